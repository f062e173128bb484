// Micro-benchmark for the direct-from-table forwards: how many bytes per cycle can every SM of a B200 pull from L2 into
// shared memory by TMA when ALL SMs pull at once, (a) unicast, (b) with the two CTAs of a cluster each fetching half a
// tile and multicasting it to both (.multicast::cluster).  The source region is small (L2-resident), tiles are 16 KB.
// nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o build/probe_multicast scripts/probe_multicast.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); return 2; } } while (0)
constexpr int NS = 8, TILE = 16384;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok, spins = 0;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (!ok && ++spins > (1u << 24)) __trap();
    } while (!ok);
}

template <int CL>
__global__ void __launch_bounds__(64, 1) pull_kernel(const uint8_t* __restrict__ src, int src_tiles, int n_tiles, int same_for_all, long long* cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t full[NS], empty[NS];
    const int tid = threadIdx.x;
    uint32_t rank = 0;
    if (CL > 1) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    if (tid == 0) {
        for (int s = 0; s < NS; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[s])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&empty[s])), "r"(CL));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (CL > 1) { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
    else __syncthreads();
    const long long t0 = clock64();
    const int cluster_id = blockIdx.x / CL;
    if (tid == 0) {                    // producer
        for (int g = 0; g < n_tiles; ++g) {
            const int slot = g % NS;
            mbar_wait(smem_u32(&empty[slot]), (uint32_t)((g / NS) & 1) ^ 1u);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[slot])), "r"((uint32_t)TILE) : "memory");
            const int t = same_for_all ? (g % src_tiles) : ((g + cluster_id * 7) % src_tiles);
            const uint8_t* s = src + (size_t)t * TILE;
            const uint32_t dst = smem_u32(smem) + slot * TILE;
            if (CL == 1) {
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(s), "r"((uint32_t)TILE),
                             "r"(smem_u32(&full[slot])) : "memory");
            } else {
                const uint32_t part = TILE / CL;
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(dst + rank * part),
                             "l"(s + rank * part), "r"(part), "r"(smem_u32(&full[slot])), "h"((uint16_t)((1u << CL) - 1)) : "memory");
            }
        }
    } else if (tid == 32) {            // consumer: frees the slot in every CTA of the cluster
        for (int g = 0; g < n_tiles; ++g) {
            const int slot = g % NS;
            mbar_wait(smem_u32(&full[slot]), (uint32_t)((g / NS) & 1));
            if (CL == 1) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[slot])) : "memory");
            else {
                for (uint32_t r = 0; r < (uint32_t)CL; ++r) {
                    uint32_t remote;
                    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(&empty[slot])), "r"(r));
                    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
                }
            }
        }
    }
    __syncthreads();
    if (CL > 1) { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
    if (tid == 0) cycles[blockIdx.x] = clock64() - t0;
}

template <int CL>
int run(const uint8_t* src, int src_tiles, int n_tiles, int same, long long* d_cyc, const char* what) {
    CK(cudaFuncSetAttribute(pull_kernel<CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, NS * TILE + 1024));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(148); cfg.blockDim = dim3(64); cfg.dynamicSmemBytes = NS * TILE + 1024;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(a);
        CK(cudaLaunchKernelEx(&cfg, pull_kernel<CL>, src, src_tiles, n_tiles, same, d_cyc));
        cudaEventRecord(b);
        CK(cudaDeviceSynchronize());
    }
    float ms; cudaEventElapsedTime(&ms, a, b);
    long long h[148]; cudaMemcpy(h, d_cyc, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
    const double bytes = (double)n_tiles * TILE;
    printf("%-46s cluster %d: %.3f ms, %.1f B/cycle/SM delivered, %.2f TB/s delivered to all SMs (L2 reads: %.2f TB/s)\n", what, CL, ms, bytes / (double)mx,
           bytes * 148 / (ms * 1e-3) * 1e-12, bytes * 148 / CL / (ms * 1e-3) * 1e-12);
    return 0;
}

int main() {
    const int src_tiles = 42;             // 688 KB: the fp16 theta copy of the C3 policy, L2-resident
    uint8_t* src; CK(cudaMalloc(&src, (size_t)src_tiles * TILE)); CK(cudaMemset(src, 1, (size_t)src_tiles * TILE));
    long long* d_cyc; CK(cudaMalloc(&d_cyc, 148 * 8));
    const int n = 4096;
    if (run<1>(src, src_tiles, n, 1, d_cyc, "unicast, every SM the same tile sequence")) return 1;
    if (run<1>(src, src_tiles, n, 0, d_cyc, "unicast, sequences offset per SM pair")) return 1;
    if (run<2>(src, src_tiles, n, 1, d_cyc, "multicast halves, same sequence everywhere")) return 1;
    if (run<2>(src, src_tiles, n, 0, d_cyc, "multicast halves, sequences offset per cluster")) return 1;
    if (run<4>(src, src_tiles, n, 0, d_cyc, "multicast quarters, sequences offset per cluster")) return 1;
    // a large source (HBM-bound): 4 GB walked once per launch by all SMs together is too slow; use 1.5 GB, distinct tiles per SM
    printf("done\n");
    return 0;
}
