// Micro-benchmark behind DESIGN.md §3.7: what does ONE tcgen05.mma of M = 128, K = 16 (kind::f16) cost as a function of
// N, of where the A operand lives (shared memory in the no-swizzle "planes" layout of the IMPALA trunk, shared memory in
// the 128-byte swizzled layout, tensor memory) and of whether consecutive MMAs write the same accumulator?  One CTA per
// SM, one thread issues `n_mma` MMAs back to back and commits; cycles = clock64 around issue + completion.  Operand
// contents are irrelevant (zeros).
// nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o build/probe_umma scripts/probe_umma.cu && ./build/probe_umma
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); return 2; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok, spins = 0;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (!ok && ++spins > (1u << 24)) __trap();
    } while (!ok);
}
__device__ __forceinline__ uint64_t desc_nosw(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024u >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ uint32_t idesc_f16(int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24); }

// mode 0: A = planes (no swizzle, LBO = plane, SBO = 128), start shifted per MMA like a filter tap
// mode 1: A = 128-byte swizzled K-major tile          mode 2: A in tensor memory
// same_acc: every MMA accumulates onto accumulator 0; else round-robin over n_acc accumulators
__global__ void __launch_bounds__(128, 1) umma_kernel(int mode, int N, int n_mma, int same_acc, int tf32, long long* cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_s;
    const int tid = threadIdx.x;
    for (int i = tid; i < 98304 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_s)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_s, s0 = smem_u32(smem);
    if (tid == 0) {
        // kind::tf32: K = 8 per MMA (format fields 2 / 2), same descriptors (contents are zeros)
        const uint32_t idesc = tf32 ? ((1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24)) : idesc_f16(N);
        const uint32_t a_base = s0, b_base = s0 + 65536;          // A: 64 KB region, B: [N <= 256 rows x 64 k] swizzled (32 KB)
        const int n_acc = same_acc ? 1 : (448 / N > 8 ? 8 : 448 / N);
        // eight (accumulator, A descriptor, B descriptor) triples prepared up front: the timed loop is eight MMAs and a
        // branch, so what is measured is the tensor core, not the issuing thread's address arithmetic
        uint32_t dd[8];
        uint64_t ad[8], bd[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            dd[u] = tmem + (uint32_t)((u % n_acc) * N);
            bd[u] = desc_sw128(b_base) + (uint64_t)((u & 3) * 2);
            ad[u] = mode == 0 ? desc_nosw(a_base + (uint32_t)((u * 35) * 16), 18496, 128)
                              : desc_sw128(a_base + (uint32_t)((u & 3) * 16384)) + (uint64_t)(((u >> 2) & 1) * 2);
        }
        const long long t0 = clock64();
        if (mode == 2 && tf32) {
#pragma unroll 1
            for (int i = 0; i < n_mma; i += 8) {
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(dd[u]),
                                 "r"(tmem + 448u), "l"(bd[u]), "r"(idesc), "r"(1u) : "memory");
            }
        } else if (mode == 2) {
#pragma unroll 1
            for (int i = 0; i < n_mma; i += 8) {
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(dd[u]),
                                 "r"(tmem + 448u), "l"(bd[u]), "r"(idesc), "r"(1u) : "memory");
            }
        } else {
#pragma unroll 1
            for (int i = 0; i < n_mma; i += 8) {
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(dd[u]),
                                 "l"(ad[u]), "l"(bd[u]), "r"(idesc), "r"(1u) : "memory");
            }
        }
        const long long t1 = clock64();
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        mbar_wait(smem_u32(&bar), 0);
        const long long t2 = clock64();
        cycles[2 * blockIdx.x] = t1 - t0;
        cycles[2 * blockIdx.x + 1] = t2 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

int main() {
    int dev = 0, sms = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    long long* d_cyc;
    CK(cudaMalloc(&d_cyc, sizeof(long long) * 2 * sms));
    const int smem = 98304 + 1024;
    CK(cudaFuncSetAttribute(umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const char* names[3] = {"A smem planes (no swizzle, shifted start)", "A smem 128B-swizzled", "A in TMEM"};
    const int n_mma = 2048;
    printf("tcgen05.mma kind::f16, M = 128, K = 16, %d MMAs issued by one thread per CTA; cycles per MMA (issue loop | until the commit lands)\n", n_mma);
    for (int grid = 1; grid <= sms; grid = grid == 1 ? sms : sms + 1)
        for (int mode = 0; mode < 3; ++mode)
            for (int same = 0; same < 2; ++same) {
                printf("grid %3d  %-42s %s:", grid, names[mode], same ? "same accumulator " : "round-robin accs ");
                for (int N = 16; N <= 256; N *= 2) {
                    umma_kernel<<<grid, 128, smem>>>(mode, N, n_mma, same, 0, d_cyc);
                    CK(cudaDeviceSynchronize());
                    long long h[2 * 256];
                    CK(cudaMemcpy(h, d_cyc, sizeof(long long) * 2 * grid, cudaMemcpyDeviceToHost));
                    double a = 0, b = 0;
                    for (int i = 0; i < grid; ++i) { a += h[2 * i]; b += h[2 * i + 1]; }
                    printf("  N=%3d %5.1f | %5.1f", N, a / grid / n_mma, b / grid / n_mma);
                }
                printf("\n");
            }
    printf("kind::tf32 (K = 8 per MMA), A in TMEM, %d SMs:", sms);
    for (int N = 16; N <= 256; N *= 2) {
        umma_kernel<<<sms, 128, smem>>>(2, N, n_mma, 0, 1, d_cyc);
        CK(cudaDeviceSynchronize());
        long long h[2 * 256];
        CK(cudaMemcpy(h, d_cyc, sizeof(long long) * 2 * sms, cudaMemcpyDeviceToHost));
        double b = 0;
        for (int i = 0; i < sms; ++i) b += h[2 * i + 1];
        printf("  N=%3d %5.1f", N, b / sms / n_mma);
    }
    printf("\n");
    cudaFree(d_cyc);
    return 0;
}
