// Hardware probe for the "direct-from-table" forward (csrc/mlp_forward_direct.cu), run once on a B200:
//   (1) cuTensorMapEncodeTiled accepts a 4-D fp16 map whose second dimension OVERLAPS the first
//       (dims {K, starts, N, replicas}, strides {16 B, K*2 B, replica bytes}) and a box {64, 1, rows, 1}
//       lands as [rows x 128 B] in the 128-byte swizzle, i.e. a member's weight tile straight from the noise table;
//   (2) tcgen05.mma kind::f16 with the A operand in TENSOR MEMORY as packed halves (two per 32-bit column,
//       even k in the low half), the negate-A bit of the instruction descriptor, B from that TMA tile.
// nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/probe_direct scripts/probe_direct.cu -lcuda
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); return 2; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok, spins = 0;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (!ok && ++spins > (1u << 22)) __trap();
    } while (!ok);
}
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024u >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

constexpr int K = 376, NROWS = 48, KC = 64;

// one CTA of 128 threads: TMA the [NROWS x 64] tile of "member" (start, replica) at k0, copy it out raw (test 1), then
// D[128 x NROWS] = (+/-)A[128 x 64] * tile^T with A in TMEM (test 2)
__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap map, int k0, int start8, int replica, int negate, const __half* __restrict__ A,
             __half* __restrict__ tile_out, float* __restrict__ D) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[2];
    __shared__ uint32_t tmem_base_s;
    const uint32_t s0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* tile = smem_raw + (s0 - smem_u32(smem_raw));
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t bar = smem_u32(&bars[0]), bar2 = smem_u32(&bars[1]);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar2));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(128u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    if (tid == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(NROWS * 128)) : "memory");
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                     ::"r"(s0), "l"(&map), "r"(k0), "r"(start8), "r"(0), "r"(replica), "r"(bar) : "memory");
    }
    mbar_wait(bar, 0);
    // test 1: un-swizzle the tile and write it out row-major [NROWS x 64]
    for (int i = tid; i < NROWS * 8; i += 128) {
        const int r = i >> 3, c = i & 7;
        const uint4 v = *reinterpret_cast<const uint4*>(tile + r * 128 + ((c ^ (r & 7)) << 4));
        *reinterpret_cast<uint4*>(tile_out + r * 64 + c * 8) = v;
    }
    // test 2: A row `tid` (64 halves) -> 32 TMEM columns, packed (k even low)
    {
        uint32_t r[32];
        const uint32_t* a = reinterpret_cast<const uint32_t*>(A + (size_t)tid * 64);
#pragma unroll
        for (int i = 0; i < 32; ++i) r[i] = a[i];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + 64u;    // A at columns [64, 96)
        asm volatile(
            "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
            "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
            "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
            "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
            "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
            "r"(r[31]) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp == 0) {
        // kind::f16: c_format F32 (1 << 4), a/b format F16 (0), negate A bit 13, N >> 3 at 17, M >> 4 at 24
        const uint32_t idesc = (1u << 4) | (negate ? (1u << 13) : 0u) | ((uint32_t)(NROWS >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t bdesc = make_desc_sw128(s0);
#pragma unroll
        for (int j = 0; j < KC / 16; ++j) {
            asm volatile(
                "{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem),
                "r"(tmem + 64u + (uint32_t)(j * 8)), "l"(bdesc + (uint64_t)(j * 2)), "r"(idesc), "r"(j ? 1u : 0u) : "memory");
        }
        asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}\n" ::"r"(bar2) : "memory");
    }
    mbar_wait(bar2, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c = 0; c < NROWS; c += 16) {
        uint32_t r[16];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 16; ++i) D[(size_t)tid * NROWS + c + i] = __uint_as_float(r[i]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u) : "memory");
    }
}

typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                              const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    const int64_t size = 1 << 20, stride = size + 64;       // halves per replica
    std::vector<__half> table(size);
    for (int64_t j = 0; j < size; ++j) table[j] = __float2half((float)((j * 7919) % 2039) / 256.0f - 4.0f);
    std::vector<__half> rep(8 * stride, __float2half(0.f));
    for (int s = 0; s < 8; ++s)
        for (int64_t j = 0; j + s < size; ++j) rep[s * stride + j] = table[j + s];
    __half* d_rep; CK(cudaMalloc(&d_rep, rep.size() * 2)); CK(cudaMemcpy(d_rep, rep.data(), rep.size() * 2, cudaMemcpyHostToDevice));
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    encode_fn encode = (encode_fn)fn;
    CUtensorMap map;
    // order A: {K, starts (16 B apart), rows (K*2 B apart), replicas}
    const cuuint64_t dims[4] = {(cuuint64_t)K, (cuuint64_t)((size - (int64_t)NROWS * K) / 8), (cuuint64_t)NROWS, 8};
    const cuuint64_t strides[3] = {16, (cuuint64_t)K * 2, (cuuint64_t)stride * 2};
    const cuuint32_t box[4] = {KC, 1, NROWS, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, d_rep, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode overlapping 4-D map {K, starts, rows, replicas}: CUresult %d\n", (int)r);
    if (r != CUDA_SUCCESS) {
        // order B: {K, rows, starts, replicas} (second stride smaller than the first)
        const cuuint64_t dimsB[4] = {(cuuint64_t)K, (cuuint64_t)NROWS, dims[1], 8};
        const cuuint64_t stridesB[3] = {(cuuint64_t)K * 2, 16, (cuuint64_t)stride * 2};
        const cuuint32_t boxB[4] = {KC, NROWS, 1, 1};
        r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, d_rep, dimsB, stridesB, boxB, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode order B {K, rows, starts, replicas}: CUresult %d\n", (int)r);
        printf("PROBE FAIL (no overlapping map)\n");
        return 1;
    }
    std::vector<__half> A(128 * 64);
    for (int i = 0; i < 128 * 64; ++i) A[i] = __float2half((float)((i * 31) % 97) / 64.0f - 0.75f);
    __half *d_A, *d_tile; float* d_D;
    CK(cudaMalloc(&d_A, A.size() * 2)); CK(cudaMemcpy(d_A, A.data(), A.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_tile, NROWS * 64 * 2)); CK(cudaMalloc(&d_D, 128 * NROWS * 4));
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, NROWS * 128 + 2048));
    int bad_total = 0;
    const int64_t idxs[4] = {0, 12345, 777777, 500003};
    for (int t = 0; t < 8; ++t) {
        const int64_t idx = idxs[t & 3] + (t >> 2) * 3;     // member's row starts at table[idx + off]
        const int64_t off = 1024;                         // a layer offset (multiple of 8)
        const int k0 = (t & 1) ? 320 : 64;                // 320: the last chunk, columns 376..383 are out of bounds -> zeros
        const int negate = t & 1;
        const int64_t s = idx + off;
        probe_kernel<<<1, 128, NROWS * 128 + 2048>>>(map, k0, (int)((s & ~7LL) / 8), (int)(s & 7), negate, d_A, d_tile, d_D);
        CK(cudaDeviceSynchronize());
        std::vector<__half> tile(NROWS * 64); std::vector<float> D(128 * NROWS);
        CK(cudaMemcpy(tile.data(), d_tile, tile.size() * 2, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(D.data(), d_D, D.size() * 4, cudaMemcpyDeviceToHost));
        int bad = 0; double maxerr = 0;
        std::vector<float> B(NROWS * 64);
        for (int n = 0; n < NROWS; ++n)
            for (int k = 0; k < 64; ++k) {
                const float want = (k0 + k < K) ? __half2float(table[s + (int64_t)n * K + k0 + k]) : 0.f;
                B[n * 64 + k] = want;
                if (__half2float(tile[n * 64 + k]) != want) { if (bad < 4) printf("  tile mismatch n %d k %d: got %f want %f\n", n, k, __half2float(tile[n * 64 + k]), want); ++bad; }
            }
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < NROWS; ++n) {
                double acc = 0;
                for (int k = 0; k < 64; ++k) acc += (double)__half2float(A[m * 64 + k]) * B[n * 64 + k];
                if (negate) acc = -acc;
                const double e = fabs(acc - D[m * NROWS + n]);
                if (e > maxerr) maxerr = e;
            }
        printf("case %d idx %lld k0 %d negate %d: tile mismatches %d, mma max err %.3e\n", t, (long long)idx, k0, negate, bad, maxerr);
        bad_total += bad + (maxerr > 1e-2 ? 1 : 0);
    }
    printf(bad_total ? "PROBE FAIL\n" : "PROBE OK\n");
    return bad_total ? 1 : 0;
}
