// Shared pieces of the "direct-from-table" tcgen05 forwards (csrc/mlp_forward_direct.cu, csrc/cnn_forward_tc.cu): a layer
// that is linear in its weights is evaluated as x.theta^T + s*(x.(sigma*eps)^T), both terms as kind::f16 MMAs whose weight
// operands arrive by TMA straight from an fp16 copy of theta and from the sigma-scaled fp16 mirror of the noise table
// (dfd_table_build_scaled16) - no thread ever touches a weight.
#pragma once
#include "tc_common.cuh"
#include <cuda.h>
#include <cuda_fp16.h>

namespace {

__device__ __forceinline__ void dr_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok, spins = 0;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(ok) : "r"(bar), "r"(parity), "r"(200000u) : "memory");
        if (!ok && ++spins > (1u << 22)) __trap();      // a lost arrival must fault, not hang the GPU
    } while (!ok);
}
__device__ __forceinline__ void dr_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void dr_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void dr_tma_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
                 "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void dr_tma_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst),
                 "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar) : "memory");
}
// kind::f16 (fp16 x fp16 -> fp32), A and B K-major, M = 128; bit 13 negates A
__device__ __forceinline__ uint32_t dr_idesc(int n, int negate_a) {
    return (1u << 4) | (negate_a ? (1u << 13) : 0u) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void dr_umma_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void dr_umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void dr_tmem_ld32(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void dr_tmem_st16(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
// two floats -> packed halves, `lo` in the low 16 bits (the even k of the pair)
__device__ __forceinline__ uint32_t dr_pack(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
template <bool APPROX>
__device__ __forceinline__ float dr_tanh(float x) {
    if (APPROX) {
        float y;
        asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
        return y;
    } else {
        return tanh_fast(x);
    }
}

__global__ void theta_to_f16_kernel(const float* __restrict__ theta, __half* __restrict__ out, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = __float2half_rn(theta[i]);
}

typedef CUresult (*dr_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
dr_encode_fn dr_encoder() {
    static dr_encode_fn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) {
            cudaGetLastError();
            return nullptr;
        }
        encode = (dr_encode_fn)fn;
    }
    return encode;
}


// fp16 copy of theta into the context's registered scratch (one small kernel per forward call: theta changes every step)
inline int dr_theta16(dfd_ctx* ctx, const float* theta, int64_t n, cudaStream_t st) {
    theta_to_f16_kernel<<<(int)((n + 1023) / 1024), 256, 0, st>>>(theta, (__half*)ctx->theta16, n);
    ctx->launches++;
    if (cudaPeekAtLastError() != cudaSuccess) {
        dfd_set_error("theta_to_f16_kernel launch failed");
        return 3;
    }
    return 0;
}

// 2-D map over the fp16 theta copy: [rows, cols] row-major at element offset `off`, box {64, box_rows}, SWIZZLE_128B
inline int dr_map_w(CUtensorMap* map, dfd_ctx* ctx, int64_t off, int cols, int rows, int box_rows) {
    dr_encode_fn encode = dr_encoder();
    if (!encode) return 1;
    const cuuint32_t estr[2] = {1, 1};
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
    const cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
    return encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (__half*)ctx->theta16 + off, dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS ? 0 : 2;
}

// 4-D map over the scaled table mirror: a member's [rows x cols] matrix starts at table element s = idx + off: coordinates
// {k, s >> 3, row, s & 7}; dims {cols, starts, rows, 8 replicas}, strides {16 B, cols * 2 B, replica bytes} - the second
// dimension overlaps the first, which is what makes an arbitrary start a legal box
inline int dr_map_e(CUtensorMap* map, dfd_ctx* ctx, int cols, int rows, int box_rows) {
    dr_encode_fn encode = dr_encoder();
    if (!encode) return 1;
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const int64_t s16 = ctx->scaled16_stride;
    const cuuint64_t starts = (cuuint64_t)((s16 - (int64_t)cols * rows) / 8);
    const cuuint64_t dims[4] = {(cuuint64_t)cols, starts, (cuuint64_t)rows, 8};
    const cuuint64_t strides[3] = {16, (cuuint64_t)cols * 2, (cuuint64_t)s16 * 2};
    const cuuint32_t box[4] = {64u, 1, (cuuint32_t)box_rows, 1};
    return encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, ctx->scaled16, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS ? 0 : 2;
}

}  // namespace
