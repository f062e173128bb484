// Dense tail of the IMPALA forward (Linear 2048 -> 256 + ReLU, LSTM cell 513 -> 1024, BN + policy head, softmax;
// policies/impala.py:153-186) as TMA-fed tcgen05 "swap AB" GEMMs, shared by the two IMPALA kernels (impala_forward.cu: mma.sync
// trunk; impala_forward_tc.cu: tcgen05 trunk).  The WEIGHT tiles [128 rows x 64 k] are the A operand and arrive by TMA
// straight from an fp16 repack of theta and from the sigma-scaled fp16 mirror of the noise table (direct_common.cuh:
// x.(theta + s*sigma*eps)^T = x.theta^T + s*(x.(sigma*eps)^T): the perturbed weights are never built); the activations of
// the CTA's one or two members are the B operand (N = 16 columns, column n = member n).  weight_ih rows are 257 wide (not
// 16-byte aligned): rows r = 8q + c form class c, 8 * 257 elements apart - a legal TMA stride - so an M tile is a class
// (gate row of lane q = 8q + c) and the reward column is added in the epilogue in fp32.
#pragma once
#include "direct_common.cuh"
#include "impala_layout.cuh"

namespace {

constexpr int TL_NSLOT = 8;                                  // ring of 16 KB weight tiles at [0, 131072) of the CTA's shared memory
constexpr int TL_CT = 131072, TL_HT = TL_CT + 5120;          // core / h0 tiles: 4 boxes x 1 KB (+ 1 KB the last N = 16 descriptor overhangs)
constexpr int TL_XT_BYTES = 33 * 1024;                       // x tiles: 32 boxes x 1 KB (+ 1 KB)
constexpr int ST = 260 + 256 + 1024 + 256 + 32;              // per-member fp32 scratch: core 260 | h0 256 | gates 1024 | hn 256 | logits 32
constexpr uint32_t TC_FW = 0, TC_FE = 32, TC_GW = 96, TC_GE = 224;   // TMEM columns: FC theta / eps parts, gates theta / eps parts
enum { TB_MMA = 0, TB_GO = 1, TB_FULL = 2, TB_EMPTY = TB_FULL + TL_NSLOT, TB_COUNT = TB_EMPTY + TL_NSLOT };

struct ItMaps {
    CUtensorMap w_fc, e_fc, w_ih, e_ih, w_hh, e_hh;
};
struct TailArgs {
    const float* theta; const float* bnbuf; const float* reward; const uint8_t* done; const float* h_in; const float* c_in;
    float* probs; float* h_out; float* c_out;
    float sigma;
    int nmem, nE;
    int inst[2], sgi[2];
    int64_t ids[2];
    const float* rows[2];
};

template <int NWT>
__device__ __forceinline__ void tl_wsync() { asm volatile("bar.sync 1, %0;" ::"n"(NWT) : "memory"); }
__device__ __forceinline__ float tl_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }

// fp16 repack of the dense-tail weights into the context's theta16 scratch, 16-byte aligned blocks (the flat offsets of
// fc.weight / weight_ih / weight_hh are 6 mod 8): [256 x 2048] | weight_ih[:, :256] as [1024 x 256] | [1024 x 256];
// behind them (TL_CONV16, when the scratch is large enough: with_conv) the convolution weights of all 15 layers in the
// tcgen05 trunk's K order [oc][tap][ci] (theta keeps them as [oc][ci][tap]), segment i at TL_CONV16 + L.seq_o[i]
constexpr int TL_CONV16 = 256 * 2048 + 2 * 1024 * 256;
__global__ void impala_theta16_kernel(const float* __restrict__ theta, __half* __restrict__ out, int fc_w, int wih, int whh,
                                      const __grid_constant__ ImpalaP L, int with_conv) {
    const int n = TL_CONV16;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float v;
        if (i < 524288) v = theta[fc_w + i];
        else if (i < 786432) { const int j = i - 524288; v = theta[wih + (j >> 8) * 257 + (j & 255)]; }
        else v = theta[whh + (i - 786432)];
        out[i] = __float2half_rn(v);
    }
    if (!with_conv) return;
    const int nc = L.seq_o[15];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nc; i += gridDim.x * blockDim.x) {
        int li = 0;
        while (li < 14 && i >= L.seq_o[li + 1]) ++li;
        const int j = i - L.seq_o[li], cin = L.seq_cin[li], k9 = cin * 9;
        const int oc = j / k9, r = j - oc * k9, tap = r / cin, ci = r - tap * cin;
        out[TL_CONV16 + i] = __float2half_rn(theta[L.seq_w[li] + oc * k9 + ci * 9 + tap]);
    }
}

// Executed by the NWT worker threads (warps 0 .. NWT/32 - 1; at least 12 warps).  sm / s0: the CTA's 1024-aligned shared
// memory (generic / shared address), XT_OFF: byte offset of the x tiles, fcin: fp16 [2][2048] BN'd trunk outputs (outside
// [0, TL_HT + 5120) and the x tiles), bar0: shared address of the TB_COUNT tail barriers, mph: parity of TB_MMA.
template <int NWT, int XT_OFF>
__device__ void tl_dense_tail_workers(const ImpalaP& L, const TailArgs& A, uint8_t* sm, uint32_t s0, const __half* fcin, uint32_t bar0,
                                      uint32_t tmem, uint32_t mph) {
#define TL_BAR(i) (bar0 + 8u * (uint32_t)(i))
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int q4 = warp & 3, wg = warp >> 2;
    const uint32_t lane_sel = (uint32_t)(q4 * 32) << 16;
    const float* theta = A.theta; const float* bnbuf = A.bnbuf; const float* reward = A.reward; const uint8_t* done = A.done;
    const float* h_in = A.h_in; const float* c_in = A.c_in; float* probs = A.probs; float* h_out = A.h_out; float* c_out = A.c_out;
    const float sigma = A.sigma;
    const int nmem = A.nmem, nE = A.nE;
    const int* inst = A.inst; const int* sgi = A.sgi; const float* const* rows = A.rows;
        uint8_t* xt = sm + XT_OFF;
        uint8_t* ct = sm + TL_CT;
        uint8_t* ht = sm + TL_HT;
        float* st = reinterpret_cast<float*>(sm);                 // fp32 scratch once the ring is dead
        for (int i = tid; i < (33 * 1024) / 16; i += NWT) reinterpret_cast<uint4*>(xt)[i] = make_uint4(0, 0, 0, 0);
        for (int i = tid; i < (10 * 1024) / 16; i += NWT) reinterpret_cast<uint4*>(ct)[i] = make_uint4(0, 0, 0, 0);
        tl_wsync<NWT>();
        // B operands: column n = member n; element (n, k) of box k >> 6 at n * 128 + (((k & 63) >> 3) ^ n) * 16 + (k & 7) * 2
        for (int i = tid; i < nmem * 2048; i += NWT) {
            const int n = i >> 11, k = i & 2047;
            *reinterpret_cast<__half*>(xt + (k >> 6) * 1024 + n * 128 + ((((k & 63) >> 3) ^ n) << 4) + (k & 7) * 2) = fcin[i];
        }
        for (int i = tid; i < nmem * 256; i += NWT) {
            const int n = i >> 8, k = i & 255;
            const float hv = done[inst[n]] ? 0.f : h_in[(int64_t)inst[n] * 256 + k];
            *reinterpret_cast<__half*>(ht + (k >> 6) * 1024 + n * 128 + ((((k & 63) >> 3) ^ n) << 4) + (k & 7) * 2) = __float2half_rn(hv);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        tl_wsync<NWT>();
        if (tid == 0) dr_arrive(TL_BAR(TB_GO));                    // the ring region is free: the producer may start
        const uint32_t id16 = dr_idesc(16, 0);
        // ---- Linear 2048 -> 256: tiles (kb, part, mt) ----
        int g = 0;
        if (warp == 0) {
#pragma unroll 1
            for (int kb = 0; kb < 32; ++kb) {
                const uint64_t bdesc = make_desc_sw128(s0 + XT_OFF + (uint32_t)kb * 1024u);
#pragma unroll 1
                for (int part = 0; part < 1 + nE; ++part)
#pragma unroll 1
                    for (int mt = 0; mt < 2; ++mt, ++g) {
                        const int slot = g % TL_NSLOT;
                        dr_wait(TL_BAR(TB_FULL + slot), (uint32_t)((g / TL_NSLOT) & 1));
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint64_t adesc = make_desc_sw128(s0 + (uint32_t)slot * 16384u);
                        const uint32_t d = tmem + (part == 0 ? TC_FW + (uint32_t)(mt * 16) : TC_FE + (uint32_t)(((part - 1) * 2 + mt) * 16));
#pragma unroll
                        for (int j = 0; j < 4; ++j) dr_umma_ss(d, adesc + (uint64_t)(j * 2), bdesc + (uint64_t)(j * 2), id16, (kb | j) ? 1u : 0u);
                        umma_commit_elect(TL_BAR(TB_EMPTY + slot));
                    }
            }
            umma_commit_elect(TL_BAR(TB_MMA));
        }
        if (lane == 0) dr_wait(TL_BAR(TB_MMA), mph);
        __syncwarp();
        mph ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // FC epilogue: thread = neuron o; relu(y + bias) -> core tile (column n, k = o) and the fp32 scratch is not needed
        if (warp < 8) {
            const int mt = warp >> 2, o = mt * 128 + q4 * 32 + lane;
            float vw[16], ve0[16], ve1[16];
            tmem_ld16(tmem + lane_sel + TC_FW + (uint32_t)(mt * 16), vw);
            tmem_ld16(tmem + lane_sel + TC_FE + (uint32_t)(mt * 16), ve0);
            if (nE == 2) tmem_ld16(tmem + lane_sel + TC_FE + (uint32_t)((2 + mt) * 16), ve1);
            for (int n = 0; n < nmem; ++n) {
                const float ev = (nE == 2 && n == 1) ? ve1[n] : ve0[n];
                const float b = perturb1(theta[L.fc_b + o], sigma * (float)sgi[n], rows[n][L.fc_b + o]);
                const float y = fmaxf(vw[n] + (float)sgi[n] * ev + b, 0.f);
                *reinterpret_cast<__half*>(ct + (o >> 6) * 1024 + n * 128 + ((((o & 63) >> 3) ^ n) << 4) + (o & 7) * 2) = __float2half_rn(y);
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        tl_wsync<NWT>();
        // ---- LSTM gates: classes c (gate row of lane q = 8 q + c), sources weight_ih[:, :256] . core and weight_hh . h0 ----
        if (warp == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int cc = 0; cc < 8; ++cc)
#pragma unroll 1
                for (int src = 0; src < 2; ++src)
#pragma unroll 1
                    for (int kb = 0; kb < 4; ++kb) {
                        const uint64_t bdesc = make_desc_sw128(s0 + (src ? TL_HT : TL_CT) + (uint32_t)kb * 1024u);
#pragma unroll 1
                        for (int part = 0; part < 1 + nE; ++part, ++g) {
                            const int slot = g % TL_NSLOT;
                            dr_wait(TL_BAR(TB_FULL + slot), (uint32_t)((g / TL_NSLOT) & 1));
                            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                            const uint64_t adesc = make_desc_sw128(s0 + (uint32_t)slot * 16384u);
                            const uint32_t d = tmem + (part == 0 ? TC_GW + (uint32_t)(cc * 16) : TC_GE + (uint32_t)(((part - 1) * 8 + cc) * 16));
#pragma unroll
                            for (int j = 0; j < 4; ++j) dr_umma_ss(d, adesc + (uint64_t)(j * 2), bdesc + (uint64_t)(j * 2), id16, (src | kb | j) ? 1u : 0u);
                            umma_commit_elect(TL_BAR(TB_EMPTY + slot));
                        }
                    }
            umma_commit_elect(TL_BAR(TB_MMA));
        }
        if (lane == 0) dr_wait(TL_BAR(TB_MMA), mph);
        __syncwarp();
        mph ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // gates epilogue: warp -> lane quarter q4, classes wg and wg + 4; + reward column, biases (exact fp32)
        for (int cc = wg; cc < 8 && wg < 3; cc += 3) {
            const int r = 8 * (q4 * 32 + lane) + cc;
            float vw[16], ve0[16], ve1[16];
            tmem_ld16(tmem + lane_sel + TC_GW + (uint32_t)(cc * 16), vw);
            tmem_ld16(tmem + lane_sel + TC_GE + (uint32_t)(cc * 16), ve0);
            if (nE == 2) tmem_ld16(tmem + lane_sel + TC_GE + (uint32_t)((8 + cc) * 16), ve1);
            for (int n = 0; n < nmem; ++n) {
                const float sgn = sigma * (float)sgi[n];
                const float* rw = rows[n];
                const float ev = (nE == 2 && n == 1) ? ve1[n] : ve0[n];
                const float rwd = fminf(fmaxf(reward[inst[n]], -1.f), 1.f);           // clamp(reward, -1, 1), impala.py:158
                const int64_t pr = L.wih + (int64_t)r * 257 + 256;
                st[n * ST + 516 + r] = vw[n] + (float)sgi[n] * ev + perturb1(theta[pr], sgn, rw[pr]) * rwd +
                                       perturb1(theta[L.bih + r], sgn, rw[L.bih + r]) + perturb1(theta[L.bhh + r], sgn, rw[L.bhh + r]);
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        tl_wsync<NWT>();
        // ---- LSTM cell, policy head (fp32, exactly perturbed parameters), softmax ----
        for (int n = 0; n < nmem; ++n) {
            const float sgn = sigma * (float)sgi[n];
            const float* rw = rows[n];
            const bool dn = done[inst[n]] != 0;
            const float* gates = st + n * ST + 516;
            float* hn = st + n * ST + 1540;
            for (int k = tid; k < 256; k += NWT) {
                const float c0 = dn ? 0.f : c_in[(int64_t)inst[n] * 256 + k];
                const float ig = tl_sigmoid(gates[k]), fg = tl_sigmoid(gates[256 + k]);
                const float gg = tanhf(gates[512 + k]), og = tl_sigmoid(gates[768 + k]);
                const float c1 = fg * c0 + ig * gg;
                const float h1 = og * tanhf(c1);
                c_out[(int64_t)inst[n] * 256 + k] = c1;
                h_out[(int64_t)inst[n] * 256 + k] = h1;
                const float inv = 1.0f / sqrtf(bnbuf[L.pol_bv + k] + 1e-5f);
                const float sc = perturb1(theta[L.pol_g + k], sgn, rw[L.pol_g + k]) * inv;
                hn[k] = fmaf(h1, sc, perturb1(theta[L.pol_be + k], sgn, rw[L.pol_be + k]) - bnbuf[L.pol_bm + k] * sc);
            }
        }
        tl_wsync<NWT>();
        for (int a = warp; a < nmem * L.A; a += (NWT / 32)) {
            const int n = a / L.A, ai = a - n * L.A;
            const float sgn = sigma * (float)sgi[n];
            const float* rw = rows[n];
            const float* hn = st + n * ST + 1540;
            float acc = 0.f;
            for (int k = lane; k < 256; k += 32) acc = fmaf(perturb1(theta[L.pol_w + ai * 256 + k], sgn, rw[L.pol_w + ai * 256 + k]), hn[k], acc);
            acc = warp_sum(acc);
            if (lane == 0) st[n * ST + 1796 + ai] = acc + perturb1(theta[L.pol_b + ai], sgn, rw[L.pol_b + ai]);
        }
        tl_wsync<NWT>();
        if (tid < nmem) {
            const float* lg = st + tid * ST + 1796;
            float mx = -INFINITY;
            for (int a = 0; a < L.A; ++a) mx = fmaxf(mx, lg[a]);
            float ssum = 0.f;
            for (int a = 0; a < L.A; ++a) ssum += expf(lg[a] - mx);
            const float inv = 1.0f / ssum;
            for (int a = 0; a < L.A; ++a) probs[(int64_t)inst[tid] * L.A + a] = expf(lg[a] - mx) * inv;
        }
#undef TL_BAR
}

// Executed by ONE thread outside the worker set: streams the weight tiles in the order the MMA issue loop consumes them
__device__ void tl_dense_tail_producer(const ImpalaP& L, const ItMaps& maps, const TailArgs& A, uint32_t s0, uint32_t bar0) {
#define TL_BAR(i) (bar0 + 8u * (uint32_t)(i))
    const int nE = A.nE;
    const int64_t* ids = A.ids;
        // =============================== TMA producer of the dense tail ===============================
        dr_wait(TL_BAR(TB_GO), 0);
        int g = 0;
        auto put = [&](const CUtensorMap* map, int rank, int c0, int c1, int c2, int c3) {
            const int slot = g % TL_NSLOT;
            dr_wait(TL_BAR(TB_EMPTY + slot), (uint32_t)((g / TL_NSLOT) & 1) ^ 1u);
            dr_expect_tx(TL_BAR(TB_FULL + slot), 16384u);
            const uint32_t dst = s0 + (uint32_t)slot * 16384u;
            if (rank == 4) dr_tma_4d(dst, map, c0, c1, c2, c3, TL_BAR(TB_FULL + slot));
            else if (rank == 2) dr_tma_2d(dst, map, c0, c1, TL_BAR(TB_FULL + slot));
            else asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
                              "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(TL_BAR(TB_FULL + slot)) : "memory");
            ++g;
        };
#pragma unroll 1
        for (int kb = 0; kb < 32; ++kb)
#pragma unroll 1
            for (int part = 0; part < 1 + nE; ++part)
#pragma unroll 1
                for (int mt = 0; mt < 2; ++mt) {
                    if (part == 0) put(&maps.w_fc, 2, kb * 64, mt * 128, 0, 0);
                    else {
                        const int64_t s = ids[part - 1] + L.fc_w;
                        put(&maps.e_fc, 4, kb * 64, (int)(s >> 3), mt * 128, (int)(s & 7));
                    }
                }
#pragma unroll 1
        for (int cc = 0; cc < 8; ++cc)
#pragma unroll 1
            for (int src = 0; src < 2; ++src)
#pragma unroll 1
                for (int kb = 0; kb < 4; ++kb)
#pragma unroll 1
                    for (int part = 0; part < 1 + nE; ++part) {
                        if (part == 0) put(src ? &maps.w_hh : &maps.w_ih, 3, kb * 64, 0, cc, 0);
                        else {
                            const int64_t s = ids[part - 1] + (src ? (int64_t)L.whh + 256 * cc : (int64_t)L.wih + 257 * cc);
                            put(src ? &maps.e_hh : &maps.e_ih, 4, kb * 64, (int)(s >> 3), 0, (int)(s & 7));
                        }
                    }
#undef TL_BAR
}

// 3-D map over an fp16 [1024 x 256] block of the theta scratch: {k, q (rows 8 apart), class c}: box {64, 128, 1}
inline int it_map_w3(CUtensorMap* map, const __half* base) {
    dr_encode_fn encode = dr_encoder();
    if (!encode) return 1;
    const cuuint32_t estr[3] = {1, 1, 1};
    const cuuint64_t dims[3] = {256, 128, 8};
    const cuuint64_t strides[2] = {8 * 256 * 2, 256 * 2};
    const cuuint32_t box[3] = {64, 128, 1};
    return encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<__half*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS ? 0 : 2;
}
// 4-D map over the scaled table mirror for the rows of one class of an LSTM weight matrix: 256 columns, rows `pitch`
// elements * 8 apart: {k, start, q, replica}, box {64, 1, 128, 1}
inline int it_map_e_class(CUtensorMap* map, dfd_ctx* ctx, int pitch) {
    dr_encode_fn encode = dr_encoder();
    if (!encode) return 1;
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const int64_t s16 = ctx->scaled16_stride;
    const cuuint64_t starts = (cuuint64_t)((s16 - (int64_t)1024 * pitch) / 8);
    const cuuint64_t dims[4] = {256, starts, 128, 8};
    const cuuint64_t strides[3] = {16, (cuuint64_t)pitch * 8 * 2, (cuuint64_t)s16 * 2};
    const cuuint32_t box[4] = {64, 1, 128, 1};
    return encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, ctx->scaled16, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS ? 0 : 2;
}
// the six maps of the tail + the fp16 repack of theta (one small kernel per forward call); 0 on success
inline int tl_prepare(dfd_ctx* ctx, const ImpalaP& L, const float* theta, ItMaps* maps, cudaStream_t st) {
    __half* t16 = (__half*)ctx->theta16;
    int rc = 0;
    rc |= dr_map_w(&maps->w_fc, ctx, 0, 2048, 256, 128);
    rc |= dr_map_e(&maps->e_fc, ctx, 2048, 256, 128);
    rc |= it_map_w3(&maps->w_ih, t16 + 524288);
    rc |= it_map_w3(&maps->w_hh, t16 + 786432);
    rc |= it_map_e_class(&maps->e_ih, ctx, 257);
    rc |= it_map_e_class(&maps->e_hh, ctx, 256);
    if (rc) { dfd_set_error("IMPALA tensor path: cuTensorMapEncodeTiled failed"); return 1; }
    impala_theta16_kernel<<<ctx->sm_count * 2, 512, 0, st>>>(theta, t16, L.fc_w, L.wih, L.whh, L, ctx->theta16_cap >= TL_CONV16 + L.seq_o[15] ? 1 : 0);
    ctx->launches++;
    if (cudaPeekAtLastError() != cudaSuccess) { dfd_set_error("impala_theta16_kernel launch failed"); return 3; }
    return 0;
}

}  // namespace
