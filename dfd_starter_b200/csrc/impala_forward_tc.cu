// IMPALA CNN + LSTM perturbed forward on tcgen05 / TMEM / TMA (policies/impala.py:136-186; perturbation worker/worker.py:28),
// precision level 2.  One CTA per (antithetic pair, environment): the two members' trunks run one after the other, the dense
// tail of both streams theta and the pair's shared eps row ONCE.
//
// Trunk (15 convolutions 3x3, pad 1): implicit GEMMs M = 128 output pixels, N = output channels, K = 9 * input channels
// in TAP-MAJOR order (k = tap * cin + c).  The operand of a convolution - BN(+ReLU) of the previous raw map - is written
// ONCE by the producing layer's epilogue as an fp16 CHANNEL-LAST zero-bordered map, so the im2col tile of 128 pixels is a
// handful of 16-byte copies per (pixel, tap) into the 128-byte swizzled K-major layout; the perturbed weights of a layer
// (<= 9 216 values) are built exactly in fp32 (theta + s*sigma*eps, two roundings), rounded once to fp16 and laid out as
// the B operand by the CTA.  Accumulators live in TENSOR MEMORY: the raw residual stream x of a stage stays there - the
// second convolution of a residual block ACCUMULATES onto it (x += conv(..) is the MMA's own accumulate) - and only
// BN'd / ReLU'd fp16 operand maps ever touch shared memory.  Stage convolutions are evaluated in bands of 8 rows into a
// 9-row fp32 band, max-pooled 3x3 / 2, and the pooled raw map is stored back into TMEM (tcgen05.st).
//
// Dense tail (Linear 2048 -> 256, LSTM 513 -> 1024: 90 % of the parameters): "swap AB" GEMMs whose WEIGHT tiles
// [128 rows x 64 k] are the A operand and arrive by TMA straight from an fp16 repack of theta and from the sigma-scaled
// fp16 mirror of the noise table (csrc/direct_common.cuh: x.(theta + s*sigma*eps)^T = x.theta^T + s*(x.(sigma*eps)^T), so
// the perturbed weights are never built); the activations of the pair's members are the B operand (N = 16 columns).  The
// 257-wide rows of weight_ih are not 16-byte aligned: rows r = 8q + c form class c, whose rows are 8*257 elements apart -
// a legal TMA stride - so an M tile is a class (gate row of lane q = 8q + c); the reward column is added in the epilogue.
// fp16 operands (10-bit mantissa, the precision of tf32), fp32 accumulate; BN folds, biases, LSTM cell, policy head fp32.
#include "direct_common.cuh"
#include "impala_layout.cuh"

namespace {

constexpr int IT_WORKERS = 512, IT_THREADS = IT_WORKERS + 32;
// shared memory map (bytes from the 1024-aligned base)
constexpr int IT_A = 0, IT_A_BYTES = 49152;                 // im2col tile: 3 boxes of [128 x 64] fp16
constexpr int IT_B = IT_A + IT_A_BYTES, IT_B_BYTES = 20480;  // weights: up to 5 boxes of [32 x 64]
constexpr int IT_XH_BYTES = 38400;                           // one operand map (34*34*16*2 = 36 992 is the largest)
constexpr int IT_XHA = IT_B + IT_B_BYTES, IT_XHB = IT_XHA + IT_XH_BYTES;
constexpr int IT_BAND = IT_XHB + IT_XH_BYTES, IT_BAND_BYTES = 36864;   // 9 rows x 64 x 16 (or 32 x 32) fp32
constexpr int IT_FCIN = IT_BAND + IT_BAND_BYTES;             // fp16 [2][2048]: BN'd trunk outputs of the two members
constexpr int IT_PAR = IT_FCIN + 8192;                       // floats: sN[32] tN[32] bias[32]
constexpr int IT_SMEM = IT_PAR + 512 + 1024;
// dense tail (the trunk buffers are dead by then)
constexpr int IT_NSLOT = 8;                                  // ring of 16 KB weight tiles at [0, 131072)
constexpr int IT_CT = 131072, IT_HT = IT_CT + 5120;          // core / h0 tiles: 4 boxes x 1 KB (+ 1 KB the last N = 16 descriptor overhangs)
constexpr int IT_XT = IT_BAND;                               // x tiles: 32 boxes x 1 KB (+ 1 KB)
constexpr int ST = 260 + 256 + 1024 + 256 + 32;              // per-member fp32 scratch: core 260 | h0 256 | gates 1024 | hn 256 | logits 32
// tensor memory columns
constexpr uint32_t TC_R = 0, TC_Y = 128, TC_P = 256;         // residual stream | block-internal map | stage-conv band
constexpr uint32_t TC_FW = 0, TC_FE = 32, TC_GW = 96, TC_GE = 224;   // tail: FC theta / eps parts, gates theta / eps parts

struct ItMaps {
    CUtensorMap w_fc, e_fc, w_ih, e_ih, w_hh, e_hh;
};

enum { IB_MMA = 0, IB_GO = 1, IB_FULL = 2, IB_EMPTY = IB_FULL + IT_NSLOT, IB_COUNT = IB_EMPTY + IT_NSLOT };

__device__ __forceinline__ void it_wsync() { asm volatile("bar.sync 1, 512;" ::: "memory"); }
__device__ __forceinline__ void it_tmem_st16(uint32_t taddr, const float* v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
        "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]), "f"(v[8]), "f"(v[9]), "f"(v[10]),
        "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15])
        : "memory");
}
__device__ __forceinline__ float it_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }

// fp16 repack of the dense-tail weights into the context's theta16 scratch, 16-byte aligned blocks (the flat offsets of
// fc.weight / weight_ih / weight_hh are 6 mod 8): [256 x 2048] | weight_ih[:, :256] as [1024 x 256] | [1024 x 256]
__global__ void impala_theta16_kernel(const float* __restrict__ theta, __half* __restrict__ out, int fc_w, int wih, int whh) {
    const int n = 256 * 2048 + 2 * 1024 * 256;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float v;
        if (i < 524288) v = theta[fc_w + i];
        else if (i < 786432) { const int j = i - 524288; v = theta[wih + (j >> 8) * 257 + (j & 255)]; }
        else v = theta[whh + (i - 786432)];
        out[i] = __float2half_rn(v);
    }
}

struct ItCtx {
    const float* theta;
    const float* row;
    const float* bn;
    float sg;
    __device__ __forceinline__ float par(int p) const { return perturb1(theta[p], sg, row[p]); }
};

// weights of one convolution -> B operand [cout rows x K] fp16, K-major 128-byte swizzle, k = tap * cinp + c; conv bias;
// input-side BN of the NEXT convolution folded to scale / shift (applied by this layer's epilogue)
__device__ void it_prep_layer(const ItCtx& c, const ConvP& p, int cinp, uint8_t* Bs, float* sN, float* tN, float* bias,
                              const ConvP* nxt) {
    const int tid = threadIdx.x;
    const int k9 = p.cin * 9, n = p.cout * k9;
    if (cinp != p.cin) {       // first layer: padded channel and K tail must be zeros
        for (int i = tid; i < 16 * 128 / 16; i += IT_WORKERS) reinterpret_cast<uint4*>(Bs)[i] = make_uint4(0, 0, 0, 0);
        it_wsync();
    }
#pragma unroll 1
    for (int t0 = tid; t0 < n; t0 += IT_WORKERS * 6) {
        float a[6], e[6];
#pragma unroll
        for (int u = 0; u < 6; ++u) {
            const int t = t0 + u * IT_WORKERS;
            a[u] = 0.f; e[u] = 0.f;
            if (t < n) { a[u] = c.theta[p.w + t]; e[u] = c.row[p.w + t]; }
        }
#pragma unroll
        for (int u = 0; u < 6; ++u) {
            const int t = t0 + u * IT_WORKERS;
            if (t < n) {
                const int oc = t / k9, rem = t - oc * k9, ci = rem / 9, tap = rem - ci * 9;
                const int k = tap * cinp + ci;
                *reinterpret_cast<__half*>(Bs + (k >> 6) * (p.cout * 128) + oc * 128 + ((((k & 63) >> 3) ^ (oc & 7)) << 4) + (k & 7) * 2) =
                    __float2half_rn(perturb1(a[u], c.sg, e[u]));
            }
        }
    }
    if (tid < p.cout) bias[tid] = c.par(p.b + tid);
    if (nxt != nullptr && tid >= 32 && tid < 32 + nxt->cin) {
        const int ch = tid - 32;
        const float inv = 1.0f / sqrtf(c.bn[nxt->bv + ch] + 1e-5f);
        const float s = c.par(nxt->g + ch) * inv;
        sN[ch] = s;
        tN[ch] = c.par(nxt->be + ch) - c.bn[nxt->bm + ch] * s;
    }
}

// im2col of M tile T: rows = output pixels T*128 + r of a W x W map (W = 1 << lw), chunks [q0, q0 + nq) of 8 halves:
// chunk q = tap q / cpt, channels 8 (q % cpt) .. + 8 (cpt = cin / 8) of the zero-bordered channel-last map `xh`
__device__ __forceinline__ void it_im2col(uint32_t A0, const uint8_t* xh, int lw, int cin, int lcpt, int T, int M, int q0, int nq) {
    const int W = 1 << lw, cpt = 1 << lcpt;
    for (int task = threadIdx.x; task < 128 * nq; task += IT_WORKERS) {
        const int r = task / nq, j = task - r * nq, p = T * 128 + r;
        if (p < M) {
            const int q = q0 + j, tap = q >> lcpt, part = q & (cpt - 1);
            const int dy = (tap * 11) >> 5, dx = tap - 3 * dy, y = p >> lw, x = p & (W - 1);
            const uint4 v = *reinterpret_cast<const uint4*>(xh + (((y + dy) * (W + 2) + x + dx) * cin + part * 8) * 2);
            const uint32_t d = A0 + (uint32_t)((j >> 3) * 16384 + r * 128 + (((j & 7) ^ (r & 7)) << 4));
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(d), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
        }
    }
}

// the first convolution: 4 (3 + 1 zero) channels per pixel, K = 36 padded to 48: chunk q = taps 2q, 2q + 1
__device__ __forceinline__ void it_im2col_first(uint32_t A0, const uint8_t* xh, int T) {
    for (int task = threadIdx.x; task < 128 * 6; task += IT_WORKERS) {
        const int r = task / 6, q = task - r * 6, p = T * 128 + r, y = p >> 6, x = p & 63;
        const int t0 = 2 * q, t1 = 2 * q + 1;
        const int dy0 = (t0 * 11) >> 5, dx0 = t0 - 3 * dy0, dy1 = (t1 * 11) >> 5, dx1 = t1 - 3 * dy1;
        const uint2 v0 = *reinterpret_cast<const uint2*>(xh + ((y + dy0) * 66 + x + dx0) * 8);
        uint2 v1 = make_uint2(0, 0);
        if (t1 < 9) v1 = *reinterpret_cast<const uint2*>(xh + ((y + dy1) * 66 + x + dx1) * 8);
        const uint32_t d = A0 + (uint32_t)(r * 128 + ((q ^ (r & 7)) << 4));
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(d), "r"(v0.x), "r"(v0.y), "r"(v1.x), "r"(v1.y) : "memory");
    }
}

// nk MMAs (K = 16 each) of tile A (local k from 0) against the weights from global k index kk0 on
__device__ __forceinline__ void it_mma(uint32_t d_tmem, uint32_t A0, uint32_t B0, int cout, int kk0, int nk, uint32_t idesc, bool acc_first) {
#pragma unroll 1
    for (int ks = 0; ks < nk; ++ks) {
        const int kk = kk0 + ks * 16;
        const uint64_t adesc = make_desc_sw128(A0 + (uint32_t)((ks >> 2) * 16384)) + (uint64_t)((ks & 3) * 2);
        const uint64_t bdesc = make_desc_sw128(B0 + (uint32_t)((kk >> 6) * cout * 128)) + (uint64_t)(((kk & 63) >> 3));
        dr_umma_ss(d_tmem, adesc, bdesc, idesc, (ks > 0 || acc_first) ? 1u : 0u);
    }
}

template <int C>
__device__ __forceinline__ void it_tmem_ld(uint32_t taddr, float* v) {
    if (C == 16) tmem_ld16(taddr, v);
    else tmem_ld32(taddr, v);
}

__global__ void __launch_bounds__(IT_THREADS, 1)
impala_direct_kernel(const ImpalaP L, const __grid_constant__ ItMaps maps, const float* __restrict__ replicas, int64_t stride,
                     const float* __restrict__ theta, const float* __restrict__ bnbuf, const int64_t* __restrict__ idx,
                     const int8_t* __restrict__ sign, float sigma, const float* __restrict__ frame, const float* __restrict__ reward,
                     const uint8_t* __restrict__ done, const float* __restrict__ h_in, const float* __restrict__ c_in, int E,
                     float* __restrict__ probs, float* __restrict__ h_out, float* __restrict__ c_out, int n_members, int pair_mode,
                     long long* __restrict__ prof) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[IB_COUNT];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const bool worker = tid < IT_WORKERS;
    const uint32_t bar0 = smem_u32(&bars[0]);
#define IT_BAR(i) (bar0 + 8u * (uint32_t)(i))
    const uint32_t sraw = smem_u32(smem_raw);
    const uint32_t s0 = (sraw + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (s0 - sraw);
    float* par_s = reinterpret_cast<float*>(sm + IT_PAR);
    float *sN = par_s, *tN = par_s + 32, *bias = par_s + 64;
    __half* fcin = reinterpret_cast<__half*>(sm + IT_FCIN);
    float* band = reinterpret_cast<float*>(sm + IT_BAND);
    int stamp_i = 0;
    auto stamp = [&]() {
        if (prof != nullptr && blockIdx.x == 7 && tid == 0) prof[stamp_i] = clock64();
        ++stamp_i;
    };

    const int nmem = pair_mode ? 2 : 1;
    const int mb = blockIdx.x / E, env = blockIdx.x - mb * E;
    const int ms[2] = {mb, pair_mode ? mb + (n_members >> 1) : mb};
    const int inst[2] = {ms[0] * E + env, ms[1] * E + env};
    const int sgi[2] = {(int)sign[ms[0]], (int)sign[ms[1]]};
    const int64_t ids[2] = {idx[ms[0]], idx[ms[1]]};
    const float* rows[2] = {table_row_ptr(replicas, stride, ids[0]), table_row_ptr(replicas, stride, ids[1])};
    const bool shared_row = nmem == 2 && ids[0] == ids[1];
    const int nE = shared_row ? 1 : nmem;

    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(IT_BAR(IB_MMA)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(IT_BAR(IB_GO)));
        for (int s = 0; s < IT_NSLOT; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(IT_BAR(IB_FULL + s)));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(IT_BAR(IB_EMPTY + s)));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    const int q4 = warp & 3, wg = warp >> 2;                    // TMEM lane quarter of this warp, its group of four warps
    const uint32_t lane_sel = (uint32_t)(q4 * 32) << 16;
    uint32_t mph = 0;
    stamp();

    if (worker) {
        // issue the MMAs of one tile (warp 0), then every worker waits for their completion (the im2col tile is free again)
        auto run_mma = [&](uint32_t d_col, int cout, int kk0, int nk, bool acc_first) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            it_wsync();
            if (warp == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                it_mma(tmem + d_col, s0 + IT_A, s0 + IT_B, cout, kk0, nk, dr_idesc(cout, 0), acc_first);
                umma_commit_elect(IT_BAR(IB_MMA));
            }
            dr_wait(IT_BAR(IB_MMA), mph); mph ^= 1u;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        };
        auto zero_map = [&](uint8_t* xh) {
            for (int i = tid; i < IT_XH_BYTES / 16; i += IT_WORKERS) reinterpret_cast<uint4*>(xh)[i] = make_uint4(0, 0, 0, 0);
        };
        // a whole convolution of a W x W map with cin = C input channels into TMEM column block d_col (+ T * cout per tile)
        auto conv_block = [&](const uint8_t* xh, int lw, int cin, int cout, uint32_t d_col, bool accumulate) {
            const int M = 1 << (2 * lw), ntile = (M + 127) >> 7, lcpt = cin == 16 ? 1 : 2;
            for (int T = 0; T < ntile; ++T) {
                if (cin == 16) {
                    it_im2col(s0 + IT_A, xh, lw, 16, lcpt, T, M, 0, 18);
                    run_mma(d_col + (uint32_t)(T * cout), cout, 0, 9, accumulate);
                } else {
                    it_im2col(s0 + IT_A, xh, lw, 32, lcpt, T, M, 0, 24);
                    run_mma(d_col + (uint32_t)(T * cout), cout, 0, 12, accumulate);
                    it_im2col(s0 + IT_A, xh, lw, 32, lcpt, T, M, 24, 12);
                    run_mma(d_col + (uint32_t)(T * cout), cout, 192, 6, true);
                }
            }
        };

        for (int mem = 0; mem < nmem; ++mem) {
            ItCtx c;
            c.theta = theta; c.bn = bnbuf; c.sg = sigma * (float)sgi[mem]; c.row = rows[mem];
            uint8_t* xa = sm + IT_XHA;
            uint8_t* xb = sm + IT_XHB;
            it_wsync();
            // ---- frame / 255 -> BN of the first convolution -> fp16 [66 x 66][4] zero-bordered (impala.py:142) ----
            zero_map(xa);
            if (tid < 3) {
                const ConvP& p0 = L.feat[0];
                const float inv = 1.0f / sqrtf(bnbuf[p0.bv + tid] + 1e-5f);
                const float s = c.par(p0.g + tid) * inv;
                sN[tid] = s;
                tN[tid] = c.par(p0.be + tid) - bnbuf[p0.bm + tid] * s;
            }
            it_wsync();
            {
                const float* fr = frame + (int64_t)inst[mem] * 12288;
                for (int p = tid; p < 4096; p += IT_WORKERS) {
                    const float f0 = fr[p] / 255.0f, f1 = fr[4096 + p] / 255.0f, f2 = fr[8192 + p] / 255.0f;
                    const uint32_t lo = dr_pack(fmaf(f0, sN[0], tN[0]), fmaf(f1, sN[1], tN[1]));
                    const uint32_t hi = dr_pack(fmaf(f2, sN[2], tN[2]), 0.f);
                    *reinterpret_cast<uint2*>(xa + (((p >> 6) + 1) * 66 + (p & 63) + 1) * 8) = make_uint2(lo, hi);
                }
            }
            it_wsync();
            if (mem == 0) stamp();
            int li = 0;
            for (int s = 0; s < 3; ++s) {
                // =========== stage convolution (BN on the input, no ReLU) + max-pool 3x3 / 2 pad 1 ===========
                const ConvP& fp = L.feat[s];
                const int lwc = 6 - s, Wc = 1 << lwc, C = fp.cout, cinp = s == 0 ? 4 : fp.cin;
                const int lwo = lwc - 1, Wo = 1 << lwo;
                const ConvP& nx = L.res[0][s][0];
                it_prep_layer(c, fp, cinp, sm + IT_B, sN, tN, bias, &nx);
                zero_map(xb);                                   // becomes the pooled operand map (new geometry)
                it_wsync();
                const int tiles_band = (8 * Wc) >> 7;           // 4, 2, 1 tiles of 128 conv pixels per band of 8 rows
                const int npool = 4 * Wo;                       // pooled pixels per band: 128, 64, 32
                for (int b = 0; b < Wc / 8; ++b) {
                    for (int tl = 0; tl < tiles_band; ++tl) {
                        const int T = b * tiles_band + tl;
                        if (s == 0) {
                            it_im2col_first(s0 + IT_A, xa, T);
                            run_mma(TC_P + (uint32_t)(tl * 16), 16, 0, 3, false);
                        } else if (s == 1) {
                            it_im2col(s0 + IT_A, xa, lwc, 16, 1, T, Wc * Wc, 0, 18);
                            run_mma(TC_P + (uint32_t)(tl * 32), 32, 0, 9, false);
                        } else {
                            it_im2col(s0 + IT_A, xa, lwc, 32, 2, T, Wc * Wc, 0, 24);
                            run_mma(TC_P + (uint32_t)(tl * 32), 32, 0, 12, false);
                            it_im2col(s0 + IT_A, xa, lwc, 32, 2, T, Wc * Wc, 24, 12);
                            run_mma(TC_P + (uint32_t)(tl * 32), 32, 192, 6, true);
                        }
                    }
                    // band epilogue: conv rows 8b .. 8b + 7 (+ bias) -> circular band [row % 9][x][c] fp32
                    if (wg < tiles_band) {
                        const int p = (b * tiles_band + wg) * 128 + q4 * 32 + lane, y = p >> lwc, x = p & (Wc - 1);
                        float* dst = band + ((y % 9) * Wc + x) * C;
                        float v[32];
                        if (C == 16) it_tmem_ld<16>(tmem + lane_sel + TC_P + (uint32_t)(wg * 16), v);
                        else it_tmem_ld<32>(tmem + lane_sel + TC_P + (uint32_t)(wg * 32), v);
                        for (int ch = 0; ch < C; ch += 4)
                            *reinterpret_cast<float4*>(dst + ch) = make_float4(v[ch] + bias[ch], v[ch + 1] + bias[ch + 1], v[ch + 2] + bias[ch + 2], v[ch + 3] + bias[ch + 3]);
                        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    }
                    it_wsync();
                    // pool: pooled pixel P = b * npool + i lives in lane P % 128 of residual tile P / 128
                    const int l0 = (b * npool) & 127;
                    if (tid >= l0 && tid < l0 + npool) {
                        const int Pp = b * npool + (tid - l0), py = Pp >> lwo, px = Pp & (Wo - 1);
                        float m[32];
#pragma unroll
                        for (int ch = 0; ch < 32; ++ch) m[ch] = -INFINITY;
                        for (int dy = -1; dy <= 1; ++dy) {
                            const int yy = 2 * py + dy;
                            if (yy < 0 || yy >= Wc) continue;
                            for (int dx = -1; dx <= 1; ++dx) {
                                const int xx = 2 * px + dx;
                                if (xx < 0 || xx >= Wc) continue;
                                const float4* src = reinterpret_cast<const float4*>(band + ((yy % 9) * Wc + xx) * C);
#pragma unroll
                                for (int g = 0; g < 8; ++g) {
                                    if (g * 4 < C) {
                                        const float4 f = src[g];
                                        m[4 * g] = fmaxf(m[4 * g], f.x); m[4 * g + 1] = fmaxf(m[4 * g + 1], f.y);
                                        m[4 * g + 2] = fmaxf(m[4 * g + 2], f.z); m[4 * g + 3] = fmaxf(m[4 * g + 3], f.w);
                                    }
                                }
                            }
                        }
                        // raw pooled map -> TMEM residual stream; BN + ReLU of the first block convolution -> operand map
                        const uint32_t rt = tmem + lane_sel + TC_R + (uint32_t)((Pp >> 7) * C);
                        it_tmem_st16(rt, m);
                        if (C == 32) it_tmem_st16(rt + 16u, m + 16);
                        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                        uint32_t o[16];
#pragma unroll
                        for (int g = 0; g < 16; ++g)
                            o[g] = dr_pack(fmaxf(fmaf(m[2 * g], sN[2 * g], tN[2 * g]), 0.f), fmaxf(fmaf(m[2 * g + 1], sN[2 * g + 1], tN[2 * g + 1]), 0.f));
                        uint4* dst = reinterpret_cast<uint4*>(xb + (((py + 1) * (Wo + 2) + px + 1) * C) * 2);
                        dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
                        dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
                        if (C == 32) { dst[2] = make_uint4(o[8], o[9], o[10], o[11]); dst[3] = make_uint4(o[12], o[13], o[14], o[15]); }
                        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    }
                    it_wsync();
                }
                ++li;
                { uint8_t* t_ = xa; xa = xb; xb = t_; }            // xa: operand map of the first block convolution
                zero_map(xb);                                       // old-geometry map: becomes the block-internal operand map
                if (mem == 0) stamp();
                // =========== two residual blocks: x += conv_b(relu(BN_b(conv_a(relu(BN_a(x)))))) ===========
                const int lw = lwo, W = Wo, M = W * W, ntile = (M + 127) >> 7;
                for (int blk = 0; blk < 2; ++blk) {
                    const ConvP& pa = L.res[blk][s][0];
                    const ConvP& pb = L.res[blk][s][1];
                    // ---- conv a -> Y; epilogue: relu(BN_b(y + bias)) -> xb ----
                    it_prep_layer(c, pa, pa.cin, sm + IT_B, sN, tN, bias, &pb);
                    it_wsync();
                    conv_block(xa, lw, C, C, TC_Y, false);
                    for (int T = wg; T < ntile; T += 4) {
                        const int p = T * 128 + q4 * 32 + lane;
                        float v[32];
                        if (C == 16) it_tmem_ld<16>(tmem + lane_sel + TC_Y + (uint32_t)(T * 16), v);
                        else it_tmem_ld<32>(tmem + lane_sel + TC_Y + (uint32_t)(T * 32), v);
                        if (p < M) {
                            const int y = p >> lw, x = p & (W - 1);
                            uint4* dst = reinterpret_cast<uint4*>(xb + (((y + 1) * (W + 2) + x + 1) * C) * 2);
                            uint32_t o[16];
#pragma unroll
                            for (int g = 0; g < 16; ++g)
                                if (2 * g < C)
                                    o[g] = dr_pack(fmaxf(fmaf(v[2 * g] + bias[2 * g], sN[2 * g], tN[2 * g]), 0.f),
                                                   fmaxf(fmaf(v[2 * g + 1] + bias[2 * g + 1], sN[2 * g + 1], tN[2 * g + 1]), 0.f));
                            dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
                            dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
                            if (C == 32) { dst[2] = make_uint4(o[8], o[9], o[10], o[11]); dst[3] = make_uint4(o[12], o[13], o[14], o[15]); }
                        }
                        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    }
                    ++li;
                    it_wsync();
                    // ---- conv b accumulates ONTO the residual stream in TMEM; epilogue: x += bias, next operand map -> xa ----
                    const bool last = (s == 2 && blk == 1);
                    const ConvP* nxt = last ? nullptr : (blk == 0 ? &L.res[1][s][0] : &L.feat[s + 1]);
                    const bool relu_next = blk == 0;                 // the next consumer is a block convolution (BN + ReLU) or a stage convolution (BN only)
                    it_prep_layer(c, pb, pb.cin, sm + IT_B, sN, tN, bias, nxt);
                    it_wsync();
                    conv_block(xb, lw, C, C, TC_R, true);
                    for (int T = wg; T < ntile; T += 4) {
                        const int p = T * 128 + q4 * 32 + lane;
                        float v[32];
                        const uint32_t rt = tmem + lane_sel + TC_R + (uint32_t)(T * C);
                        if (C == 16) it_tmem_ld<16>(rt, v);
                        else it_tmem_ld<32>(rt, v);
#pragma unroll
                        for (int ch = 0; ch < 32; ++ch)
                            if (ch < C) v[ch] += bias[ch];
                        if (!last) {
                            it_tmem_st16(rt, v);
                            if (C == 32) it_tmem_st16(rt + 16u, v + 16);
                            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                        }
                        if (p < M) {
                            if (!last) {
                                const int y = p >> lw, x = p & (W - 1);
                                uint4* dst = reinterpret_cast<uint4*>(xa + (((y + 1) * (W + 2) + x + 1) * C) * 2);
                                uint32_t o[16];
#pragma unroll
                                for (int g = 0; g < 16; ++g)
                                    if (2 * g < C) {
                                        float u0 = fmaf(v[2 * g], sN[2 * g], tN[2 * g]), u1 = fmaf(v[2 * g + 1], sN[2 * g + 1], tN[2 * g + 1]);
                                        if (relu_next) { u0 = fmaxf(u0, 0.f); u1 = fmaxf(u1, 0.f); }
                                        o[g] = dr_pack(u0, u1);
                                    }
                                dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
                                dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
                                if (C == 32) { dst[2] = make_uint4(o[8], o[9], o[10], o[11]); dst[3] = make_uint4(o[12], o[13], o[14], o[15]); }
                            } else {
                                // relu -> flatten (C, H, W) -> BN1d(2048), exactly perturbed gamma / beta (impala.py:153-155)
                                for (int ch = 0; ch < 32; ++ch) {
                                    const int k = ch * 64 + p;
                                    const float inv = 1.0f / sqrtf(bnbuf[L.fc_bv + k] + 1e-5f);
                                    const float sc = c.par(L.fc_g + k) * inv;
                                    fcin[mem * 2048 + k] = __float2half_rn(fmaf(fmaxf(v[ch], 0.f), sc, c.par(L.fc_be + k) - bnbuf[L.fc_bm + k] * sc));
                                }
                            }
                        }
                        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    }
                    ++li;
                    it_wsync();
                    if (mem == 0) stamp();
                }
                // the next stage convolution reads xa (BN only); its pooled output goes to xb after a zero fill
            }
        }
        // ======================================= dense tail =======================================
        stamp_i = 11;
        stamp();
        uint8_t* xt = sm + IT_XT;
        uint8_t* ct = sm + IT_CT;
        uint8_t* ht = sm + IT_HT;
        float* st = reinterpret_cast<float*>(sm);                 // fp32 scratch once the ring is dead
        for (int i = tid; i < (33 * 1024) / 16; i += IT_WORKERS) reinterpret_cast<uint4*>(xt)[i] = make_uint4(0, 0, 0, 0);
        for (int i = tid; i < (10 * 1024) / 16; i += IT_WORKERS) reinterpret_cast<uint4*>(ct)[i] = make_uint4(0, 0, 0, 0);
        it_wsync();
        // B operands: column n = member n; element (n, k) of box k >> 6 at n * 128 + (((k & 63) >> 3) ^ n) * 16 + (k & 7) * 2
        for (int i = tid; i < nmem * 2048; i += IT_WORKERS) {
            const int n = i >> 11, k = i & 2047;
            *reinterpret_cast<__half*>(xt + (k >> 6) * 1024 + n * 128 + ((((k & 63) >> 3) ^ n) << 4) + (k & 7) * 2) = fcin[i];
        }
        for (int i = tid; i < nmem * 256; i += IT_WORKERS) {
            const int n = i >> 8, k = i & 255;
            const float hv = done[inst[n]] ? 0.f : h_in[(int64_t)inst[n] * 256 + k];
            *reinterpret_cast<__half*>(ht + (k >> 6) * 1024 + n * 128 + ((((k & 63) >> 3) ^ n) << 4) + (k & 7) * 2) = __float2half_rn(hv);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        it_wsync();
        if (tid == 0) dr_arrive(IT_BAR(IB_GO));                    // the ring region is free: the producer may start
        const uint32_t id16 = dr_idesc(16, 0);
        // ---- Linear 2048 -> 256: tiles (kb, part, mt) ----
        int g = 0;
        if (warp == 0) {
#pragma unroll 1
            for (int kb = 0; kb < 32; ++kb) {
                const uint64_t bdesc = make_desc_sw128(s0 + IT_XT + (uint32_t)kb * 1024u);
#pragma unroll 1
                for (int part = 0; part < 1 + nE; ++part)
#pragma unroll 1
                    for (int mt = 0; mt < 2; ++mt, ++g) {
                        const int slot = g % IT_NSLOT;
                        dr_wait(IT_BAR(IB_FULL + slot), (uint32_t)((g / IT_NSLOT) & 1));
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint64_t adesc = make_desc_sw128(s0 + (uint32_t)slot * 16384u);
                        const uint32_t d = tmem + (part == 0 ? TC_FW + (uint32_t)(mt * 16) : TC_FE + (uint32_t)(((part - 1) * 2 + mt) * 16));
#pragma unroll
                        for (int j = 0; j < 4; ++j) dr_umma_ss(d, adesc + (uint64_t)(j * 2), bdesc + (uint64_t)(j * 2), id16, (kb | j) ? 1u : 0u);
                        umma_commit_elect(IT_BAR(IB_EMPTY + slot));
                    }
            }
            umma_commit_elect(IT_BAR(IB_MMA));
        }
        dr_wait(IT_BAR(IB_MMA), mph); mph ^= 1u;
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
        it_wsync();
        // ---- LSTM gates: classes c (gate row of lane q = 8 q + c), sources weight_ih[:, :256] . core and weight_hh . h0 ----
        if (warp == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int cc = 0; cc < 8; ++cc)
#pragma unroll 1
                for (int src = 0; src < 2; ++src)
#pragma unroll 1
                    for (int kb = 0; kb < 4; ++kb) {
                        const uint64_t bdesc = make_desc_sw128(s0 + (src ? IT_HT : IT_CT) + (uint32_t)kb * 1024u);
#pragma unroll 1
                        for (int part = 0; part < 1 + nE; ++part, ++g) {
                            const int slot = g % IT_NSLOT;
                            dr_wait(IT_BAR(IB_FULL + slot), (uint32_t)((g / IT_NSLOT) & 1));
                            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                            const uint64_t adesc = make_desc_sw128(s0 + (uint32_t)slot * 16384u);
                            const uint32_t d = tmem + (part == 0 ? TC_GW + (uint32_t)(cc * 16) : TC_GE + (uint32_t)(((part - 1) * 8 + cc) * 16));
#pragma unroll
                            for (int j = 0; j < 4; ++j) dr_umma_ss(d, adesc + (uint64_t)(j * 2), bdesc + (uint64_t)(j * 2), id16, (src | kb | j) ? 1u : 0u);
                            umma_commit_elect(IT_BAR(IB_EMPTY + slot));
                        }
                    }
            umma_commit_elect(IT_BAR(IB_MMA));
        }
        dr_wait(IT_BAR(IB_MMA), mph); mph ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // gates epilogue: warp -> lane quarter q4, classes wg and wg + 4; + reward column, biases (exact fp32)
        for (int cc = wg; cc < 8; cc += 4) {
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
        it_wsync();
        // ---- LSTM cell, policy head (fp32, exactly perturbed parameters), softmax ----
        for (int n = 0; n < nmem; ++n) {
            const float sgn = sigma * (float)sgi[n];
            const float* rw = rows[n];
            const bool dn = done[inst[n]] != 0;
            const float* gates = st + n * ST + 516;
            float* hn = st + n * ST + 1540;
            for (int k = tid; k < 256; k += IT_WORKERS) {
                const float c0 = dn ? 0.f : c_in[(int64_t)inst[n] * 256 + k];
                const float ig = it_sigmoid(gates[k]), fg = it_sigmoid(gates[256 + k]);
                const float gg = tanhf(gates[512 + k]), og = it_sigmoid(gates[768 + k]);
                const float c1 = fg * c0 + ig * gg;
                const float h1 = og * tanhf(c1);
                c_out[(int64_t)inst[n] * 256 + k] = c1;
                h_out[(int64_t)inst[n] * 256 + k] = h1;
                const float inv = 1.0f / sqrtf(bnbuf[L.pol_bv + k] + 1e-5f);
                const float sc = perturb1(theta[L.pol_g + k], sgn, rw[L.pol_g + k]) * inv;
                hn[k] = fmaf(h1, sc, perturb1(theta[L.pol_be + k], sgn, rw[L.pol_be + k]) - bnbuf[L.pol_bm + k] * sc);
            }
        }
        it_wsync();
        for (int a = warp; a < nmem * L.A; a += IT_WORKERS / 32) {
            const int n = a / L.A, ai = a - n * L.A;
            const float sgn = sigma * (float)sgi[n];
            const float* rw = rows[n];
            const float* hn = st + n * ST + 1540;
            float acc = 0.f;
            for (int k = lane; k < 256; k += 32) acc = fmaf(perturb1(theta[L.pol_w + ai * 256 + k], sgn, rw[L.pol_w + ai * 256 + k]), hn[k], acc);
            acc = warp_sum(acc);
            if (lane == 0) st[n * ST + 1796 + ai] = acc + perturb1(theta[L.pol_b + ai], sgn, rw[L.pol_b + ai]);
        }
        it_wsync();
        if (tid < nmem) {
            const float* lg = st + tid * ST + 1796;
            float mx = -INFINITY;
            for (int a = 0; a < L.A; ++a) mx = fmaxf(mx, lg[a]);
            float ssum = 0.f;
            for (int a = 0; a < L.A; ++a) ssum += expf(lg[a] - mx);
            const float inv = 1.0f / ssum;
            for (int a = 0; a < L.A; ++a) probs[(int64_t)inst[tid] * L.A + a] = expf(lg[a] - mx) * inv;
        }
        stamp();
    } else if (lane == 0) {
        // =============================== TMA producer of the dense tail ===============================
        dr_wait(IT_BAR(IB_GO), 0);
        int g = 0;
        auto put = [&](const CUtensorMap* map, int rank, int c0, int c1, int c2, int c3) {
            const int slot = g % IT_NSLOT;
            dr_wait(IT_BAR(IB_EMPTY + slot), (uint32_t)((g / IT_NSLOT) & 1) ^ 1u);
            dr_expect_tx(IT_BAR(IB_FULL + slot), 16384u);
            const uint32_t dst = s0 + (uint32_t)slot * 16384u;
            if (rank == 4) dr_tma_4d(dst, map, c0, c1, c2, c3, IT_BAR(IB_FULL + slot));
            else if (rank == 2) dr_tma_2d(dst, map, c0, c1, IT_BAR(IB_FULL + slot));
            else asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
                              "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(IT_BAR(IB_FULL + slot)) : "memory");
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
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
#undef IT_BAR
}

// 3-D map over an fp16 [1024 x 256] block of the theta scratch: {k, q (rows 8 apart), class c}: box {64, 128, 1}
int it_map_w3(CUtensorMap* map, const __half* base) {
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
int it_map_e_class(CUtensorMap* map, dfd_ctx* ctx, int pitch) {
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

}  // namespace

// returns -1 when this path does not serve the call (no scaled mirror of this table for this sigma)
int dfd_impala_forward_direct_impl(dfd_ctx* ctx, const dfd_policy_desc* desc, const dfd_table* table, const float* theta,
                                   const float* bn_buffers, const int64_t* idx, const int8_t* sign, int n_members, float sigma,
                                   const float* frame, const float* reward, const uint8_t* done, const float* h_in,
                                   const float* c_in, int obs_per_member, float* probs, float* h_out, float* c_out,
                                   cudaStream_t st) {
    if (getenv("DFD_TC_NO_DIRECT")) return -1;
    if (!ctx->scaled16 || ctx->scaled_src != table->replicas || ctx->scaled_sigma != sigma) return -1;
    const ImpalaP L = make_impala(desc->n_act);
    if (ctx->theta16_cap < 1048576) return -1;
    ItMaps maps;
    __half* t16 = (__half*)ctx->theta16;
    int rc = 0;
    rc |= dr_map_w(&maps.w_fc, ctx, 0, 2048, 256, 128);
    rc |= dr_map_e(&maps.e_fc, ctx, 2048, 256, 128);
    rc |= it_map_w3(&maps.w_ih, t16 + 524288);
    rc |= it_map_w3(&maps.w_hh, t16 + 786432);
    rc |= it_map_e_class(&maps.e_ih, ctx, 257);
    rc |= it_map_e_class(&maps.e_hh, ctx, 256);
    DFD_CHECK_ARG(rc == 0, "IMPALA tensor path: cuTensorMapEncodeTiled failed");
    impala_theta16_kernel<<<ctx->sm_count * 2, 512, 0, st>>>(theta, t16, L.fc_w, L.wih, L.whh);
    DFD_LAUNCHED(ctx);
    DFD_CUDA(cudaFuncSetAttribute(impala_direct_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, IT_SMEM));
    const int pair_mode = n_members % 2 == 0 ? 1 : 0;
    const int grid = (pair_mode ? n_members / 2 : n_members) * obs_per_member;
    long long* prof = nullptr;
    if (getenv("DFD_IMPALA_PROF")) { cudaMalloc(&prof, 32 * 8); cudaMemset(prof, 0, 32 * 8); }
    impala_direct_kernel<<<grid, IT_THREADS, IT_SMEM, st>>>(L, maps, table->replicas, table->replica_stride, theta, bn_buffers, idx, sign,
                                                            sigma, frame, reward, done, h_in, c_in, obs_per_member, probs, h_out, c_out,
                                                            n_members, pair_mode, prof);
    DFD_LAUNCHED(ctx);
    if (prof) {
        cudaStreamSynchronize(st);
        long long h[32];
        cudaMemcpy(h, prof, sizeof(h), cudaMemcpyDeviceToHost);
        fprintf(stderr, "[impala tcgen05 timeline] CTA 7, cycles per phase of its first member: frame %lld | s0 conv+pool %lld res %lld %lld | "
                        "s1 conv+pool %lld res %lld %lld | s2 conv+pool %lld res %lld %lld | all trunks done at %lld | dense tail %lld | total %lld\n",
                h[1] - h[0], h[2] - h[1], h[3] - h[2], h[4] - h[3], h[5] - h[4], h[6] - h[5], h[7] - h[6], h[8] - h[7], h[9] - h[8],
                h[10] - h[9], h[11] - h[0], h[12] - h[11], h[12] - h[0]);
        cudaFree(prof);
    }
    return 0;
}
