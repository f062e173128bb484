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
#include "impala_tail.cuh"

namespace {

constexpr int IT_WORKERS = 384, IT_THREADS = IT_WORKERS + 32, IT_GROUPS = IT_WORKERS / 128;   // worker warps come in groups of four TMEM lane quarters
// shared memory map (bytes from the 1024-aligned base)
constexpr int IT_A = 0, IT_A_BYTES = 49152;                  // ring of 3 im2col buffers [128 pixels x 64 k] fp16
constexpr int IT_B = IT_A + IT_A_BYTES, IT_B_BYTES = 20480;   // weights of a layer: up to 5 boxes of [32 x 64]; two buffers
constexpr int IT_XH_BYTES = 38400;                            // one operand map (34*34*16*2 = 36 992 is the largest)
constexpr int IT_XHA = IT_B + 2 * IT_B_BYTES, IT_XHB = IT_XHA + IT_XH_BYTES;
constexpr int IT_BAND = IT_XHB + IT_XH_BYTES, IT_BAND_BYTES = 36864;   // 9 rows x 64 x 16 (or 32 x 32) fp32
constexpr int IT_FCIN = IT_BAND + IT_BAND_BYTES;              // fp16 [2][2048]: BN'd trunk outputs of the two members
constexpr int IT_PAR = IT_FCIN + 8192, IT_PAR_WORDS = 160;    // two buffers of: sN[32] tN[32] bias[32] offs[40]
constexpr int IT_SMEM = IT_PAR + 2 * IT_PAR_WORDS * 4 + 1024;
constexpr int IT_XT = IT_BAND;                                // dense tail: x tiles (impala_tail.cuh) in the band buffer
// tensor memory columns of the trunk
constexpr uint32_t TC_R = 0, TC_Y = 128, TC_P = 256;          // residual stream | block-internal map | stage-conv band
// barriers: the tail's (impala_tail.cuh; TB_MMA doubles as "layer / band done" in the trunk), then the im2col pipeline's
enum { IB_MMA = TB_MMA, IB_AFULL = TB_COUNT, IB_AEMPTY = IB_AFULL + 3, IB_COUNT = IB_AEMPTY + 3 };

__device__ __forceinline__ void it_wsync() { asm volatile("bar.sync 1, %0;" ::"n"(IT_WORKERS) : "memory"); }
__device__ __forceinline__ void it_tmem_st16(uint32_t taddr, const float* v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
        "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]), "f"(v[8]), "f"(v[9]), "f"(v[10]),
        "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15])
        : "memory");
}
struct ItCtx {
    const float* theta;
    const float* row;
    const float* bn;
    float sg;
    __device__ __forceinline__ float par(int p) const { return perturb1(theta[p], sg, row[p]); }
};

// Everything a layer needs that is not its input map: weights -> B operand [cout rows x K] fp16, K-major 128-byte swizzle,
// k = tap * cinp + c (one 64-wide box per im2col pass); conv bias; byte offsets of the 16-byte chunks of a pixel's 3x3 window
// in the zero-bordered channel-last input map (Wp = padded width); input-side BN of the NEXT convolution folded to scale /
// shift (applied by THIS layer's epilogue).  par: sN[32] tN[32] bias[32] offs[40].
__device__ __noinline__ void it_prep_layer(const ItCtx& c, const ConvP& p, int cinp, int Wp, uint8_t* Bs, float* par, const ConvP* nxt) {
    const int tid = threadIdx.x;
    const int k9 = p.cin * 9, n = p.cout * k9;
    float *sN = par, *tN = par + 32, *bias = par + 64;
    int* offs = reinterpret_cast<int*>(par + 96);
    if (cinp != p.cin) {       // first layer: padded channel and K tail are zeros
        for (int i = tid; i < 16 * 128 / 16; i += IT_WORKERS) reinterpret_cast<uint4*>(Bs)[i] = make_uint4(0, 0, 0, 0);
        it_wsync();
    }
    // compact code on purpose (the kernel's instruction footprint is what a layer pays for first): 6 weights per round
#pragma unroll 1
    for (int t0 = tid; t0 < n; t0 += 6 * IT_WORKERS) {
        float a[6], e[6];
#pragma unroll
        for (int u = 0; u < 6; ++u) {
            const int t = min(t0 + u * IT_WORKERS, n - 1);
            a[u] = c.theta[p.w + t]; e[u] = c.row[p.w + t];
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
    else if (nxt != nullptr && tid >= 32 && tid < 32 + nxt->cin) {
        const int ch = tid - 32;
        const float inv = 1.0f / sqrtf(c.bn[nxt->bv + ch] + 1e-5f);
        const float s = c.par(nxt->g + ch) * inv;
        sN[ch] = s;
        tN[ch] = c.par(nxt->be + ch) - c.bn[nxt->bm + ch] * s;
    } else if (tid >= 64 && tid < 64 + 40) {
        const int q = tid - 64;
        if (cinp == 4) {                 // first layer: offs[tap] of the 8-byte pixel
            const int dy = (q * 11) >> 5, dx = q - 3 * dy;
            offs[q] = (dy * Wp + dx) * 8;
        } else {
            const int cpt = cinp >> 3, tap = q / cpt, part = q - tap * cpt, dy = (tap * 11) >> 5, dx = tap - 3 * dy;
            offs[q] = ((dy * Wp + dx) * cinp + part * 8) * 2;
        }
    }
}

// one im2col unit: output pixels T*128 + r of a W x W map, chunks [q0, q0 + nq) (nq <= 8: one 64-wide K box) of the 3x3
// window: a 16-byte copy per (pixel, chunk) from the channel-last map into the 128-byte swizzled K-major tile
__device__ __forceinline__ void it_build(uint32_t Ab, const uint8_t* xh, const int* offs, int lw, int cin, int T, int M, int q0, int nq) {
    const int t = threadIdx.x, r = t & 127, j0 = t >> 7, p = T * 128 + r;
    if (p < M) {
        const int W = 1 << lw, y = p >> lw, x = p & (W - 1);
        const uint8_t* pb = xh + ((y * (W + 2) + x) * cin) * 2;
        const uint32_t db = Ab + (uint32_t)(r * 128);
#pragma unroll
        for (int h = 0; h < 3; ++h) {
            const int j = j0 + IT_GROUPS * h;
            if (j < nq) {
                const uint4 v = *reinterpret_cast<const uint4*>(pb + offs[q0 + j]);
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(db + (uint32_t)((j ^ (r & 7)) << 4)), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
            }
        }
    }
}
// the first convolution: 4 (3 + 1 zero) halves per pixel, K = 36 padded to 48: chunk q = taps 2q, 2q + 1 (taps >= 9 are zeros)
__device__ __forceinline__ void it_build_first(uint32_t Ab, const uint8_t* xh, const int* offs, int T) {
    const int t = threadIdx.x, r = t & 127, j0 = t >> 7, p = T * 128 + r, y = p >> 6, x = p & 63;
    const uint8_t* pb = xh + (y * 66 + x) * 8;
    const uint32_t db = Ab + (uint32_t)(r * 128);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int j = j0 + IT_GROUPS * h;
        if (j < 6) {
            uint2 v0 = make_uint2(0, 0), v1 = make_uint2(0, 0);
            if (2 * j < 9) v0 = *reinterpret_cast<const uint2*>(pb + offs[2 * j]);
            if (2 * j + 1 < 9) v1 = *reinterpret_cast<const uint2*>(pb + offs[2 * j + 1]);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(db + (uint32_t)((j ^ (r & 7)) << 4)), "r"(v0.x), "r"(v0.y), "r"(v1.x), "r"(v1.y) : "memory");
        }
    }
}

template <int C>
__device__ __forceinline__ void it_tmem_ld(uint32_t taddr, float* v) {
    if (C == 16) tmem_ld16(taddr, v);
    else tmem_ld32(taddr, v);
}

// passes (one 64-wide K box each) of a layer with `cin` (padded) input channels and the k-steps of pass ps
__device__ __forceinline__ int it_npass(int cin) { return cin == 4 ? 1 : (cin == 16 ? 3 : 5); }
__device__ __forceinline__ int it_nchunk(int cin, int ps) { return cin == 4 ? 6 : (cin == 16 ? (ps < 2 ? 8 : 2) : (ps < 4 ? 8 : 4)); }

// worker side of the im2col pipeline (3 buffers): wait until the MMAs that read the buffer have completed (one poller per
// warp), copy, make the copies visible to the tensor core's proxy, publish.  `ug` = running unit number, the same
// sequence in the workers and in the MMA warp.
__device__ __noinline__ int it_build_tile(uint32_t bar0, uint32_t a0, int ug, const uint8_t* xh, const int* offs, int lw, int cin, int T, int M, long long* prof) {
#define IT_US(slot) do { if (prof != nullptr && blockIdx.x == 7 && threadIdx.x == 0 && ug >= 60 && ug < 62) prof[32 + (ug - 60) * 8 + (slot)] = clock64(); } while (0)
    const int lane = threadIdx.x & 31;
    const int np = it_npass(cin);
#pragma unroll 1
    for (int ps = 0; ps < np; ++ps) {
        const int bufi = ug % 3;
        IT_US(0);
        if (lane == 0) dr_wait(bar0 + 8u * (uint32_t)(IB_AEMPTY + bufi), (uint32_t)((ug / 3) & 1) ^ 1u);
        __syncwarp();
        IT_US(1);
        const uint32_t Ab = a0 + (uint32_t)bufi * 16384u;
        if (cin == 4) it_build_first(Ab, xh, offs, T);
        else it_build(Ab, xh, offs, lw, cin, T, M, ps * 8, it_nchunk(cin, ps));
        IT_US(2);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        IT_US(3);
        if (lane == 0) dr_arrive(bar0 + 8u * (uint32_t)(IB_AFULL + bufi));
        ++ug;
    }
    return ug;
#undef IT_US
}
__device__ __forceinline__ void it_layer_wait(uint32_t bar0, uint32_t& mph) {
    if ((threadIdx.x & 31) == 0) dr_wait(bar0 + 8u * (uint32_t)IB_MMA, mph);
    __syncwarp();
    mph ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void it_zero_map(uint8_t* xh) {
    for (int i = threadIdx.x; i < IT_XH_BYTES / 16; i += IT_WORKERS) reinterpret_cast<uint4*>(xh)[i] = make_uint4(0, 0, 0, 0);
}

// MMA-warp side of the pipeline: the passes of one output tile; frees each buffer with a commit
__device__ __noinline__ int it_mma_tile(uint32_t bar0, uint32_t s0, uint32_t tmem, int ug, int li, uint32_t d_col, int cin, int cout, bool accumulate, long long* prof) {
#define IT_MS(slot) do { if (prof != nullptr && blockIdx.x == 7 && (threadIdx.x & 31) == 0 && ug >= 60 && ug < 62) prof[32 + (ug - 60) * 8 + (slot)] = clock64(); } while (0)
    const int np = it_npass(cin);
    const uint32_t idesc = dr_idesc(cout, 0);
    const uint32_t Bb = s0 + IT_B + (uint32_t)((li & 1) * IT_B_BYTES);
#pragma unroll 1
    for (int ps = 0; ps < np; ++ps) {
        const int bufi = ug % 3;
        IT_MS(4);
        dr_wait(bar0 + 8u * (uint32_t)(IB_AFULL + bufi), (uint32_t)((ug / 3) & 1));
        IT_MS(5);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int nk = cin == 4 ? 3 : it_nchunk(cin, ps) >> 1;
        const uint64_t adesc = make_desc_sw128(s0 + IT_A + (uint32_t)bufi * 16384u);
        const uint64_t bdesc = make_desc_sw128(Bb + (uint32_t)(ps * cout * 128));
#pragma unroll 1
        for (int ks = 0; ks < nk; ++ks)
            dr_umma_ss(tmem + d_col, adesc + (uint64_t)(ks * 2), bdesc + (uint64_t)(ks * 2), idesc, ((ps | ks) != 0 || accumulate) ? 1u : 0u);
        umma_commit_elect(bar0 + 8u * (uint32_t)(IB_AEMPTY + bufi));
        IT_MS(6);
        ++ug;
    }
    return ug;
#undef IT_MS
}

__global__ void __launch_bounds__(IT_THREADS, 1)
impala_direct_kernel(const __grid_constant__ ImpalaP L, const __grid_constant__ ItMaps maps, const float* __restrict__ replicas, int64_t stride,
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
    TailArgs targs;
    targs.theta = theta; targs.bnbuf = bnbuf; targs.reward = reward; targs.done = done; targs.h_in = h_in; targs.c_in = c_in;
    targs.probs = probs; targs.h_out = h_out; targs.c_out = c_out; targs.sigma = sigma; targs.nmem = nmem; targs.nE = nE;
    for (int i = 0; i < 2; ++i) { targs.inst[i] = inst[i]; targs.sgi[i] = sgi[i]; targs.ids[i] = ids[i]; targs.rows[i] = rows[i]; }

    if (tid == 0) {
        for (int s = 0; s < TB_COUNT; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(IT_BAR(s)));
        for (int s = 0; s < 3; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(IT_BAR(IB_AFULL + s)), "r"(IT_WORKERS / 32));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(IT_BAR(IB_AEMPTY + s)));
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
    int ug = 0, li = 0;                  // running im2col unit / layer numbers: identical sequences in the workers and the MMA warp
    stamp();

    if (worker) {
        auto Bbuf = [&](int l) { return sm + IT_B + (l & 1) * IT_B_BYTES; };
        auto Pbuf = [&](int l) { return par_s + (l & 1) * IT_PAR_WORDS; };

#pragma unroll 1
        for (int mem = 0; mem < nmem; ++mem) {
            ItCtx c;
            c.theta = theta; c.bn = bnbuf; c.sg = sigma * (float)sgi[mem]; c.row = rows[mem];
            uint8_t* xa = sm + IT_XHA;
            uint8_t* xb = sm + IT_XHB;
            it_wsync();
            // ---- frame / 255 -> BN of the first convolution -> fp16 [66 x 66][4] zero-bordered (impala.py:142) ----
            float* fpar = reinterpret_cast<float*>(band);        // scale / shift of the three input channels
            it_zero_map(xa);
            if (tid < 3) {
                const ConvP& p0 = L.feat[0];
                const float inv = 1.0f / sqrtf(bnbuf[p0.bv + tid] + 1e-5f);
                const float s = c.par(p0.g + tid) * inv;
                fpar[tid] = s;
                fpar[4 + tid] = c.par(p0.be + tid) - bnbuf[p0.bm + tid] * s;
            }
            it_prep_layer(c, L.feat[0], 4, 66, Bbuf(li), Pbuf(li), &L.res[0][0][0]);
            it_wsync();
            {
                const float* fr = frame + (int64_t)inst[mem] * 12288;
                constexpr int NF = (4096 + IT_WORKERS - 1) / IT_WORKERS;
                float f[3 * NF];
#pragma unroll
                for (int u = 0; u < NF; ++u) {
                    const int p = min(tid + u * IT_WORKERS, 4095);
                    f[3 * u] = fr[p]; f[3 * u + 1] = fr[4096 + p]; f[3 * u + 2] = fr[8192 + p];
                }
#pragma unroll
                for (int u = 0; u < NF; ++u) {
                    const int p = min(tid + u * IT_WORKERS, 4095);
                    const uint32_t lo = dr_pack(fmaf(f[3 * u] / 255.0f, fpar[0], fpar[4]), fmaf(f[3 * u + 1] / 255.0f, fpar[1], fpar[5]));
                    const uint32_t hi = dr_pack(fmaf(f[3 * u + 2] / 255.0f, fpar[2], fpar[6]), 0.f);
                    *reinterpret_cast<uint2*>(xa + (((p >> 6) + 1) * 66 + (p & 63) + 1) * 8) = make_uint2(lo, hi);
                }
            }
            it_wsync();
            if (mem == 0) stamp();
#pragma unroll 1
            for (int s = 0; s < 3; ++s) {
                // =========== stage convolution (BN on the input, no ReLU) + max-pool 3x3 / 2 pad 1 ===========
                const ConvP& fp = L.feat[s];
                const int lwc = 6 - s, Wc = 1 << lwc, C = fp.cout, cinp = s == 0 ? 4 : fp.cin;
                const int lwo = lwc - 1, Wo = 1 << lwo;
                const float* par = Pbuf(li);
                const float *sN = par, *tN = par + 32, *bias = par + 64;
                const int* offs = reinterpret_cast<const int*>(par + 96);
                it_zero_map(xb);                                   // becomes the pooled operand map (new geometry)
                const int tiles_band = (8 * Wc) >> 7;           // 4, 2, 1 tiles of 128 conv pixels per band of 8 rows
                const int npool = 4 * Wo;                       // pooled pixels per band: 128, 64, 32
                const ConvP& pa0 = L.res[0][s][0];
#pragma unroll 1
                for (int b = 0; b < Wc / 8; ++b) {
                    if (prof != nullptr && blockIdx.x == 7 && tid == 0 && mem == 0 && s == 0 && b == 3) prof[48] = clock64();
#pragma unroll 1
                    for (int tl = 0; tl < tiles_band; ++tl) ug = it_build_tile(bar0, s0 + IT_A, ug, xa, offs, lwc, cinp, b * tiles_band + tl, Wc * Wc, prof);
                    auto bst = [&](int i) { if (prof != nullptr && blockIdx.x == 7 && tid == 0 && mem == 0 && s == 0 && b == 3) prof[48 + i] = clock64(); };
                    bst(1);
                    if (b == 0)                                 // the first block convolution's weights, under this layer's MMAs
                        it_prep_layer(c, pa0, pa0.cin, Wo + 2, Bbuf(li + 1), Pbuf(li + 1), &L.res[0][s][1]);
                    it_layer_wait(bar0, mph);
                    bst(2);
                    // band epilogue: conv rows 8b .. 8b + 7 (+ bias) -> circular band [row % 9][x][c] fp32
#pragma unroll 1
                    for (int tg = wg; tg < tiles_band; tg += IT_GROUPS) {
                        const int p = (b * tiles_band + tg) * 128 + q4 * 32 + lane, y = p >> lwc, x = p & (Wc - 1);
                        float* dst = band + ((y % 9) * Wc + x) * C;
                        float v[32];
                        if (C == 16) it_tmem_ld<16>(tmem + lane_sel + TC_P + (uint32_t)(tg * 16), v);
                        else it_tmem_ld<32>(tmem + lane_sel + TC_P + (uint32_t)(tg * 32), v);
#pragma unroll
                        for (int ch = 0; ch < 32; ch += 4)
                            if (ch < C)
                                *reinterpret_cast<float4*>(dst + ch) = make_float4(v[ch] + bias[ch], v[ch + 1] + bias[ch + 1], v[ch + 2] + bias[ch + 2], v[ch + 3] + bias[ch + 3]);
                        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    }
                    bst(3);
                    it_wsync();
                    bst(4);
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
                    bst(5);
                    it_wsync();
                    bst(6);
                }
                ++li;
                { uint8_t* t_ = xa; xa = xb; xb = t_; }            // xa: operand map of the first block convolution
                it_zero_map(xb);                                       // old-geometry map: becomes the block-internal operand map
                it_wsync();
                if (mem == 0) stamp();
                // =========== two residual blocks: x += conv_b(relu(BN_b(conv_a(relu(BN_a(x)))))) ===========
                const int lw = lwo, W = Wo, M = W * W, ntile = (M + 127) >> 7;
#pragma unroll 1
                for (int blk = 0; blk < 2; ++blk) {
                    const ConvP& pb = L.res[blk][s][1];
                    const bool last = (s == 2 && blk == 1);
                    const ConvP* nxt = last ? nullptr : (blk == 0 ? &L.res[1][s][0] : &L.feat[s + 1]);
                    // ---- conv a -> Y; epilogue: relu(BN_b(y + bias)) -> xb ----
                    {
                        const float* pr = Pbuf(li);
                        const float *sA = pr, *tA = pr + 32, *bA = pr + 64;
                        const int* offs_a = reinterpret_cast<const int*>(pr + 96);
                        auto fine = [&](int i) { if (prof != nullptr && blockIdx.x == 7 && tid == 0 && mem == 0 && blk == 0) prof[16 + 4 * s + i] = clock64(); };
                        fine(0);
#pragma unroll 1
                        for (int T = 0; T < ntile; ++T) ug = it_build_tile(bar0, s0 + IT_A, ug, xa, offs_a, lw, C, T, M, prof);
                        fine(1);
                        it_prep_layer(c, pb, pb.cin, W + 2, Bbuf(li + 1), Pbuf(li + 1), nxt);     // conv b's weights, under conv a's MMAs
                        fine(2);
                        it_layer_wait(bar0, mph);
                        fine(3);
#pragma unroll 1
                        for (int T = wg; T < ntile; T += IT_GROUPS) {
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
                                        o[g] = dr_pack(fmaxf(fmaf(v[2 * g] + bA[2 * g], sA[2 * g], tA[2 * g]), 0.f),
                                                       fmaxf(fmaf(v[2 * g + 1] + bA[2 * g + 1], sA[2 * g + 1], tA[2 * g + 1]), 0.f));
                                dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
                                dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
                                if (C == 32) { dst[2] = make_uint4(o[8], o[9], o[10], o[11]); dst[3] = make_uint4(o[12], o[13], o[14], o[15]); }
                            }
                            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                        }
                        ++li;
                        it_wsync();
                    }
                    // ---- conv b accumulates ONTO the residual stream in TMEM; epilogue: x += bias, next operand map -> xa ----
                    {
                        const float* pr = Pbuf(li);
                        const float *sB = pr, *tB = pr + 32, *bB = pr + 64;
                        const int* offs_b = reinterpret_cast<const int*>(pr + 96);
                        const bool relu_next = blk == 0;             // next consumer: a block convolution (BN + ReLU) or a stage convolution (BN only)
#pragma unroll 1
                        for (int T = 0; T < ntile; ++T) ug = it_build_tile(bar0, s0 + IT_A, ug, xb, offs_b, lw, C, T, M, prof);
                        if (nxt != nullptr) {                        // the next layer's weights, under conv b's MMAs
                            const bool nxt_stage = blk == 1;
                            it_prep_layer(c, *nxt, nxt->cin, W + 2, Bbuf(li + 1), Pbuf(li + 1),
                                          nxt_stage ? &L.res[0][s + 1][0] : &L.res[1][s][1]);
                        }
                        it_layer_wait(bar0, mph);
                        float* fscr = band;                          // relu(x) of the last layer, [c * 64 + pixel]
#pragma unroll 1
                        for (int T = wg; T < ntile; T += IT_GROUPS) {
                            const int p = T * 128 + q4 * 32 + lane;
                            float v[32];
                            const uint32_t rt = tmem + lane_sel + TC_R + (uint32_t)(T * C);
                            if (C == 16) it_tmem_ld<16>(rt, v);
                            else it_tmem_ld<32>(rt, v);
#pragma unroll
                            for (int ch = 0; ch < 32; ++ch)
                                if (ch < C) v[ch] += bB[ch];
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
                                            float u0 = fmaf(v[2 * g], sB[2 * g], tB[2 * g]), u1 = fmaf(v[2 * g + 1], sB[2 * g + 1], tB[2 * g + 1]);
                                            if (relu_next) { u0 = fmaxf(u0, 0.f); u1 = fmaxf(u1, 0.f); }
                                            o[g] = dr_pack(u0, u1);
                                        }
                                    dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
                                    dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
                                    if (C == 32) { dst[2] = make_uint4(o[8], o[9], o[10], o[11]); dst[3] = make_uint4(o[12], o[13], o[14], o[15]); }
                                } else {
#pragma unroll
                                    for (int ch = 0; ch < 32; ++ch) fscr[ch * 64 + p] = fmaxf(v[ch], 0.f);
                                }
                            }
                            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                        }
                        ++li;
                        it_wsync();
                        if (last) {
                            // relu -> flatten (C, H, W) -> BN1d(2048), exactly perturbed gamma / beta (impala.py:153-155)
                            constexpr int NK = (2048 + IT_WORKERS - 1) / IT_WORKERS;
                            float gt[NK], ge[NK], bt[NK], be[NK], vm[NK], vv[NK];
#pragma unroll
                            for (int u = 0; u < NK; ++u) {
                                const int k = min(tid + u * IT_WORKERS, 2047);
                                gt[u] = theta[L.fc_g + k]; ge[u] = c.row[L.fc_g + k];
                                bt[u] = theta[L.fc_be + k]; be[u] = c.row[L.fc_be + k];
                                vm[u] = bnbuf[L.fc_bm + k]; vv[u] = bnbuf[L.fc_bv + k];
                            }
#pragma unroll
                            for (int u = 0; u < NK; ++u) {
                                const int k = min(tid + u * IT_WORKERS, 2047);
                                const float inv = 1.0f / sqrtf(vv[u] + 1e-5f);
                                const float sc = perturb1(gt[u], c.sg, ge[u]) * inv;
                                fcin[mem * 2048 + k] = __float2half_rn(fmaf(fscr[k], sc, perturb1(bt[u], c.sg, be[u]) - vm[u] * sc));
                            }
                            it_wsync();
                        }
                    }
                    if (mem == 0) stamp();
                }
            }
        }
        // ======================================= dense tail =======================================
        stamp_i = 11;
        stamp();
        tl_dense_tail_workers<IT_WORKERS, IT_XT>(L, targs, sm, s0, fcin, bar0, tmem, mph);
        stamp();
    } else {
        // =============================== MMA issue warp of the trunk ===============================
        // mirrors the workers' sequence of im2col units: waits for a buffer, issues its k-steps, frees it with a commit;
        // a second commit at the end of every layer (stage convolutions: every band) tells the workers the accumulator is ready
#pragma unroll 1
        for (int mem = 0; mem < nmem; ++mem) {
#pragma unroll 1
            for (int s = 0; s < 3; ++s) {
                const int Wc = 64 >> s, C = s == 0 ? 16 : 32, cinp = s == 0 ? 4 : (s == 1 ? 16 : 32);
                const int tiles_band = (8 * Wc) >> 7;
#pragma unroll 1
                for (int b = 0; b < Wc / 8; ++b) {
#pragma unroll 1
                    for (int tl = 0; tl < tiles_band; ++tl) ug = it_mma_tile(bar0, s0, tmem, ug, li, TC_P + (uint32_t)(tl * C), cinp, C, false, prof);
                    umma_commit_elect(IT_BAR(IB_MMA));
                }
                ++li;
                const int W = Wc >> 1, M = W * W, ntile = (M + 127) >> 7;
#pragma unroll 1
                for (int blk = 0; blk < 2; ++blk) {
#pragma unroll 1
                    for (int T = 0; T < ntile; ++T) ug = it_mma_tile(bar0, s0, tmem, ug, li, TC_Y + (uint32_t)(T * C), C, C, false, prof);
                    umma_commit_elect(IT_BAR(IB_MMA));
                    ++li;
#pragma unroll 1
                    for (int T = 0; T < ntile; ++T) ug = it_mma_tile(bar0, s0, tmem, ug, li, TC_R + (uint32_t)(T * C), C, C, true, prof);
                    umma_commit_elect(IT_BAR(IB_MMA));
                    ++li;
                }
            }
        }
        __syncwarp();
        if (lane == 0) {
            tl_dense_tail_producer(L, maps, targs, s0, bar0);
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
    if (tl_prepare(ctx, L, theta, &maps, st)) return 3;
    DFD_CUDA(cudaFuncSetAttribute(impala_direct_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, IT_SMEM));
    const int pair_mode = n_members % 2 == 0 ? 1 : 0;
    const int grid = (pair_mode ? n_members / 2 : n_members) * obs_per_member;
    long long* prof = nullptr;
    if (getenv("DFD_IMPALA_PROF")) { cudaMalloc(&prof, 64 * 8); cudaMemset(prof, 0, 64 * 8); }
    impala_direct_kernel<<<grid, IT_THREADS, IT_SMEM, st>>>(L, maps, table->replicas, table->replica_stride, theta, bn_buffers, idx, sign,
                                                            sigma, frame, reward, done, h_in, c_in, obs_per_member, probs, h_out, c_out,
                                                            n_members, pair_mode, prof);
    DFD_LAUNCHED(ctx);
    if (prof) {
        cudaStreamSynchronize(st);
        long long h[64];
        cudaMemcpy(h, prof, sizeof(h), cudaMemcpyDeviceToHost);
        for (int u = 0; u < 2; ++u)
            fprintf(stderr, "[impala tcgen05 timeline] unit %d: worker wait-empty %lld copy %lld fence %lld | mma: wait-full from %lld to %lld (rel. worker start), issued+commit %lld\n", 60 + u,
                    h[33 + 8 * u] - h[32 + 8 * u], h[34 + 8 * u] - h[33 + 8 * u], h[35 + 8 * u] - h[34 + 8 * u], h[36 + 8 * u] - h[32 + 8 * u], h[37 + 8 * u] - h[32 + 8 * u], h[38 + 8 * u] - h[37 + 8 * u]);
        fprintf(stderr, "[impala tcgen05 timeline] stage 0 band 3: builds %lld | wait MMA %lld | band epilogue %lld | sync %lld | pool %lld | sync %lld\n",
                h[49] - h[48], h[50] - h[49], h[51] - h[50], h[52] - h[51], h[53] - h[52], h[54] - h[53]);
        for (int s = 0; s < 3; ++s)
            fprintf(stderr, "[impala tcgen05 timeline] stage %d first block conv a: builds %lld | prep next %lld | wait MMA %lld | (start at %lld after stage conv)\n", s,
                    h[17 + 4 * s] - h[16 + 4 * s], h[18 + 4 * s] - h[17 + 4 * s], h[19 + 4 * s] - h[18 + 4 * s], h[16 + 4 * s] - h[2 + 3 * s]);
        fprintf(stderr, "[impala tcgen05 timeline] CTA 7, cycles per phase of its first member: frame %lld | s0 conv+pool %lld res %lld %lld | "
                        "s1 conv+pool %lld res %lld %lld | s2 conv+pool %lld res %lld %lld | all trunks done at %lld | dense tail %lld | total %lld\n",
                h[1] - h[0], h[2] - h[1], h[3] - h[2], h[4] - h[3], h[5] - h[4], h[6] - h[5], h[7] - h[6], h[8] - h[7], h[9] - h[8],
                h[10] - h[9], h[11] - h[0], h[12] - h[11], h[12] - h[0]);
        cudaFree(prof);
    }
    return 0;
}
