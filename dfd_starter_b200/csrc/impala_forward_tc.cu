// IMPALA CNN + LSTM perturbed forward on tcgen05 / TMEM / TMA (policies/impala.py:136-186; perturbation worker/worker.py:28),
// precision level 3.  One CTA per (antithetic pair, environment): the two members' trunks run one after the other, the dense
// tail of both streams theta and the pair's shared eps row ONCE.
//
// Trunk (15 convolutions 3x3, pad 1): implicit GEMMs M = 128 output positions, N = output channels, K = 16 input channels
// of ONE filter tap per MMA - and NO im2col.  The operand of a convolution - BN(+ReLU) of the previous raw map - is written
// ONCE by the producing layer's epilogue as fp16 PLANES: plane j = channels 8j .. 8j + 7 of every position of the
// zero-bordered (W + 2) x (W + 2) map, 16 bytes per position, positions row-major.  That IS the UMMA no-swizzle K-major
// layout with the position as the row (8 consecutive positions = one 128-byte core matrix, SBO = 128, LBO = one plane),
// so a filter tap (dy, dx) is the same descriptor with its start address moved by (dy (W + 2) + dx) positions, and an
// output tile is 128 consecutive PADDED positions (the two border columns of a row give garbage rows that nobody reads).
// The first convolution (3 channels in one plane) pairs neighbouring taps: the second core matrix of an MMA is the next
// position of the same plane (LBO = 16 bytes - overlapping core matrices; measured correct on B200).  The perturbed
// weights of a layer (<= 9 216 values) become the B operand [cout x K] (K = tap-major, 128-byte swizzle) from 16-bit
// sources: the per-call fp16 repack of theta (already in that K order) + sign * the sigma-scaled fp16 table mirror.
// Accumulators live in TENSOR MEMORY: the raw residual stream x of a stage stays there - the second convolution of a
// residual block ACCUMULATES onto it (x += conv(..) is the MMA's own accumulate) - and only BN'd / ReLU'd fp16 operand
// planes ever touch shared memory.  Stage convolutions run in bands of 8 rows over TWO band accumulators (the MMA warp -
// one lane issues, descriptors advance by adds - works a band ahead of the workers), each band goes through a 9-row fp32
// band buffer ([row][channel quad][x] float4: conflict-free for the pool), is max-pooled 3x3 / 2, and the pooled raw map
// is stored back into TMEM (tcgen05.st) at its position of the next geometry.
// Measured on B200 (C5 share, 256 pairs): 587 us against 762 us for the mma.sync trunk (level 2) and 1 590 us for the
// first tcgen05 trunk (im2col copies through shared memory).  An MMA of M = 128, K = 16 costs 63 cycles for EVERY N <= 128
// (scripts/probe_umma.cu; 78 in this kernel) - the pipe's floor is max(64, N / 2) - so stage 0 (564 MMAs per member, N = 16)
// is tensor-issue bound and the later stages wait for the weight build.
//
// Dense tail (Linear 2048 -> 256, LSTM 513 -> 1024: 90 % of the parameters): "swap AB" GEMMs whose WEIGHT tiles
// [128 rows x 64 k] are the A operand and arrive by TMA straight from an fp16 repack of theta and from the sigma-scaled
// fp16 mirror of the noise table (csrc/direct_common.cuh: x.(theta + s*sigma*eps)^T = x.theta^T + s*(x.(sigma*eps)^T), so
// the perturbed weights are never built); the activations of the pair's members are the B operand (N = 16 columns).  The
// 257-wide rows of weight_ih are not 16-byte aligned: rows r = 8q + c form class c, whose rows are 8*257 elements apart -
// a legal TMA stride - so an M tile is a class (gate row of lane q = 8q + c); the reward column is added in the epilogue.
// fp16 operands (10-bit mantissa, the precision of tf32), fp32 accumulate; BN folds, biases, LSTM cell, policy head fp32.
#include <type_traits>
#include "impala_tail.cuh"

namespace {

constexpr int IT_WORKERS = 384, IT_THREADS = IT_WORKERS + 32, IT_GROUPS = IT_WORKERS / 128;   // worker warps come in groups of four TMEM lane quarters
// shared memory map (bytes from the 1024-aligned base).  Operand maps are PLANES: plane j holds channels 8j .. 8j + 7 of
// every position of the zero-bordered (W + 2) x (W + 2) map, 16 bytes per position, positions in row-major order - the
// UMMA no-swizzle K-major layout with the POSITION as the row: a core matrix is 8 consecutive positions (128 contiguous
// bytes), the next 8-row group follows at +128 (SBO), the next 8 channels are one plane further (LBO).  A filter tap
// (dy, dx) is then nothing but the descriptor's start address moved by (dy * (W + 2) + dx) positions: no im2col copy.
constexpr int IT_M0 = 0, IT_M0_BYTES = 38912;                 // operand map buffer 0 (the B operands behind it need 1024-byte aligned swizzle atoms)
constexpr int IT_B = IT_M0 + IT_M0_BYTES, IT_B_BYTES = 20480; // weights of a layer: up to 5 boxes of [32 x 64]; two buffers
constexpr int IT_XH_BYTES = 38400;                            // one operand map (34*34*16*2 = 36 992 is the largest)
constexpr int IT_M1 = IT_B + 2 * IT_B_BYTES, IT_M2 = IT_M1 + IT_XH_BYTES;   // buffers 1, 2; the 66 x 66 frame map (69 696 B) spans both
constexpr int IT_BAND = IT_M2 + IT_XH_BYTES, IT_BAND_BYTES = 36864;   // 9 conv rows x (C / 4) x W float4 (fp32)
constexpr int IT_FCIN = IT_BAND + IT_BAND_BYTES;              // fp16 [2][2048]: BN'd trunk outputs of the two members
constexpr int IT_PAR = IT_FCIN + 8192, IT_PAR_WORDS = 96;     // per layer (15, execution order): sN[32] tN[32] bias[32]
constexpr int IT_SMEM = IT_PAR + 15 * IT_PAR_WORDS * 4 + 1024;
static_assert(IT_SMEM <= 232448, "shared memory");
static_assert(IT_B % 1024 == 0 && IT_B_BYTES % 1024 == 0 && IT_BAND % 1024 == 0, "swizzle atoms");
constexpr int IT_XT = IT_BAND;                                // dense tail: x tiles (impala_tail.cuh) in the band buffer
static_assert(IT_M2 + IT_XH_BYTES >= 131072 + 2 * 5120, "the dense tail's tile ring and core / h0 tiles live below the band buffer");
// tensor memory columns of the trunk
constexpr uint32_t TC_R = 0, TC_Y = 160, TC_P = 320, TC_PW = 96;   // residual stream (<= 144) | block-internal map | two stage-conv band accumulators (<= 96 each)
// barriers: the tail's (impala_tail.cuh; TB_MMA doubles as "layer / band done" in the trunk), then "operands ready"
// IB_PD + i: band accumulator i holds a finished band; IB_PF + i: the workers have read it (the MMA warp runs a band ahead)
enum { IB_MMA = TB_MMA, IB_GO = TB_COUNT, IB_PD = IB_GO + 1, IB_PF = IB_PD + 2, IB_COUNT = IB_PF + 2 };

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
    const __half* w16;      // convolution weights of theta, fp16, [oc][tap][ci] per layer (impala_theta16_kernel)
    const __half* e16;      // the member's slice of the sigma-scaled fp16 table mirror: e16[p] = fp16(fl32(sigma * eps[p]))
    float sgn;              // the member's sign (0: unperturbed)
    float sg;
    __device__ __forceinline__ float par(int p) const { return perturb1(theta[p], sg, row[p]); }
};

// Everything a layer needs that is not its input map: weights -> B operand [cout rows x K] fp16, K-major 128-byte swizzle
// in 64-wide boxes.  K order = the order the MMAs walk the map: 16-channel layers k = tap * 16 + c (one MMA per tap);
// 32-channel layers k = tap * 32 + c (two MMAs per tap: planes 0-1, planes 2-3); the first convolution (3 channels in a
// single plane) pairs neighbouring taps - the second core matrix of an MMA is the NEXT position of the same plane (LBO =
// 16 bytes) - so k = (dy * 2 + (dx >> 1)) * 16 + (dx & 1) * 8 + c, K = 96 with zeros for the fourth tap of a row and the
// five missing channels.  Also: conv bias; input-side BN of the NEXT convolution folded to scale / shift (applied by THIS
// layer's epilogue).  par: sN[32] tN[32] bias[32].
template <int CIN>
__device__ __noinline__ void it_prep_layer_t(const ItCtx& c, const ConvP& p, uint8_t* Bs, float* par, const ConvP* nxt, const __half* w16_layer) {
    constexpr bool first = CIN == 3;
    const int tid = threadIdx.x;
    constexpr int k9 = CIN * 9;
    if (first) {
        // phantom tap and padded channels are zeros; 432 weights: one per thread and round, generic index arithmetic
        for (int i = tid; i < 2 * 16 * 128 / 16; i += IT_WORKERS) reinterpret_cast<uint4*>(Bs)[i] = make_uint4(0, 0, 0, 0);
        it_wsync();
        const int n = p.cout * k9;
#pragma unroll 1
        for (int t = tid; t < n; t += IT_WORKERS) {
            const int oc = t / k9, rem = t - oc * k9, ci = rem / 9, tap = rem - ci * 9;
            const int dy = (tap * 11) >> 5, dx = tap - 3 * dy;
            const int k = (dy * 2 + (dx >> 1)) * 16 + (dx & 1) * 8 + ci;
            *reinterpret_cast<__half*>(Bs + (k >> 6) * (p.cout * 128) + oc * 128 + ((((k & 63) >> 3) ^ (oc & 7)) << 4) + (k & 7) * 2) =
                __float2half_rn(c.par(p.w + t));
        }
    } else {
        // A lane is an INPUT CHANNEL (the dimension that is contiguous in the K-major operand): the 32 lanes of a warp
        // write 64 contiguous-but-swizzled bytes per tap - no bank conflicts, no index arithmetic beyond shifts.  The
        // sources are 16-bit: theta's weights from the per-call fp16 repack ALREADY in this K order (a warp reads 64
        // contiguous bytes per tap), eps from the sigma-scaled fp16 mirror of the table in its own order (stride of 9
        // halves; the nine loads of an output channel walk the same five lines, so all but the first hit L1): 4 bytes per
        // weight instead of 8 - the build is bound by L2 -> SM delivery while the other SMs stream their dense-tail tiles -
        // and the same split of one rounding into two as the dense tail's (fp16(theta) + s * fp16(sigma eps), summed in
        // fp32, rounded once more).  Measured alternatives with fp32 sources (coalesced loads + 2-byte scatter stores
        // with or without an offset table; a two-pass build through a staging buffer): all at 5 - 8 k cycles per 32 -> 32
        // layer.
        // A lane is a PAIR of input channels: theta's two halves are one 32-bit load, the two results leave as one 32-bit
        // store (half as many shared-memory store instructions and no 2-byte ones: the build tracked the number of
        // sub-word memory instructions in every variant measured), eps stays two 2-byte loads at a stride of 9 halves.
        constexpr int LPO = CIN / 2, OPW = 32 / LPO;                  // lanes per output channel, output channels per warp and round
        const int w = tid >> 5, l = tid & 31, ci = 2 * (l & (LPO - 1)), o2 = l / LPO;
        const int rowb = p.cout * 128;
        const unsigned short* w16 = reinterpret_cast<const unsigned short*>(w16_layer);
        const unsigned short* e16 = reinterpret_cast<const unsigned short*>(c.e16 + p.w);
#pragma unroll 1
        for (int oc = w * OPW + o2; oc < p.cout; oc += (IT_WORKERS / 32) * OPW) {
            const uint32_t* wt = reinterpret_cast<const uint32_t*>(w16 + oc * k9 + ci);      // [oc][tap][ci]: 4-byte aligned (ci even, CIN even)
            const unsigned short* we = e16 + oc * k9 + ci * 9;
            uint32_t a[9];
            unsigned short e0[9], e1[9];
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) { a[tap] = __ldg(wt + tap * (CIN / 2)); e0[tap] = __ldg(we + tap); e1[tap] = __ldg(we + 9 + tap); }
            uint8_t* dst = Bs + oc * 128 + (ci & 7) * 2;
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
                const int k = tap * CIN + ci;
                const float2 th = __half22float2(*reinterpret_cast<const __half2*>(&a[tap]));
                const float v0 = fmaf(c.sgn, __half2float(__ushort_as_half(e0[tap])), th.x);                     // exact in fp32
                const float v1 = fmaf(c.sgn, __half2float(__ushort_as_half(e1[tap])), th.y);
                *reinterpret_cast<uint32_t*>(dst + (k >> 6) * rowb + ((((k & 63) >> 3) ^ (oc & 7)) << 4)) = dr_pack(v0, v1);
            }
        }
    }
}

// layer li of the trunk in execution order: per stage the stage convolution, then the two residual blocks (a, b each)
__device__ __forceinline__ const ConvP& it_layer(const ImpalaP& L, int li) {
    const int s = li / 5, r = li - 5 * s;
    return r == 0 ? L.feat[s] : L.res[(r - 1) >> 1][s][(r - 1) & 1];
}
// Everything a layer's epilogue needs besides the accumulator, for ALL 15 layers of a member at once (one round of loads
// in flight instead of a dependent chain per layer): par[li] = sN[32] tN[32] bias[32] - the layer's own conv bias and the
// input-side BN of the NEXT convolution folded to scale / shift (applied by THIS layer's epilogue), exactly perturbed.
__device__ __noinline__ void it_prep_params(const ImpalaP& L, const ItCtx& c, float* par_all) {
#pragma unroll 1
    for (int i = threadIdx.x; i < 15 * 64; i += IT_WORKERS) {
        const int li = i >> 6, ch = i & 31;
        float* par = par_all + li * IT_PAR_WORDS;
        if (i & 32) {
            if (li < 14) {
                const ConvP& nx = it_layer(L, li + 1);
                if (ch < nx.cin) {
                    const float inv = 1.0f / sqrtf(c.bn[nx.bv + ch] + 1e-5f);
                    const float s = c.par(nx.g + ch) * inv;
                    par[ch] = s;
                    par[32 + ch] = c.par(nx.be + ch) - c.bn[nx.bm + ch] * s;
                }
            }
        } else {
            const ConvP& p = it_layer(L, li);
            if (ch < p.cout) par[64 + ch] = c.par(p.b + ch);
        }
    }
}

__device__ __forceinline__ void it_prep_layer(const ImpalaP& L, const ItCtx& c, const ConvP& p, bool first, uint8_t* Bs, float* par, const ConvP* nxt) {
    int seg = 0;
    while (seg < 14 && L.seq_w[seg] != p.w) ++seg;
    const __half* w16 = c.w16 + L.seq_o[seg];
    if (first) it_prep_layer_t<3>(c, p, Bs, par, nxt, w16);
    else if (p.cin == 16) it_prep_layer_t<16>(c, p, Bs, par, nxt, w16);
    else it_prep_layer_t<32>(c, p, Bs, par, nxt, w16);
}

template <int C>
__device__ __forceinline__ void it_tmem_ld(uint32_t taddr, float* v) {
    if (C == 16) tmem_ld16(taddr, v);
    else tmem_ld32(taddr, v);
}

// workers -> MMA warp: everything the next batch of MMAs reads (operand map, weights) is written and visible to the
// tensor core's proxy, and every TMEM access of the previous epilogue has retired
__device__ __forceinline__ void it_go(uint32_t bar) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    it_wsync();
    if (threadIdx.x == 0) dr_arrive(bar);
}
__device__ __forceinline__ void it_layer_wait(uint32_t bar0, uint32_t& mph) {
    if ((threadIdx.x & 31) == 0) dr_wait(bar0 + 8u * (uint32_t)IB_MMA, mph);
    __syncwarp();
    mph ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void it_zero(uint8_t* xh, int bytes) {
    for (int i = threadIdx.x; i < bytes / 16; i += IT_WORKERS) reinterpret_cast<uint4*>(xh)[i] = make_uint4(0, 0, 0, 0);
}
// eight BN'd (optionally ReLU'd) channels of one position -> one 16-byte chunk of a plane
__device__ __forceinline__ uint4 it_pack8(const float* v, const float* sc, const float* sh, const float* bias, bool relu) {
    uint32_t o[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        float u0 = fmaf(v[2 * g] + (bias ? bias[2 * g] : 0.f), sc[2 * g], sh[2 * g]);
        float u1 = fmaf(v[2 * g + 1] + (bias ? bias[2 * g + 1] : 0.f), sc[2 * g + 1], sh[2 * g + 1]);
        if (relu) { u0 = fmaxf(u0, 0.f); u1 = fmaxf(u1, 0.f); }
        o[g] = dr_pack(u0, u1);
    }
    return make_uint4(o[0], o[1], o[2], o[3]);
}

// MMA-warp side: all MMAs of a batch of `nt` output tiles of 128 positions (tile T starts at position q0 + 128 T, relative
// to the first interior pixel; accumulator columns d_tmem + T * cout).  map: shared address of the input map, Wp its
// padded width, cin its channels (3: the first convolution).  TAPS OUTER, TILES INNER: consecutive MMAs write different
// accumulators - back-to-back MMAs onto the same small accumulator wait for each other (measured: ~100 cycles each).
__device__ __forceinline__ void it_umma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem), "l"(adesc),
                 "l"(bdesc), "r"(idesc), "r"(acc)
                 : "memory");
}
// called by ONE lane (which also commits): the inner loop is two adds and the MMA
__device__ __forceinline__ void it_mma_batch(uint32_t map, int Wp, int cin, uint32_t Bb, int cout, int q0, int nt, uint32_t d_tmem, bool accumulate) {
    const uint32_t idesc = dr_idesc(cout, 0);
    const uint32_t plane = (uint32_t)(Wp * Wp * 16);
    const int halves = cin == 3 ? 1 : cin >> 4, nj = cin == 3 ? 6 : 9 * halves;
#pragma unroll 1
    for (int j = 0; j < nj; ++j) {
        uint32_t a, lbo;
        if (cin == 3) {                                       // (dy, tap pair): taps dx = 2 * (j & 1), + 1; second core matrix = next position
            a = map + (uint32_t)((q0 + (j >> 1) * Wp + (j & 1) * 2) * 16);
            lbo = 16;
        } else {                                              // K = 16 per MMA: two planes
            const int tap = halves == 1 ? j : (j >> 1), hf = halves == 1 ? 0 : (j & 1), dy = (tap * 11) >> 5, dx = tap - 3 * dy;
            a = map + (uint32_t)hf * 2u * plane + (uint32_t)((q0 + dy * Wp + dx) * 16);
            lbo = plane;
        }
        const uint64_t bdesc = make_desc_sw128(Bb + (uint32_t)((j >> 2) * cout * 128)) + (uint64_t)((j & 3) * 2);
        const uint32_t acc = (j != 0 || accumulate) ? 1u : 0u;
        uint64_t adesc = make_desc(a, lbo, 128);
        uint32_t d = d_tmem;
#pragma unroll 3
        for (int T = 0; T < nt; ++T, adesc += 128, d += (uint32_t)cout) it_umma(d, adesc, bdesc, idesc, acc);     // next tile: 128 positions = 2048 bytes
    }
}

__global__ void __launch_bounds__(IT_THREADS, 1)
impala_direct_kernel(const __grid_constant__ ImpalaP L, const __grid_constant__ ItMaps maps, const float* __restrict__ replicas, int64_t stride,
                     const float* __restrict__ theta, const float* __restrict__ bnbuf, const int64_t* __restrict__ idx,
                     const int8_t* __restrict__ sign, float sigma, const float* __restrict__ frame, const float* __restrict__ reward,
                     const uint8_t* __restrict__ done, const float* __restrict__ h_in, const float* __restrict__ c_in, int E,
                     float* __restrict__ probs, float* __restrict__ h_out, float* __restrict__ c_out, int n_members, int pair_mode,
                     const __half* __restrict__ conv16, const __half* __restrict__ mirror16, long long* __restrict__ prof) {
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
        for (int s = 0; s < IB_COUNT; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(IT_BAR(s)));
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
    int li = 0;                          // running layer number: the same sequence in the workers and the MMA warp
    stamp();

    if (worker) {
        auto Bbuf = [&](int l) { return sm + IT_B + (l & 1) * IT_B_BYTES; };
        auto Pbuf = [&](int l) { return par_s + (l % 15) * IT_PAR_WORDS; };
        const int r128 = q4 * 32 + lane;                        // this thread's TMEM lane = row of every tile
        int gb = 0;                                             // running band number (both roles): accumulator gb & 1, its use number gb >> 1

#pragma unroll 1
        for (int mem = 0; mem < nmem; ++mem) {
            ItCtx c;
            c.theta = theta; c.bn = bnbuf; c.sg = sigma * (float)sgi[mem]; c.row = rows[mem];
            c.w16 = conv16; c.e16 = mirror16 + ids[mem]; c.sgn = (float)sgi[mem];
            uint8_t* xa = sm + IT_M1;                           // the frame map spans buffers 1 and 2
            uint8_t* xb = sm + IT_M0;
            // the member's convolution parameters (the first 0.4 MB of its table row) and its frame: into L2 now, so that
            // the weight builds of the 15 layers find them there
            if (tid < 8 && (mem == 0 || !shared_row)) {
                const int span = L.fc_g, part = (span + 7) / 8, lo = tid * part;
                if (lo < span) l2_prefetch(c.e16 + lo, (size_t)min(part, span - lo) * 2);
            } else if (tid == 32) {
                l2_prefetch(frame + (int64_t)inst[mem] * 12288, 12288 * 4);
                if (mem == 0 && nmem == 2) l2_prefetch(frame + (int64_t)inst[1] * 12288, 12288 * 4);
            }
            it_wsync();
            // ---- frame / 255 -> BN of the first convolution -> fp16 plane [66 x 66][8] (3 channels + 5 zeros), zero border ----
            float* fpar = reinterpret_cast<float*>(band);        // scale / shift of the three input channels
            it_zero(xa, 2 * IT_XH_BYTES);
            if (tid < 3) {
                const ConvP& p0 = L.feat[0];
                const float inv = 1.0f / sqrtf(bnbuf[p0.bv + tid] + 1e-5f);
                const float s = c.par(p0.g + tid) * inv;
                fpar[tid] = s;
                fpar[4 + tid] = c.par(p0.be + tid) - bnbuf[p0.bm + tid] * s;
            }
            it_prep_layer(L, c, L.feat[0], true, Bbuf(li), Pbuf(li), &L.res[0][0][0]);
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
                    const uint32_t lo = dr_pack(fmaf(f[3 * u] * (1.0f / 255.0f), fpar[0], fpar[4]), fmaf(f[3 * u + 1] * (1.0f / 255.0f), fpar[1], fpar[5]));
                    const uint32_t hi = dr_pack(fmaf(f[3 * u + 2] * (1.0f / 255.0f), fpar[2], fpar[6]), 0.f);
                    *reinterpret_cast<uint4*>(xa + (((p >> 6) + 1) * 66 + (p & 63) + 1) * 16) = make_uint4(lo, hi, 0u, 0u);
                }
            }
            if (mem == 0) stamp();
            auto stage = [&](auto Cc, const int s) {
                constexpr int C = decltype(Cc)::value, C4 = C >> 2;
                // =========== stage convolution (BN on the input, no ReLU) + max-pool 3x3 / 2 pad 1 ===========
                const int Wc = 64 >> s, Wpc = Wc + 2;
                const int Wo = Wc >> 1, Wpo = Wo + 2, planeo = Wpo * Wpo * 16;
                const float* par = Pbuf(li);
                const float *sN = par, *tN = par + 32, *bias = par + 64;
                it_zero(xb, IT_XH_BYTES);                          // becomes the pooled operand map (new geometry)
                const int ntb = (8 * Wpc + 127) >> 7;           // 5, 3, 2 tiles per band of 8 conv rows
                const int ntR = (Wo * Wpo - 2 + 127) >> 7;      // tiles of the residual stream: 9, 3, 1
                const ConvP& pa0 = L.res[0][s][0];
                float4* band4 = reinterpret_cast<float4*>(band);
                it_go(IT_BAR(IB_GO));                            // the stage's map and weights are visible: the MMA warp may start its bands
                if (s == 0) it_prep_params(L, c, par_s);        // all 15 layers' folded parameters, under the first bands' MMAs (first used by the band epilogue)
                it_prep_layer(L, c, pa0, false, Bbuf(li + 1), Pbuf(li + 1), &L.res[0][s][1]);   // the first block convolution's weights, under the MMAs
#pragma unroll 1
                for (int b = 0; b < Wc / 8; ++b, ++gb) {
                    auto bst = [&](int i) { if (prof != nullptr && blockIdx.x == 7 && tid == 0 && mem == 0 && s == 0 && b == 3) prof[48 + i] = clock64(); };
                    const uint32_t pbuf = (uint32_t)(gb & 1), tcp = TC_P + pbuf * TC_PW;
                    bst(0);
                    it_wsync();                                  // the previous band's pool has read the rows this band overwrites
                    bst(1);
                    if (lane == 0) dr_wait(IT_BAR(IB_PD + pbuf), (uint32_t)((gb >> 1) & 1));
                    __syncwarp();
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    bst(2);
                    // band epilogue: conv rows 8b .. 8b + 7 (+ bias) -> circular band [row % 9][c / 4][x] float4
#pragma unroll 1
                    for (int tg = wg; tg < ntb; tg += IT_GROUPS) {
                        const int ql = tg * 128 + r128, yl = ql / Wpc, x = ql - yl * Wpc;
                        float v[32];
                        if (C == 16) it_tmem_ld<16>(tmem + lane_sel + tcp + (uint32_t)(tg * 16), v);
                        else it_tmem_ld<32>(tmem + lane_sel + tcp + (uint32_t)(tg * 32), v);
                        if (yl < 8 && x < Wc) {
                            float4* dst = band4 + (((8 * b + yl) % 9) * C4) * Wc + x;
#pragma unroll
                            for (int g = 0; g < 8; ++g)
                                if (g < C4)
                                    dst[g * Wc] = make_float4(v[4 * g] + bias[4 * g], v[4 * g + 1] + bias[4 * g + 1], v[4 * g + 2] + bias[4 * g + 2],
                                                              v[4 * g + 3] + bias[4 * g + 3]);
                        }
                    }
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    bst(3);
                    it_wsync();
                    if (tid == 0) dr_arrive(IT_BAR(IB_PF + pbuf));  // the accumulator may be overwritten (band b + 2)
                    bst(4);
                    // pool: pooled rows 4b .. 4b + 3 = positions [4b * Wpo, (4b + 4) * Wpo) of the new geometry; position q
                    // lives in lane q % 128 of residual tile q / 128
                    {
                        const int qlo = 4 * b * Wpo, tile = (qlo >> 7) + (tid >> 7), q = tile * 128 + (tid & 127);
                        const bool in_range = q >= qlo && q < qlo + 4 * Wpo;
                        if (tile < ntR && __any_sync(0xffffffffu, in_range)) {
                            const int py = q / Wpo, px = q - py * Wpo;
                            const bool valid = in_range && px < Wo;
                            float m[32];
                            const uint32_t rt = tmem + lane_sel + TC_R + (uint32_t)(tile * C);
                            if (C == 16) it_tmem_ld<16>(rt, m);           // lanes outside this band keep what they hold
                            else it_tmem_ld<32>(rt, m);
                            if (valid) {
#pragma unroll
                                for (int ch = 0; ch < 32; ++ch) m[ch] = -INFINITY;
                                for (int dy = -1; dy <= 1; ++dy) {
                                    const int yy = 2 * py + dy;
                                    if (yy < 0 || yy >= Wc) continue;
                                    for (int dx = -1; dx <= 1; ++dx) {
                                        const int xx = 2 * px + dx;
                                        if (xx < 0 || xx >= Wc) continue;
                                        const float4* src = band4 + ((yy % 9) * C4) * Wc + xx;
#pragma unroll
                                        for (int g = 0; g < 8; ++g) {
                                            if (g < C4) {
                                                const float4 f = src[g * Wc];
                                                m[4 * g] = fmaxf(m[4 * g], f.x); m[4 * g + 1] = fmaxf(m[4 * g + 1], f.y);
                                                m[4 * g + 2] = fmaxf(m[4 * g + 2], f.z); m[4 * g + 3] = fmaxf(m[4 * g + 3], f.w);
                                            }
                                        }
                                    }
                                }
                            }
                            // raw pooled map -> TMEM residual stream; BN + ReLU of the first block convolution -> operand map
                            it_tmem_st16(rt, m);
                            if (C == 32) it_tmem_st16(rt + 16u, m + 16);
                            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                            if (valid) {
                                uint8_t* dst = xb + (q + Wpo + 1) * 16;
#pragma unroll
                                for (int j = 0; j < 4; ++j)
                                    if (8 * j < C) *reinterpret_cast<uint4*>(dst + j * planeo) = it_pack8(m + 8 * j, sN + 8 * j, tN + 8 * j, nullptr, true);
                            }
                        }
                    }
                    bst(5);
                }
                ++li;
                { uint8_t* t_ = xa; xa = xb; xb = t_; }            // xa: operand map of the first block convolution
                it_wsync();
                it_zero(xb, IT_XH_BYTES);                              // old-geometry map: becomes the block-internal operand map
                if (mem == 0) stamp();
                // =========== two residual blocks: x += conv_b(relu(BN_b(conv_a(relu(BN_a(x)))))) ===========
                const int W = Wo, Wp = Wpo, plane = planeo, ntile = ntR;
#pragma unroll 1
                for (int blk = 0; blk < 2; ++blk) {
                    const ConvP& pb = L.res[blk][s][1];
                    const bool last = (s == 2 && blk == 1);
                    const ConvP* nxt = last ? nullptr : (blk == 0 ? &L.res[1][s][0] : &L.feat[s + 1]);
                    // ---- conv a -> Y; epilogue: relu(BN_b(y + bias)) -> xb ----
                    {
                        const float* pr = Pbuf(li);
                        const float *sA = pr, *tA = pr + 32, *bA = pr + 64;
                        auto fine = [&](int i) { if (prof != nullptr && blockIdx.x == 7 && tid == 0 && mem == 0 && blk == 0) prof[16 + 5 * s + i] = clock64(); };
                        fine(0);
                        it_go(IT_BAR(IB_GO));
                        fine(1);
                        it_prep_layer(L, c, pb, false, Bbuf(li + 1), Pbuf(li + 1), nxt);     // conv b's weights, under conv a's MMAs
                        fine(2);
                        it_layer_wait(bar0, mph);
                        fine(3);
#pragma unroll 1
                        for (int T = wg; T < ntile; T += IT_GROUPS) {
                            const int q = T * 128 + r128, y = q / Wp, x = q - y * Wp;
                            float v[32];
                            if (C == 16) it_tmem_ld<16>(tmem + lane_sel + TC_Y + (uint32_t)(T * 16), v);
                            else it_tmem_ld<32>(tmem + lane_sel + TC_Y + (uint32_t)(T * 32), v);
                            if (y < W && x < W) {
                                uint8_t* dst = xb + (q + Wp + 1) * 16;
#pragma unroll
                                for (int j = 0; j < 4; ++j)
                                    if (8 * j < C) *reinterpret_cast<uint4*>(dst + j * plane) = it_pack8(v + 8 * j, sA + 8 * j, tA + 8 * j, bA + 8 * j, true);
                            }
                        }
                        fine(4);
                        ++li;
                    }
                    // ---- conv b accumulates ONTO the residual stream in TMEM; epilogue: x += bias, next operand map -> xa ----
                    {
                        const float* pr = Pbuf(li);
                        const float *sB = pr, *tB = pr + 32, *bB = pr + 64;
                        const bool relu_next = blk == 0;             // next consumer: a block convolution (BN + ReLU) or a stage convolution (BN only)
                        it_go(IT_BAR(IB_GO));
                        if (nxt != nullptr) {                        // the next layer's weights, under conv b's MMAs
                            const bool nxt_stage = blk == 1;
                            it_prep_layer(L, c, *nxt, false, Bbuf(li + 1), Pbuf(li + 1), nxt_stage ? &L.res[0][s + 1][0] : &L.res[1][s][1]);
                        }
                        it_layer_wait(bar0, mph);
                        float* fscr = band;                          // relu(x) of the last layer, [c * 64 + pixel]
#pragma unroll 1
                        for (int T = wg; T < ntile; T += IT_GROUPS) {
                            const int q = T * 128 + r128, y = q / Wp, x = q - y * Wp;
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
                            if (y < W && x < W) {
                                if (!last) {
                                    uint8_t* dst = xa + (q + Wp + 1) * 16;
#pragma unroll
                                    for (int j = 0; j < 4; ++j)
                                        if (8 * j < C) *reinterpret_cast<uint4*>(dst + j * plane) = it_pack8(v + 8 * j, sB + 8 * j, tB + 8 * j, nullptr, relu_next);
                                } else {
#pragma unroll
                                    for (int ch = 0; ch < 32; ++ch) fscr[ch * 64 + y * 8 + x] = fmaxf(v[ch], 0.f);
                                }
                            }
                        }
                        ++li;
                        if (last) {
                            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                            it_wsync();
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
            };
            stage(std::integral_constant<int, 16>(), 0);
            stage(std::integral_constant<int, 32>(), 1);
            stage(std::integral_constant<int, 32>(), 2);
        }
        // ======================================= dense tail =======================================
        stamp_i = 11;
        stamp();
        tl_dense_tail_workers<IT_WORKERS, IT_XT>(L, targs, sm, s0, fcin, bar0, tmem, mph);
        stamp();
    } else {
        // =============================== MMA issue warp of the trunk ===============================
        // mirrors the workers' sequence of layers / bands: waits until the operands are ready, issues every tile of the
        // batch (a filter tap = a start address), and commits: the commit tells the workers the accumulators are ready
        uint32_t gph = 0;
        int gb = 0;
        if (lane == 0) {
        auto go_wait = [&]() {
            dr_wait(IT_BAR(IB_GO), gph);
            gph ^= 1u;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        };
#pragma unroll 1
        for (int mem = 0; mem < nmem; ++mem) {
            uint32_t xa = s0 + IT_M1, xb = s0 + IT_M0;
#pragma unroll 1
            for (int s = 0; s < 3; ++s) {
                const int Wc = 64 >> s, Wpc = Wc + 2, C = s == 0 ? 16 : 32, cin = s == 0 ? 3 : (s == 1 ? 16 : 32);
                const int ntb = (8 * Wpc + 127) >> 7;
                go_wait();
                {
                    const uint32_t Bb = s0 + IT_B + (uint32_t)((li & 1) * IT_B_BYTES);
#pragma unroll 1
                    for (int b = 0; b < Wc / 8; ++b, ++gb) {
                        const uint32_t pbuf = (uint32_t)(gb & 1);
                        if (gb >= 2) {                           // the accumulator's previous band has been read
                            dr_wait(IT_BAR(IB_PF + pbuf), (uint32_t)(((gb >> 1) - 1) & 1));
                            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        }
                        it_mma_batch(xa, Wpc, cin, Bb, C, 8 * b * Wpc, ntb, tmem + TC_P + pbuf * TC_PW, false);
                        umma_commit(IT_BAR(IB_PD + pbuf));
                    }
                }
                ++li;
                { const uint32_t t_ = xa; xa = xb; xb = t_; }
                const int W = Wc >> 1, Wp = W + 2, ntile = (W * Wp - 2 + 127) >> 7;
#pragma unroll 1
                for (int blk = 0; blk < 2; ++blk) {
                    go_wait();
                    uint32_t Bb = s0 + IT_B + (uint32_t)((li & 1) * IT_B_BYTES);
                    it_mma_batch(xa, Wp, C, Bb, C, 0, ntile, tmem + TC_Y, false);
                    umma_commit(IT_BAR(IB_MMA));
                    ++li;
                    go_wait();
                    Bb = s0 + IT_B + (uint32_t)((li & 1) * IT_B_BYTES);
                    it_mma_batch(xb, Wp, C, Bb, C, 0, ntile, tmem + TC_R, true);
                    umma_commit(IT_BAR(IB_MMA));
                    ++li;
                }
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
    if (ctx->theta16_cap < TL_CONV16 + L.seq_o[15]) return -1;
    ItMaps maps;
    if (tl_prepare(ctx, L, theta, &maps, st)) return 3;
    DFD_CUDA(cudaFuncSetAttribute(impala_direct_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, IT_SMEM));
    const int pair_mode = n_members % 2 == 0 ? 1 : 0;
    const int grid = (pair_mode ? n_members / 2 : n_members) * obs_per_member;
    long long* prof = nullptr;
    if (getenv("DFD_IMPALA_PROF")) { cudaMalloc(&prof, 64 * 8); cudaMemset(prof, 0, 64 * 8); }
    impala_direct_kernel<<<grid, IT_THREADS, IT_SMEM, st>>>(L, maps, table->replicas, table->replica_stride, theta, bn_buffers, idx, sign,
                                                            sigma, frame, reward, done, h_in, c_in, obs_per_member, probs, h_out, c_out,
                                                            n_members, pair_mode, (const __half*)ctx->theta16 + TL_CONV16, (const __half*)ctx->scaled16, prof);
    DFD_LAUNCHED(ctx);
    if (prof) {
        cudaStreamSynchronize(st);
        long long h[64];
        cudaMemcpy(h, prof, sizeof(h), cudaMemcpyDeviceToHost);
        fprintf(stderr, "[impala tcgen05 timeline] stage 0 band 3: go %lld | wait MMA %lld | band epilogue %lld | sync %lld | pool %lld\n",
                h[49] - h[48], h[50] - h[49], h[51] - h[50], h[52] - h[51], h[53] - h[52]);
        for (int s = 0; s < 3; ++s)
            fprintf(stderr, "[impala tcgen05 timeline] stage %d first block conv a: go %lld | prep next %lld | wait MMA %lld | epilogue %lld\n", s,
                    h[17 + 5 * s] - h[16 + 5 * s], h[18 + 5 * s] - h[17 + 5 * s], h[19 + 5 * s] - h[18 + 5 * s], h[20 + 5 * s] - h[19 + 5 * s]);
        fprintf(stderr, "[impala tcgen05 timeline] CTA 7, cycles per phase of its first member: frame %lld | s0 conv+pool %lld res %lld %lld | "
                        "s1 conv+pool %lld res %lld %lld | s2 conv+pool %lld res %lld %lld | all trunks done at %lld | dense tail %lld | total %lld\n",
                h[1] - h[0], h[2] - h[1], h[3] - h[2], h[4] - h[3], h[5] - h[4], h[6] - h[5], h[7] - h[6], h[8] - h[7], h[9] - h[8],
                h[10] - h[9], h[11] - h[0], h[12] - h[11], h[12] - h[0]);
        cudaFree(prof);
    }
    return 0;
}
