// IMPALA CNN + LSTM perturbed forward (policies/impala.py:136-186), one CTA per (antithetic pair, environment).
//   x = frame/255; 3 stages {BN -> conv3x3 -> maxpool(3,2,1); 2 x [x += conv(relu(BN(conv(relu(BN(x))))))]};
//   relu -> flatten(2048) -> relu(Linear(BN1d(x))) -> concat clamp(reward,-1,1) -> LSTM cell (gates i,f,g,o,
//   state zeroed where done) -> Linear(BN1d(h)) -> softmax.
// BN layers are eval-mode with shared running statistics and per-member (perturbed) gamma / beta; every
// parameter is theta + sign*sigma*eps generated in-kernel (worker/worker.py:28).  The whole conv trunk keeps its
// activations in shared memory (two ping-pong maps + a conv-row band for the pooled stages); only the
// carried LSTM state and the action probabilities touch HBM.  precision 0: exact fp32 on CUDA cores (atol 1e-5 against
// torch CPU); precision >= 1: the convolutions run on the tensor cores (fp16 operands, fp32 accumulate, see conv3x3_mma).
// One CTA evaluates the two members of an antithetic pair (trunks one after the other, ONE pass over theta and the shared
// eps row for the dense tail of both).  The fp32 path:
//   * convolutions are register-tiled: a thread owns 4 neighbouring pixels x OCT output channels (OCT = 8 / 4 / 2 chosen so
//     that every layer fills the 512 threads), reads its 4 inputs with one 16-byte shared-memory load per (channel, row),
//     applies the input-side BN (+ReLU) once, takes the two halo pixels from the neighbouring lanes by shuffle, and issues
//     packed fp32 FMAs (FFMA2: two output channels per instruction, the weight pair comes straight out of the 16-byte
//     weight load);
//   * the pooled stages compute exactly 8 new conv rows per band into a 9-row circular band (no row is convolved twice);
//   * a layer's weights arrive in ONE round of loads (every thread's theta / eps loads issued before the first store) and
//     the NEXT layer's eps segment is prefetched into L2 while the current layer computes;
//   * the dense tail (Linear 2048 -> 256, LSTM 513 -> 1024) streams 8.4 MB of theta and eps per pair: every warp keeps
//     256 B (Linear) / 272 B (LSTM, two gate rows at a time) of loads in flight per lane.
#include "impala_tail.cuh"
#include <stdlib.h>

namespace {

constexpr int IM_THREADS = 512;
constexpr int MAP = 16384;    // floats per activation map buffer (16 x 32 x 32)
constexpr int MAP_TC = 18560; // the same map padded for the tensor-core path: 16 channels x 1160 (34 x 34 = 1156 -> 8 mod 32)
constexpr int BAND = 9248;    // conv-row band for the pooled stages: 9 rows x (64 x 16 | 32 x 32), 32 x 16 x 16; staging of a
                              // layer's weights (32 rows x 289)
constexpr int WMAX = 9216;    // largest conv weight block (32 x 32 x 3 x 3)

struct Ctx {
    const float* theta;
    const float* row;
    const float* bn;
    float sg;
    __device__ __forceinline__ float par(int p) const { return perturb1(theta[p], sg, row[p]); }
};

// weights -> wsm[(ci*9+tap)*cout + oc]; input-side BN folded to per-input-channel scale/shift; conv bias.
// One round of global loads: warp w takes the weight rows oc = w, w + 16 (each k9 = cin*9 contiguous floats, lanes over k),
// every load of the layer - BN gamma / beta / statistics and the bias included - is issued before the first store.  The
// perturbed rows go to a staging area (`stage` = the conv band buffer, free while weights are loaded) with an odd row
// stride, so the transposition into the [k][oc] layout reads and writes shared memory without bank conflicts.
__device__ void load_conv(const Ctx& c, const ConvP& p, float* wsm, float* stage, float* s_in, float* sh_in, float* bias,
                          bool tc_layer) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int k9 = p.cin * 9, n = p.cout * k9;
    const int ss = k9 | 1;                     // staging row stride (odd)
    float a[2][9], e[2][9];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int oc = warp + 16 * r;
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            const int k = lane + 32 * i;
            a[r][i] = 0.f;
            e[r][i] = 0.f;
            if (oc < p.cout && k < k9) {
                a[r][i] = c.theta[p.w + oc * k9 + k];
                e[r][i] = c.row[p.w + oc * k9 + k];
            }
        }
    }
    // tid < 32: input channel tid (BN scale / shift); 32 <= tid < 64: output channel tid - 32 (bias)
    float g_t = 0.f, g_e = 0.f, b_t = 0.f, b_e = 0.f, bm = 0.f, bv = 1.f;
    if (tid < p.cin) {
        g_t = c.theta[p.g + tid]; g_e = c.row[p.g + tid];
        b_t = c.theta[p.be + tid]; b_e = c.row[p.be + tid];
        bm = c.bn[p.bm + tid]; bv = c.bn[p.bv + tid];
    } else if (tid >= 32 && tid - 32 < p.cout) {
        b_t = c.theta[p.b + tid - 32]; b_e = c.row[p.b + tid - 32];
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int oc = warp + 16 * r;
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            const int k = lane + 32 * i;
            if (oc < p.cout && k < k9) stage[oc * ss + k] = perturb1(a[r][i], c.sg, e[r][i]);
        }
    }
    if (tid < p.cin) {
        const float inv = 1.0f / sqrtf(bv + 1e-5f);
        const float sc = perturb1(g_t, c.sg, g_e) * inv;
        s_in[tid] = sc;
        sh_in[tid] = perturb1(b_t, c.sg, b_e) - bm * sc;
    } else if (tid >= 32 && tid - 32 < p.cout) {
        bias[tid - 32] = perturb1(b_t, c.sg, b_e);
    }
    __syncthreads();
    const int lc = p.cout == 32 ? 5 : 4;       // cout is 16 or 32
    if (!tc_layer) {
        for (int t = tid; t < n; t += IM_THREADS) {
            const int k = t >> lc, oc = t & (p.cout - 1);
            wsm[t] = stage[oc * ss + k];
        }
    } else {
        // tensor-core layer (conv3x3_mma / conv_first_mma): weights as packed fp16 pairs in the order the B fragments of
        // mma.m16n8k16 want them: the two registers (j = 0, 1) of a lane sit side by side (one 8-byte load), the output
        // channel is XOR-swizzled by tig (conflict-free).  Word (((kb*4 + tig)*cout + (oc ^ (tig << 2)))*2 + j) =
        // {lo: W[kk][oc], hi: W[kk + 4][oc]} with
        //   cin >= 16: kb = (16-channel block)*9 + tap, kk = channel block*16 + 8j + tig at that tap;
        //   cin = 3 (first convolution): K = 27 (ci, tap) values padded to 32, kb = 16-value step, kk = 16*kb + 8j + tig.
        uint32_t* wpk = reinterpret_cast<uint32_t*>(wsm);
        const bool first = p.cin < 16;
        const int nw = first ? 16 * p.cout : n / 2;
        for (int t = tid; t < nw; t += IM_THREADS) {
            const int oc = t & (p.cout - 1), r = t >> lc;
            const int jj = r & 7, kb = r >> 3, tg = jj & 3, j = jj >> 2;
            float lo, hi;
            if (first) {
                const int kk = 16 * kb + 8 * j + tg;
                lo = kk < 27 ? stage[oc * ss + kk] : 0.f;
                hi = kk + 4 < 27 ? stage[oc * ss + kk + 4] : 0.f;
            } else {
                const int blk = kb / 9, tap = kb - blk * 9;
                const int ch = blk * 16 + 8 * j + tg;
                lo = stage[oc * ss + ch * 9 + tap];
                hi = stage[oc * ss + (ch + 4) * 9 + tap];
            }
            uint32_t u;
            asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(u) : "f"(hi), "f"(lo));
            wpk[(((kb * 4 + tg) * p.cout + (oc ^ (tg << 2))) << 1) + j] = u;
        }
    }
}

// pull the eps segment of a later layer towards L2 while the current layer computes (theta is L2-hot: every CTA reads it)
__device__ __forceinline__ void prefetch_l2(const float* p, int n) {
    for (int t = threadIdx.x * 32; t < n; t += IM_THREADS * 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + t));
}

// conv3x3 pad 1 over output rows [r0, r1) of an H x W map (W a multiple of 4).  in: [cin][H][W]; the input is BN'd
// (scale/shift) and optionally ReLU'd on the fly, zero padding applies AFTER that (torch pads the BN output).
// dst element (oc, r, x) at dst[oc*dst_cs + ((r - r0 + dst_r0) % dst_rmod)*W + x]; ACCUM adds to dst (residual).
// A work item = 4 neighbouring pixels of one row x OCT output channels; consecutive lanes take consecutive 4-pixel
// tiles (W/4 divides 32, so lane 0 / lane 31 always sit on a row edge and the halo shuffles never cross an item group).
template <int OCT, bool RELU_IN, bool ACCUM>
__device__ __forceinline__ void conv3x3(const float* __restrict__ in, int cin, int H, int W, const float* __restrict__ wsm,
                                        int cout, const float* __restrict__ s_in, const float* __restrict__ sh_in,
                                        const float* __restrict__ bias, int r0, int r1, float* __restrict__ dst, int dst_cs,
                                        int dst_r0, int dst_rmod) {
    static_assert(OCT == 8 || OCT == 4 || OCT == 2, "OCT");
    const int lane = threadIdx.x & 31;
    const int tpr = W >> 2, ntile = (r1 - r0) * tpr, nitem = ntile * (cout / OCT);
    const int HW = H * W;
    for (int base = threadIdx.x & ~31; base < nitem; base += IM_THREADS) {
        const int item = base + lane;
        const bool valid = item < nitem;
        const int it = valid ? item : nitem - 1;
        const int og = it / ntile, tile = it - og * ntile;
        const int tr = tile / tpr;
        const int r = r0 + tr, x0 = (tile - tr * tpr) << 2;
        const bool ledge = x0 == 0, redge = x0 + 4 == W;
        float2 acc[OCT / 2][4];
#pragma unroll
        for (int j = 0; j < OCT / 2; ++j) {
            const float2 b = *reinterpret_cast<const float2*>(bias + og * OCT + 2 * j);
#pragma unroll
            for (int px = 0; px < 4; ++px) acc[j][px] = b;
        }
        int yo[3];
        bool yok[3];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int yy = r + ky - 1;
            yok[ky] = yy >= 0 && yy < H;
            yo[ky] = min(max(yy, 0), H - 1) * W + x0;
        }
        const float* wp = wsm + og * OCT;
        for (int ci = 0; ci < cin; ++ci) {
            const float s = s_in[ci], sh = sh_in[ci];
            const float* ip = in + ci * HW;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                const float4 v = *reinterpret_cast<const float4*>(ip + yo[ky]);
                float a[6];
                a[1] = fmaf(v.x, s, sh);
                a[2] = fmaf(v.y, s, sh);
                a[3] = fmaf(v.z, s, sh);
                a[4] = fmaf(v.w, s, sh);
                if (RELU_IN) {
                    a[1] = fmaxf(a[1], 0.f);
                    a[2] = fmaxf(a[2], 0.f);
                    a[3] = fmaxf(a[3], 0.f);
                    a[4] = fmaxf(a[4], 0.f);
                }
                if (!yok[ky]) a[1] = a[2] = a[3] = a[4] = 0.f;
                a[0] = __shfl_up_sync(0xffffffffu, a[4], 1);
                a[5] = __shfl_down_sync(0xffffffffu, a[1], 1);
                if (ledge) a[0] = 0.f;
                if (redge) a[5] = 0.f;
                float2 aa[6];
#pragma unroll
                for (int i = 0; i < 6; ++i) aa[i] = make_float2(a[i], a[i]);
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const float* wq = wp + (ci * 9 + ky * 3 + kx) * cout;
                    float2 w2[OCT / 2];
                    if (OCT == 8) {
                        const float4 wa = *reinterpret_cast<const float4*>(wq);
                        const float4 wb = *reinterpret_cast<const float4*>(wq + 4);
                        w2[0] = make_float2(wa.x, wa.y);
                        w2[1] = make_float2(wa.z, wa.w);
                        w2[OCT / 2 - 2] = make_float2(wb.x, wb.y);
                        w2[OCT / 2 - 1] = make_float2(wb.z, wb.w);
                    } else if (OCT == 4) {
                        const float4 wa = *reinterpret_cast<const float4*>(wq);
                        w2[0] = make_float2(wa.x, wa.y);
                        w2[OCT / 2 - 1] = make_float2(wa.z, wa.w);
                    } else {
                        w2[0] = *reinterpret_cast<const float2*>(wq);
                    }
#pragma unroll
                    for (int j = 0; j < OCT / 2; ++j)
#pragma unroll
                        for (int px = 0; px < 4; ++px) acc[j][px] = __ffma2_rn(w2[j], aa[px + kx], acc[j][px]);
                }
            }
        }
        if (valid) {
            const int slot = (r - r0 + dst_r0) % dst_rmod;
#pragma unroll
            for (int j = 0; j < OCT / 2; ++j) {
                float4* d0 = reinterpret_cast<float4*>(dst + (og * OCT + 2 * j) * dst_cs + slot * W + x0);
                float4* d1 = reinterpret_cast<float4*>(dst + (og * OCT + 2 * j + 1) * dst_cs + slot * W + x0);
                float4 o0 = make_float4(acc[j][0].x, acc[j][1].x, acc[j][2].x, acc[j][3].x);
                float4 o1 = make_float4(acc[j][0].y, acc[j][1].y, acc[j][2].y, acc[j][3].y);
                if (ACCUM) {
                    const float4 p0 = *d0, p1 = *d1;
                    o0.x += p0.x; o0.y += p0.y; o0.z += p0.z; o0.w += p0.w;
                    o1.x += p1.x; o1.y += p1.y; o1.z += p1.z; o1.w += p1.w;
                }
                *d0 = o0;
                *d1 = o1;
            }
        }
    }
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// maxpool 3x3 stride 2 pad 1 (-inf padding) of conv rows [cr0 - 1, cr0 + 2*npy) held in `src` (row r of channel oc at
// src[oc*src_cs + (r % src_rmod)*W]) into pooled rows [py0, py0 + npy): a thread owns one (channel, pooled column), takes the
// 3-wide maximum of each conv row once and rolls the 3-row window down (27 loads for 4 outputs instead of 36, no per-output
// index arithmetic).  Wo = W/2 = 1 << lwo.
__device__ __forceinline__ void pool_rows(const float* __restrict__ src, int src_cs, int W, int src_rmod, int cr0, int npy, int py0,
                                          int cout, int lwo, float* __restrict__ dst_o, int dst_cs, int dst_rs) {
    const int Wo = 1 << lwo;
    for (int o = threadIdx.x; o < (cout << lwo); o += IM_THREADS) {
        const int oc = o >> lwo, px = o & (Wo - 1);
        const float* sb = src + oc * src_cs + 2 * px;
        auto hmax = [&](int r) {
            const float* row = sb + (r % src_rmod) * W;
            float m = fmaxf(row[0], row[1]);
            if (px > 0) m = fmaxf(m, row[-1]);
            return m;
        };
        float prev = cr0 > 0 ? hmax(cr0 - 1) : -INFINITY;
        float* d = dst_o + oc * dst_cs + py0 * dst_rs + px;
        for (int j = 0; j < npy; ++j) {
            const float a = hmax(cr0 + 2 * j), b = hmax(cr0 + 2 * j + 1);
            d[j * dst_rs] = fmaxf(fmaxf(prev, a), b);
            prev = b;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// Tensor-core convolution (precision >= 1): implicit GEMM on mma.sync m16n8k16, fp16 operands (the 10-bit mantissa of tf32;
// values saturate at +-65504), fp32 accumulate.
//   M = 16 neighbouring pixels, N = 8 output channels, K = 16 input channels at one filter tap (9 taps x cin/16 K-steps).
// The A operand is read straight from the activation map - no im2col copy: maps are stored PADDED ([c][H+2][W+2], channel
// stride = 8 mod 32 so the (pixel, channel) fragment loads are bank-conflict-free) with a NaN border, the input-side BN + ReLU
// is applied as the fragment is loaded (fmaxf(NaN, 0) = 0 turns the border into torch's zero padding of the BN output; the
// ReLU-less stage convolutions use a self-compare), and two channels are packed into one fp16x2 register.  The K index of the
// MMA is a free permutation of the channels as long as A and B agree: lane (g, tig) takes channels tig, tig + 4 (low / high
// half of a0, a1) and tig + 8, tig + 12 (a2, a3), so its four loads per pixel hit four different bank groups.  Weights are
// packed the same way when the layer is loaded (load_conv): one 8-byte load per n-tile fetches both B registers.
// Why mma.sync and not tcgen05 (yet): tcgen05's A operand must sit in shared memory in the UMMA core-matrix layout, i.e.
// one materialised im2col tile per filter tap (or channel-last maps addressed through shifted descriptors), while this loop
// feeds the fragments from the map as it is.  Measured on B200 (DFD_IMPALA_PROF timeline, 16 -> 16 layer at 32 x 32): packed
// FMAs 28 k cycles, tf32 m16n8k8 15 k, this loop 10 k - it is bound by instruction issue (fragment loads, BN, packing:
// ~11 instructions per MMA), not by the legacy HMMA pipe.
// A warp owns MT m-tiles x NT n-tiles; every layer is cut into exactly 16 such units (one per warp).
__device__ __forceinline__ void mma_f16(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
template <bool RELU_IN>
__device__ __forceinline__ float norm_in(float v, float s, float sh) {
    const float f = fmaf(v, s, sh);
    if (RELU_IN) return fmaxf(f, 0.f);      // NaN border -> 0
    return (f == f) ? f : 0.f;
}
__device__ __forceinline__ uint32_t pack_f16(float lo, float hi) {
    uint32_t u;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(u) : "f"(hi), "f"(lo));
    return u;
}

// in_o: padded input map at its (0, 0) element, channel stride in_cs, row stride in_rs.  Output rows [r0, r1) of the
// H x W map; element (oc, r, x) goes to dst[oc*dst_cs + ((r - r0 + dst_r0) % dst_rmod)*dst_rs + x]; `accum` adds (residual) -
// a run-time flag, so the two convolutions of a residual block run the SAME code (the second one finds it in the
// instruction cache; the kernel is ~200 KB of fully unrolled SASS).
// The map geometry (W, CIN, COUT) is a template parameter: every fragment load then carries its tap offset as an immediate
// and the loop is ~5 instructions per k8-equivalent MMA instead of ~10 (it is issue-bound, not MMA-bound).
__host__ __device__ constexpr int tc_rs(int W) { return W + 2; }
__host__ __device__ constexpr int tc_cs(int W) { return ((W + 2) * (W + 2) + 23) / 32 * 32 + 8; }     // = 8 mod 32
template <int MT, int NT, int W, int CIN, int COUT, bool RELU_IN>
__device__ __noinline__ void conv3x3_mma(const float* __restrict__ in_o, const float* __restrict__ wsm,
                                            const float* __restrict__ s_in, const float* __restrict__ sh_in,
                                            const float* __restrict__ bias, int r0, int r1,
                                            float* __restrict__ dst, int dst_cs, int dst_rs, int dst_r0, int dst_rmod, bool accum) {
    constexpr int cin = CIN, cout = COUT, in_cs = tc_cs(W), in_rs = tc_rs(W);
    const uint32_t* wpk = reinterpret_cast<const uint32_t*>(wsm);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, tig = lane & 3;
    constexpr int lw = W == 64 ? 6 : (W == 32 ? 5 : (W == 16 ? 4 : 3));
    static_assert((1 << lw) == W, "W must be 8, 16, 32 or 64");
    const int mblks = (((r1 - r0) << lw) >> 4) / MT;
    constexpr int nblks = (cout >> 3) / NT;
    const int swz = tig << 2;
    for (int unit = warp; unit < mblks * nblks; unit += IM_THREADS / 32) {
        const int mblk = unit % mblks, nblk = unit / mblks;
        int ao[MT][2], pr[MT][2], px[MT][2];
#pragma unroll
        for (int i = 0; i < MT; ++i)
#pragma unroll
            for (int p = 0; p < 2; ++p) {
                const int pm = ((mblk * MT + i) << 4) + g + 8 * p;
                pr[i][p] = r0 + (pm >> lw);
                px[i][p] = pm & (W - 1);
                ao[i][p] = pr[i][p] * in_rs + px[i][p];
            }
        float acc[MT][NT][4];
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const float2 b = *reinterpret_cast<const float2*>(bias + ((nblk * NT + j) << 3) + 2 * tig);
#pragma unroll
            for (int i = 0; i < MT; ++i) {
                acc[i][j][0] = b.x; acc[i][j][1] = b.y; acc[i][j][2] = b.x; acc[i][j][3] = b.y;
            }
        }
        int nph[NT];
#pragma unroll
        for (int j = 0; j < NT; ++j) nph[j] = ((((nblk * NT + j) << 3) + g) ^ swz);
        for (int cb = 0; cb < cin; cb += 16) {
            const float* ch[4];
            float sc[4], sh[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                ch[q] = in_o + (cb + tig + 4 * q) * in_cs;
                sc[q] = s_in[cb + tig + 4 * q];
                sh[q] = sh_in[cb + tig + 4 * q];
            }
            const uint32_t* wb = wpk + ((((cb >> 4) * 9) * 4 + tig) * cout << 1);
#pragma unroll 1   // filter-row loop rolled: a third of the code per layer shape
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const int tap = ky * 3 + kx;
                    const int toff = (ky - 1) * in_rs + (kx - 1);
                    uint32_t b0[NT], b1[NT];
#pragma unroll
                    for (int j = 0; j < NT; ++j) {
                        const uint2 bb = *reinterpret_cast<const uint2*>(wb + ((tap * 4 * cout + nph[j]) << 1));
                        b0[j] = bb.x;
                        b1[j] = bb.y;
                    }
#pragma unroll
                    for (int i = 0; i < MT; ++i) {
                        float v[4][2];
#pragma unroll
                        for (int q = 0; q < 4; ++q)
#pragma unroll
                            for (int p = 0; p < 2; ++p) v[q][p] = norm_in<RELU_IN>(ch[q][ao[i][p] + toff], sc[q], sh[q]);
                        const uint32_t a0 = pack_f16(v[0][0], v[1][0]), a1 = pack_f16(v[0][1], v[1][1]);
                        const uint32_t a2 = pack_f16(v[2][0], v[3][0]), a3 = pack_f16(v[2][1], v[3][1]);
#pragma unroll
                        for (int j = 0; j < NT; ++j) mma_f16(acc[i][j], a0, a1, a2, a3, b0[j], b1[j]);
                    }
                }
        }
#pragma unroll
        for (int i = 0; i < MT; ++i)
#pragma unroll
            for (int p = 0; p < 2; ++p) {
                const int slot = (pr[i][p] - r0 + dst_r0) % dst_rmod;
#pragma unroll
                for (int j = 0; j < NT; ++j)
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const int oc = ((nblk * NT + j) << 3) + 2 * tig + q;
                        float* d = dst + oc * dst_cs + slot * dst_rs + px[i][p];
                        if (accum) *d += acc[i][j][2 * p + q]; else *d = acc[i][j][2 * p + q];
                    }
            }
    }
}

// The first convolution (3 input channels, BN without ReLU) on the same MMA: its K = 27 (channel, tap) values are padded to
// 32 = two K-steps; lane (g, tig) owns the values kk = 16*step + 8j + tig (+ 4), whose map offsets it keeps in registers
// (padding values point at the pixel itself and meet zero weights).  8 conv rows x 64 pixels = 32 m-tiles, 2 per warp.
__device__ __forceinline__ void conv_first_mma(const float* __restrict__ in_o, const float* __restrict__ wsm,
                                               const float* __restrict__ s_in, const float* __restrict__ sh_in,
                                               const float* __restrict__ bias, int r0, float* __restrict__ dst, int dst_cs,
                                               int dst_rs, int dst_r0, int dst_rmod) {
    constexpr int MT = 2, NT = 2, W = 64, COUT = 16, in_cs = tc_cs(64), in_rs = tc_rs(64);
    const uint32_t* wpk = reinterpret_cast<const uint32_t*>(wsm);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, tig = lane & 3;
    const int swz = tig << 2;
    int koff[2][2][2];          // [step][j][half]
    float ksc[2][2][2], ksh[2][2][2];
#pragma unroll
    for (int st = 0; st < 2; ++st)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int kk = 16 * st + 8 * j + tig + 4 * h;
                const bool ok = kk < 27;
                const int ci = ok ? kk / 9 : 0, tap = ok ? kk - ci * 9 : 4;
                koff[st][j][h] = ci * in_cs + (tap / 3 - 1) * in_rs + (tap % 3 - 1);
                ksc[st][j][h] = s_in[ci];
                ksh[st][j][h] = sh_in[ci];
            }
    int ao[MT][2], pr[MT][2], px[MT][2];
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            const int pm = ((warp * MT + i) << 4) + g + 8 * p;
            pr[i][p] = r0 + (pm >> 6);
            px[i][p] = pm & (W - 1);
            ao[i][p] = pr[i][p] * in_rs + px[i][p];
        }
    float acc[MT][NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j) {
        const float2 b = *reinterpret_cast<const float2*>(bias + (j << 3) + 2 * tig);
#pragma unroll
        for (int i = 0; i < MT; ++i) {
            acc[i][j][0] = b.x; acc[i][j][1] = b.y; acc[i][j][2] = b.x; acc[i][j][3] = b.y;
        }
    }
#pragma unroll
    for (int st = 0; st < 2; ++st) {
        uint32_t b0[NT], b1[NT];
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const int nph = ((j << 3) + g) ^ swz;
            const uint2 bb = *reinterpret_cast<const uint2*>(wpk + (((st * 4 + tig) * COUT + nph) << 1));
            b0[j] = bb.x;
            b1[j] = bb.y;
        }
#pragma unroll
        for (int i = 0; i < MT; ++i) {
            float v[2][2][2];       // [j][half][pixel row]
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int p = 0; p < 2; ++p)
                        v[j][h][p] = norm_in<false>(in_o[ao[i][p] + koff[st][j][h]], ksc[st][j][h], ksh[st][j][h]);
            const uint32_t a0 = pack_f16(v[0][0][0], v[0][1][0]), a1 = pack_f16(v[0][0][1], v[0][1][1]);
            const uint32_t a2 = pack_f16(v[1][0][0], v[1][1][0]), a3 = pack_f16(v[1][0][1], v[1][1][1]);
#pragma unroll
            for (int j = 0; j < NT; ++j) mma_f16(acc[i][j], a0, a1, a2, a3, b0[j], b1[j]);
        }
    }
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            const int slot = (pr[i][p] - r0 + dst_r0) % dst_rmod;
#pragma unroll
            for (int j = 0; j < NT; ++j)
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int oc = (j << 3) + 2 * tig + q;
                    dst[oc * dst_cs + slot * dst_rs + px[i][p]] = acc[i][j][2 * p + q];
                }
        }
}

// ---------------------------------------------------------------------------------------------------------------------------
// Dense tail for the NM members of this CTA (NM = 2: the two members of an antithetic pair, whose eps row is streamed ONCE).
// fc[q]: BN'd flattened trunk output of member q (2048); st: per-member scratch (core 260 | h0 256 | gates 1024 | hn 256 |
// logits 32 = ST floats).
template <int NM>
__device__ __forceinline__ void dense_tail(const ImpalaP& L, const float* __restrict__ theta, const float* __restrict__ row,
                                           const float* __restrict__ bnbuf, const float (&sg)[2], const int (&inst)[2],
                                           float* const (&fc)[2], float* st, const float* __restrict__ reward,
                                           const uint8_t* __restrict__ done, const float* __restrict__ h_in,
                                           const float* __restrict__ c_in, float* __restrict__ probs,
                                           float* __restrict__ h_out, float* __restrict__ c_out) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    auto par = [&](int q, int64_t p) { return perturb1(theta[p], sg[q], row[p]); };
    bool dn[NM];
#pragma unroll
    for (int q = 0; q < NM; ++q) {
        dn[q] = done[inst[q]] != 0;
        float* hst = st + q * ST + 260;
        for (int k = tid; k < 256; k += IM_THREADS) hst[k] = dn[q] ? 0.f : h_in[(int64_t)inst[q] * 256 + k];
    }
    __syncthreads();
    // Linear 2048 -> 256 (+ReLU): warp per output row, 8-byte loads (row starts are 8-byte aligned only), 16 + 16 loads
    // (256 B) in flight per lane
    for (int o = warp; o < 256; o += IM_THREADS / 32) {
        const int64_t base = L.fc_w + (int64_t)o * 2048;
        float acc[NM];
#pragma unroll
        for (int q = 0; q < NM; ++q) acc[q] = 0.f;
        for (int v0 = 0; v0 < 1024; v0 += 512) {
            float2 tw[16], ew[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                const int v = v0 + u * 32 + lane;
                tw[u] = *reinterpret_cast<const float2*>(theta + base + 2 * v);
                ew[u] = *reinterpret_cast<const float2*>(row + base + 2 * v);
            }
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                const int v = v0 + u * 32 + lane;
#pragma unroll
                for (int q = 0; q < NM; ++q) {
                    const float2 f = *reinterpret_cast<const float2*>(fc[q] + 2 * v);
                    acc[q] = fmaf(perturb1(tw[u].x, sg[q], ew[u].x), f.x, acc[q]);
                    acc[q] = fmaf(perturb1(tw[u].y, sg[q], ew[u].y), f.y, acc[q]);
                }
            }
        }
#pragma unroll
        for (int q = 0; q < NM; ++q) {
            const float a = warp_sum(acc[q]);
            if (lane == 0) st[q * ST + o] = fmaxf(a + par(q, L.fc_b + o), 0.f);
        }
    }
    if (tid < NM) st[tid * ST + 256] = fminf(fmaxf(reward[inst[tid]], -1.f), 1.f);   // clamp(reward, -1, 1), impala.py:158
    __syncthreads();
    // LSTM gates = W_ih [x;r] + b_ih + W_hh h0 + b_hh   (rows: i | f | g | o, 256 each).  A warp takes two gate rows at a
    // time and issues all their loads first: W_ih rows (257 floats, 4-byte aligned only) lane-strided, W_hh rows (256
    // floats, 8-byte aligned) as float2 - 52 loads = 272 B in flight per lane
    for (int o0 = 2 * warp; o0 < 1024; o0 += 2 * (IM_THREADS / 32)) {
        float ti[2][9], ei[2][9];
        float2 th[2][4], eh[2][4];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int64_t bi = L.wih + (int64_t)(o0 + r) * 257, bh = L.whh + (int64_t)(o0 + r) * 256;
#pragma unroll
            for (int u = 0; u < 9; ++u) {
                const int k = lane + 32 * u;
                ti[r][u] = 0.f;
                ei[r][u] = 0.f;
                if (k < 257) {
                    ti[r][u] = theta[bi + k];
                    ei[r][u] = row[bi + k];
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int v = lane + 32 * u;
                th[r][u] = *reinterpret_cast<const float2*>(theta + bh + 2 * v);
                eh[r][u] = *reinterpret_cast<const float2*>(row + bh + 2 * v);
            }
        }
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int q = 0; q < NM; ++q) {
                const float* core = st + q * ST;
                const float* hst = core + 260;
                float acc = 0.f;
#pragma unroll
                for (int u = 0; u < 9; ++u) {
                    const int k = lane + 32 * u;
                    if (k < 257) acc = fmaf(perturb1(ti[r][u], sg[q], ei[r][u]), core[k], acc);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int v = lane + 32 * u;
                    const float2 hv = *reinterpret_cast<const float2*>(hst + 2 * v);
                    acc = fmaf(perturb1(th[r][u].x, sg[q], eh[r][u].x), hv.x, acc);
                    acc = fmaf(perturb1(th[r][u].y, sg[q], eh[r][u].y), hv.y, acc);
                }
                acc = warp_sum(acc);
                if (lane == 0) st[q * ST + 516 + o0 + r] = acc + par(q, L.bih + o0 + r) + par(q, L.bhh + o0 + r);
            }
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < NM; ++q) {
        const float* gates = st + q * ST + 516;
        float* hn = st + q * ST + 1540;
        for (int k = tid; k < 256; k += IM_THREADS) {
            const float c0 = dn[q] ? 0.f : c_in[(int64_t)inst[q] * 256 + k];
            const float ig = sigmoidf_(gates[k]), fg = sigmoidf_(gates[256 + k]);
            const float gg = tanhf(gates[512 + k]), og = sigmoidf_(gates[768 + k]);
            const float c1 = fg * c0 + ig * gg;
            const float h1 = og * tanhf(c1);
            c_out[(int64_t)inst[q] * 256 + k] = c1;
            h_out[(int64_t)inst[q] * 256 + k] = h1;
            const float inv = 1.0f / sqrtf(bnbuf[L.pol_bv + k] + 1e-5f);
            const float s = par(q, L.pol_g + k) * inv;
            hn[k] = fmaf(h1, s, par(q, L.pol_be + k) - bnbuf[L.pol_bm + k] * s);
        }
    }
    __syncthreads();
    for (int a = warp; a < NM * L.A; a += IM_THREADS / 32) {
        const int q = a / L.A, ai = a - q * L.A;
        const float* hn = st + q * ST + 1540;
        float acc = 0.f;
        for (int k = lane; k < 256; k += 32) acc = fmaf(par(q, L.pol_w + ai * 256 + k), hn[k], acc);
        acc = warp_sum(acc);
        if (lane == 0) st[q * ST + 1796 + ai] = acc + par(q, L.pol_b + ai);
    }
    __syncthreads();
    if (tid < NM) {
        const float* lg = st + tid * ST + 1796;
        float mx = -INFINITY;
        for (int a = 0; a < L.A; ++a) mx = fmaxf(mx, lg[a]);
        float ssum = 0.f;
        for (int a = 0; a < L.A; ++a) ssum += expf(lg[a] - mx);
        const float inv = 1.0f / ssum;
        for (int a = 0; a < L.A; ++a) probs[(int64_t)inst[tid] * L.A + a] = expf(lg[a] - mx) * inv;
    }
}

// TC = false: exact fp32 (unpadded maps).  TC = true: tensor-core convolutions (padded maps with a NaN border); the dense
// tail stays fp32.
// pair_order: 0 = CTA per (member, env) in plain order; 1 = the same with the two members of an antithetic pair on
// neighbouring CTAs; 2 = CTA per (pair, env): members j and j + M/2 run their trunks one after the other and share the dense
// tail, so theta and the pair's eps row (8.4 MB per pair) are streamed once for both.
template <bool TC>
__global__ void __launch_bounds__(IM_THREADS, 1) impala_forward_kernel(ImpalaP L, const float* __restrict__ replicas,
                                                                       int64_t stride, const float* __restrict__ theta,
                                                                       const float* __restrict__ bnbuf,
                                                                       const int64_t* __restrict__ idx,
                                                                       const int8_t* __restrict__ sign, float sigma,
                                                                       const float* __restrict__ frame,
                                                                       const float* __restrict__ reward,
                                                                       const uint8_t* __restrict__ done,
                                                                       const float* __restrict__ h_in,
                                                                       const float* __restrict__ c_in, int E,
                                                                       float* __restrict__ probs, float* __restrict__ h_out,
                                                                       float* __restrict__ c_out, int n_members, int pair_order,
                                                                       long long* __restrict__ prof,
                                                                       const __grid_constant__ ItMaps maps, int tail_tc) {
    constexpr int MAPF = TC ? MAP_TC : MAP;
    extern __shared__ __align__(16) float sm[];
    __shared__ __align__(8) uint64_t tbars[TB_COUNT];      // dense tail on tcgen05 / TMA (impala_tail.cuh), tail_tc != 0
    __shared__ uint32_t tmem_base_s;
    float* bufA = sm;
    float* bufB = bufA + MAPF;
    float* band = bufB + MAPF;
    float* wsm = band + BAND;
    float* s_in = wsm + WMAX;      // 32
    float* sh_in = s_in + 32;      // 32
    float* bias = sh_in + 32;      // 32
    float* fc0 = bias + 32;        // 2048: BN'd trunk output of the CTA's first member (the second one's goes to `band`)

    const int tid = threadIdx.x;
    if (tail_tc) {
        if (tid == 0) {
            for (int i = 0; i < TB_COUNT; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&tbars[i])));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        if ((tid >> 5) == 1) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    int stamp_i = 0;
    // DFD_IMPALA_PROF=1: cycle stamps of CTA 7 at the phase boundaries (all stamps follow a CTA barrier)
    auto stamp = [&]() {
        if (prof != nullptr && blockIdx.x == 7 && tid == 0) prof[stamp_i] = clock64();
        ++stamp_i;
    };
    stamp();
    const int nmem = pair_order == 2 ? 2 : 1;
    const int mb = blockIdx.x / E, env = blockIdx.x - mb * E;
    const int m0 = pair_order == 1 ? ((mb & 1) ? (n_members >> 1) + (mb >> 1) : (mb >> 1)) : mb;
    const int ms[2] = {m0, pair_order == 2 ? m0 + (n_members >> 1) : m0};
    const int inst[2] = {ms[0] * E + env, ms[1] * E + env};          // (member, env)
    const float sgs[2] = {sigma * (float)sign[ms[0]], sigma * (float)sign[ms[1]]};
    const float* rows[2] = {table_row_ptr(replicas, stride, idx[ms[0]]), table_row_ptr(replicas, stride, idx[ms[1]])};
    float* const fcs[2] = {fc0, band};

    for (int mem = 0; mem < nmem; ++mem) {
        Ctx c;
        c.theta = theta;
        c.bn = bnbuf;
        c.sg = sgs[mem];
        c.row = rows[mem];
        __syncthreads();        // the previous member's trunk output has been read
        // frame / 255 -> bufA  (impala.py:142); 24 loads per thread in three rounds of 8.  Tensor-core path: padded
        // [3][66][66] (channel stride 4360 = 8 mod 32) inside a NaN border
        constexpr int FCS = TC ? tc_cs(64) : 4096, FRS = TC ? tc_rs(64) : 64, FORG = TC ? tc_rs(64) + 1 : 0;
        if (TC) {
            for (int i = tid; i < 3 * FCS; i += IM_THREADS) bufA[i] = __int_as_float(0x7fc00000);
            __syncthreads();
        }
        const float* fr = frame + (int64_t)inst[mem] * 12288;
        for (int t0 = tid; t0 < 12288; t0 += IM_THREADS * 8) {
            float f[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) f[u] = fr[t0 + u * IM_THREADS];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int e = t0 + u * IM_THREADS;
                bufA[FORG + (e >> 12) * FCS + ((e >> 6) & 63) * FRS + (e & 63)] = f[u] / 255.0f;
            }
        }
        __syncthreads();
        if (mem == 0) stamp();            // 1: frame in shared memory
        float* x = bufA;    // current map
        float* t = bufB;    // the other buffer
        int H = 64;
        int xcs = FCS, xrs = FRS, xorg = FORG;     // geometry of x: channel stride, row stride, offset of element (0, 0)
        int li = 0;         // position in the execution-order list of conv layers
        auto next_layer = [&](const ConvP& p, bool tc_layer) {
            load_conv(c, p, wsm, band, s_in, sh_in, bias, tc_layer);
            ++li;
            if (li < 15) prefetch_l2(c.row + L.seq_w[li], L.seq_n[li]);
            else prefetch_l2(c.row + L.fc_w, 65536);      // first 32 rows of the Linear
        };
        auto nan_fill = [&](float* buf) {
            for (int i = tid; i < MAPF; i += IM_THREADS) buf[i] = __int_as_float(0x7fc00000);
        };
        for (int s = 0; s < 3; ++s) {
            const ConvP& fp = L.feat[s];
            const bool tc_conv = TC;
            __syncthreads();
            next_layer(fp, tc_conv);
            // conv (BN on the input, no ReLU) + maxpool 3x3 stride 2 pad 1 (-inf padding)
            const int W = H, Ho = H / 2, Wo = W / 2;
            const int lwo = 31 - __clz(Wo);
            // geometry of the pooled map
            const int trs = TC ? Wo + 2 : Wo;
            const int tcs = TC ? (((Wo + 2) * (Wo + 2) + 23) / 32 * 32 + 8) : Wo * Wo;
            const int torg = TC ? trs + 1 : 0;
            if (TC) nan_fill(t);        // t is free: nobody reads it until the pool below has written the new map
            __syncthreads();
            if (s < 2) {
                // bands of 4 pooled rows: 8 NEW conv rows per band go into a 9-row circular band (conv row r lives in slot
                // r % 9), the row above them is still there from the previous band
                for (int py0 = 0; py0 < Ho; py0 += 4) {
                    const int cr0 = 2 * py0;
                    if (tc_conv && s == 0)
                        conv_first_mma(x + xorg, wsm, s_in, sh_in, bias, cr0, band, 9 * W, W, cr0 % 9, 9);
                    else if (tc_conv)
                        conv3x3_mma<1, 4, 32, 16, 32, false>(x + xorg, wsm, s_in, sh_in, bias, cr0, cr0 + 8, band, 9 * W, W, cr0 % 9, 9, false);
                    else
                        conv3x3<4, false, false>(x, fp.cin, H, W, wsm, fp.cout, s_in, sh_in, bias, cr0, cr0 + 8, band, 9 * W, cr0 % 9, 9);
                    __syncthreads();
                    pool_rows(band, 9 * W, W, 9, cr0, 4, py0, fp.cout, lwo, t + torg, tcs, trs);
                    __syncthreads();
                }
            } else {
                // 16 x 16: the whole conv output (32 x 16 x 16) fits the band buffer
                if (tc_conv)
                    conv3x3_mma<1, 4, 16, 32, 32, false>(x + xorg, wsm, s_in, sh_in, bias, 0, H, band, H * W, W, 0, H, false);
                else
                    conv3x3<4, false, false>(x, fp.cin, H, W, wsm, fp.cout, s_in, sh_in, bias, 0, H, band, H * W, 0, H);
                __syncthreads();
                pool_rows(band, H * W, W, H, 0, Ho, 0, fp.cout, lwo, t + torg, tcs, trs);
                __syncthreads();
            }
            { float* tmp = x; x = t; t = tmp; }
            H = Ho;
            xcs = tcs; xrs = trs; xorg = torg;
            if (TC) nan_fill(t);        // the old input map becomes the block-internal map: NaN border in the new geometry
            if (mem == 0) stamp();        // 2, 5, 8: stage conv + pool done
            // two residual blocks at this resolution: x += conv_b(relu(BN_b(conv_a(relu(BN_a(x))))))
            for (int blk = 0; blk < 2; ++blk) {
                const ConvP& pa = L.res[blk][s][0];
                const ConvP& pb = L.res[blk][s][1];
                auto fine = [&](int i) {       // finer stamps inside the first residual block of each stage
                    if (prof != nullptr && blockIdx.x == 7 && tid == 0 && mem == 0 && blk == 0) prof[16 + 4 * s + i] = clock64();
                };
                next_layer(pa, TC);
                __syncthreads();
                fine(0);
                if (TC) {
                    if (s == 0) conv3x3_mma<4, 2, 32, 16, 16, true>(x + xorg, wsm, s_in, sh_in, bias, 0, H, t + xorg, xcs, xrs, 0, H, false);
                    else if (s == 1) conv3x3_mma<1, 4, 16, 32, 32, true>(x + xorg, wsm, s_in, sh_in, bias, 0, H, t + xorg, xcs, xrs, 0, H, false);
                    else conv3x3_mma<1, 1, 8, 32, 32, true>(x + xorg, wsm, s_in, sh_in, bias, 0, H, t + xorg, xcs, xrs, 0, H, false);
                } else {
                    if (s == 0) conv3x3<8, true, false>(x, pa.cin, H, H, wsm, pa.cout, s_in, sh_in, bias, 0, H, t, H * H, 0, H);
                    else if (s == 1) conv3x3<4, true, false>(x, pa.cin, H, H, wsm, pa.cout, s_in, sh_in, bias, 0, H, t, H * H, 0, H);
                    else conv3x3<2, true, false>(x, pa.cin, H, H, wsm, pa.cout, s_in, sh_in, bias, 0, H, t, H * H, 0, H);
                }
                __syncthreads();
                fine(1);
                next_layer(pb, TC);
                __syncthreads();
                fine(2);
                if (TC) {
                    if (s == 0) conv3x3_mma<4, 2, 32, 16, 16, true>(t + xorg, wsm, s_in, sh_in, bias, 0, H, x + xorg, xcs, xrs, 0, H, true);
                    else if (s == 1) conv3x3_mma<1, 4, 16, 32, 32, true>(t + xorg, wsm, s_in, sh_in, bias, 0, H, x + xorg, xcs, xrs, 0, H, true);
                    else conv3x3_mma<1, 1, 8, 32, 32, true>(t + xorg, wsm, s_in, sh_in, bias, 0, H, x + xorg, xcs, xrs, 0, H, true);
                } else {
                    if (s == 0) conv3x3<8, true, true>(t, pb.cin, H, H, wsm, pb.cout, s_in, sh_in, bias, 0, H, x, H * H, 0, H);
                    else if (s == 1) conv3x3<4, true, true>(t, pb.cin, H, H, wsm, pb.cout, s_in, sh_in, bias, 0, H, x, H * H, 0, H);
                    else conv3x3<2, true, true>(t, pb.cin, H, H, wsm, pb.cout, s_in, sh_in, bias, 0, H, x, H * H, 0, H);
                }
                __syncthreads();
                fine(3);
                if (mem == 0) stamp();    // residual block done
            }
        }
        // x: [32][8][8].  relu -> flatten (C,H,W) -> BN1d(2048)
        float* fcin = fcs[mem];
        for (int k = tid; k < 2048; k += IM_THREADS) {
            const float inv = 1.0f / sqrtf(bnbuf[L.fc_bv + k] + 1e-5f);
            const float s = c.par(L.fc_g + k) * inv;
            const float xv = x[xorg + (k >> 6) * xcs + ((k >> 3) & 7) * xrs + (k & 7)];
            fcin[k] = fmaf(fmaxf(xv, 0.f), s, c.par(L.fc_be + k) - bnbuf[L.fc_bm + k] * s);
        }
    }
    __syncthreads();            // trunks done: the map buffers are dead and hold the dense tail's scratch from here on
    stamp_i = 11;
    stamp();                    // 11: trunks done
    if (tail_tc) {
        // TMA-fed tcgen05 dense tail: warps 0-14 are the workers, lane 0 of warp 15 streams the weight tiles
        const uint32_t sraw = smem_u32(sm), s0 = (sraw + 1023u) & ~1023u;
        uint8_t* smb = reinterpret_cast<uint8_t*>(sm) + (s0 - sraw);
        constexpr int XT_OFF = TL_HT + 5120, FCIN_OFF = XT_OFF + TL_XT_BYTES;
        __half* fcin = reinterpret_cast<__half*>(smb + FCIN_OFF);
        for (int i = tid; i < nmem * 2048; i += IM_THREADS) fcin[i] = __float2half_rn(fcs[i >> 11][i & 2047]);
        __syncthreads();
        TailArgs ta;
        ta.theta = theta; ta.bnbuf = bnbuf; ta.reward = reward; ta.done = done; ta.h_in = h_in; ta.c_in = c_in;
        ta.probs = probs; ta.h_out = h_out; ta.c_out = c_out; ta.sigma = sigma; ta.nmem = nmem;
        ta.nE = (nmem == 2 && rows[0] != rows[1]) ? 2 : 1;
        for (int i = 0; i < 2; ++i) { ta.inst[i] = inst[i]; ta.sgi[i] = (int)sign[ms[i]]; ta.ids[i] = idx[ms[i]]; ta.rows[i] = rows[i]; }
        const uint32_t tb0 = smem_u32(&tbars[0]);
        if (tid < 480) tl_dense_tail_workers<480, XT_OFF>(L, ta, smb, s0, fcin, tb0, tmem_base_s, 0u);
        else if (tid == 480) tl_dense_tail_producer(L, maps, ta, s0, tb0);
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if ((tid >> 5) == 1) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base_s), "r"(512u) : "memory");
        }
    } else if (nmem == 2 && rows[0] == rows[1]) {
        dense_tail<2>(L, theta, rows[0], bnbuf, sgs, inst, fcs, bufA, reward, done, h_in, c_in, probs, h_out, c_out);
    } else {
        for (int mem = 0; mem < nmem; ++mem) {
            const float sg1[2] = {sgs[mem], sgs[mem]};
            const int in1[2] = {inst[mem], inst[mem]};
            float* const fc1[2] = {fcs[mem], fcs[mem]};
            dense_tail<1>(L, theta, rows[mem], bnbuf, sg1, in1, fc1, bufA, reward, done, h_in, c_in, probs, h_out, c_out);
            __syncthreads();
        }
    }
    stamp();                    // 12: end
}

}  // namespace

int dfd_impala_forward_direct_impl(dfd_ctx* ctx, const dfd_policy_desc* desc, const dfd_table* table, const float* theta,
                                   const float* bn_buffers, const int64_t* idx, const int8_t* sign, int n_members, float sigma,
                                   const float* frame, const float* reward, const uint8_t* done, const float* h_in,
                                   const float* c_in, int obs_per_member, float* probs, float* h_out, float* c_out,
                                   cudaStream_t st);

extern "C" size_t dfd_impala_scratch_bytes(int n_members, int obs_per_member) {
    (void)n_members;
    (void)obs_per_member;
    return 256;   // the trunk keeps its activations in shared memory; no global scratch is needed
}

extern "C" int dfd_impala_forward(dfd_ctx* ctx, const dfd_policy_desc* desc, const dfd_table* table, const float* theta,
                                  const float* bn_buffers, const int64_t* idx, const int8_t* sign, int n_members,
                                  float sigma, const float* frame, const float* reward, const uint8_t* done,
                                  const float* h_in, const float* c_in, int obs_per_member, float* probs, float* h_out,
                                  float* c_out, void* scratch, size_t scratch_bytes, dfd_stream stream) {
    (void)scratch;
    (void)scratch_bytes;
    DFD_CHECK_ARG(ctx && desc && table && theta && bn_buffers && idx && sign && frame && reward && done && h_in && c_in &&
                      probs && h_out && c_out, "dfd_impala_forward: NULL argument");
    DFD_CHECK_ARG(desc->kind == DFD_POLICY_IMPALA, "dfd_impala_forward: desc.kind must be DFD_POLICY_IMPALA");
    DFD_CHECK_ARG(desc->n_act >= 1 && desc->n_act <= 32, "dfd_impala_forward: n_act %d out of range (1..32)", desc->n_act);
    if (n_members == 0 || obs_per_member == 0) return 0;
    DFD_CHECK_ARG(n_members > 0 && obs_per_member > 0, "dfd_impala_forward: negative sizes");
    const ImpalaP L = make_impala(desc->n_act);
    DFD_CHECK_ARG(L.P == dfd_policy_num_params(desc) && L.P < table->size, "dfd_impala_forward: parameter count mismatch");
    DFD_CHECK_ARG((((uintptr_t)theta) & 15) == 0, "dfd_impala_forward: theta must be 16-byte aligned");
    DFD_CHECK_ARG((int64_t)n_members * obs_per_member < 2147483647LL, "dfd_impala_forward: grid too large");
    if (desc->precision >= 3) {       // tcgen05 trunk + TMA-fed dense tail (csrc/impala_forward_tc.cu) when the scaled table mirror is registered
        const int rd = dfd_impala_forward_direct_impl(ctx, desc, table, theta, bn_buffers, idx, sign, n_members, sigma, frame, reward,
                                                      done, h_in, c_in, obs_per_member, probs, h_out, c_out, (cudaStream_t)stream);
        if (rd >= 0) return rd;
    }
    const bool tc = desc->precision >= 1;
    const size_t smem = (size_t)(2 * (tc ? MAP_TC : MAP) + BAND + WMAX + 96 + 2048) * sizeof(float);
    // 2: one CTA per (antithetic pair, env) - members j and j + M/2 of [plus | minus] batches; 1: CTA per (member, env),
    // pair-adjacent order; 0: plain order.  DFD_IMPALA_NO_PAIR=1 keeps mode 1 (A/B measurements)
    static const bool no_pair = getenv("DFD_IMPALA_NO_PAIR") != nullptr;
    const int mode = (n_members % 2 == 0) ? (no_pair ? 1 : 2) : 0;
    const int grid = (mode == 2 ? n_members / 2 : n_members) * obs_per_member;
    long long* prof = nullptr;
    if (getenv("DFD_IMPALA_PROF")) { cudaMalloc(&prof, 32 * 8); cudaMemset(prof, 0, 32 * 8); }   // 0..12 phases, 16..27 fine stamps
    // level 2: the dense tail (90 % of the parameters) as TMA-fed tcgen05 GEMMs (impala_tail.cuh) when the sigma-scaled
    // fp16 mirror of this table is registered with the context
    ItMaps maps;
    memset(&maps, 0, sizeof(maps));
    int tail_tc = 0;
    if (desc->precision >= 2 && !getenv("DFD_TC_NO_DIRECT") && ctx->scaled16 && ctx->scaled_src == table->replicas &&
        ctx->scaled_sigma == sigma && ctx->theta16_cap >= 1048576) {
        if (tl_prepare(ctx, L, theta, &maps, (cudaStream_t)stream)) return 3;
        tail_tc = 1;
    }
    if (tc) {
        DFD_CUDA(cudaFuncSetAttribute(impala_forward_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        impala_forward_kernel<true><<<grid, IM_THREADS, smem, (cudaStream_t)stream>>>(
            L, table->replicas, table->replica_stride, theta, bn_buffers, idx, sign, sigma, frame, reward, done, h_in, c_in,
            obs_per_member, probs, h_out, c_out, n_members, mode, prof, maps, tail_tc);
    } else {
        DFD_CUDA(cudaFuncSetAttribute(impala_forward_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        impala_forward_kernel<false><<<grid, IM_THREADS, smem, (cudaStream_t)stream>>>(
            L, table->replicas, table->replica_stride, theta, bn_buffers, idx, sign, sigma, frame, reward, done, h_in, c_in,
            obs_per_member, probs, h_out, c_out, n_members, mode, prof, maps, tail_tc);
    }
    DFD_LAUNCHED(ctx);
    if (prof) {
        cudaStreamSynchronize((cudaStream_t)stream);
        long long h[32];
        cudaMemcpy(h, prof, sizeof(h), cudaMemcpyDeviceToHost);
        fprintf(stderr, "[impala timeline] CTA 7 (mode %d, %s), cycles per phase of its first member: frame %lld | s0 conv+pool %lld res %lld %lld | "
                        "s1 conv+pool %lld res %lld %lld | s2 conv+pool %lld res %lld %lld | all trunks done at %lld | dense tail %lld | total %lld\n",
                mode, tc ? "fp16 mma" : "fp32", h[1] - h[0], h[2] - h[1], h[3] - h[2], h[4] - h[3], h[5] - h[4], h[6] - h[5], h[7] - h[6],
                h[8] - h[7], h[9] - h[8], h[10] - h[9], h[11] - h[0], h[12] - h[11], h[12] - h[0]);
        for (int s = 0; s < 3; ++s)
            fprintf(stderr, "[impala timeline] stage %d first residual block: weights a %lld | conv a %lld | weights b %lld | conv b %lld\n", s,
                    h[16 + 4 * s] - h[2 + 3 * s], h[17 + 4 * s] - h[16 + 4 * s], h[18 + 4 * s] - h[17 + 4 * s], h[19 + 4 * s] - h[18 + 4 * s]);
        cudaFree(prof);
    }
    return 0;
}
