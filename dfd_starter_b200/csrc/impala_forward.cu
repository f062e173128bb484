// IMPALA CNN + LSTM perturbed forward (policies/impala.py:136-186), one CTA per (member, environment).
//   x = frame/255; 3 stages {BN -> conv3x3 -> maxpool(3,2,1); 2 x [x += conv(relu(BN(conv(relu(BN(x))))))]};
//   relu -> flatten(2048) -> relu(Linear(BN1d(x))) -> concat clamp(reward,-1,1) -> LSTM cell (gates i,f,g,o,
//   state zeroed where done) -> Linear(BN1d(h)) -> softmax.
// BN layers are eval-mode with shared running statistics and per-member (perturbed) gamma / beta; every
// parameter is theta + sign*sigma*eps generated in-kernel (worker/worker.py:28).  The whole conv trunk keeps its
// activations in shared memory (two ping-pong maps + a conv-row band for the pooled stages); only the
// carried LSTM state and the action probabilities touch HBM.  Exact fp32 on CUDA cores (atol 1e-5 against torch CPU):
//   * convolutions are register-tiled: a thread owns 4 neighbouring pixels x OCT output channels (OCT = 8 / 4 / 2 chosen so
//     that every layer fills the 512 threads), reads its 4 inputs with one 16-byte shared-memory load per (channel, row),
//     applies the input-side BN (+ReLU) once, takes the two halo pixels from the neighbouring lanes by shuffle, and issues
//     packed fp32 FMAs (FFMA2: two output channels per instruction, the weight pair comes straight out of the 16-byte
//     weight load);
//   * the pooled stages compute exactly 8 new conv rows per band into a 9-row circular band (no row is convolved twice);
//   * a layer's weights arrive in ONE round of loads (every thread's theta / eps loads issued before the first store) and
//     the NEXT layer's eps segment is prefetched into L2 while the current layer computes;
//   * the dense tail (Linear 2048 -> 256, LSTM 513 -> 1024) streams 8.4 MB per member: every warp keeps 256 B (Linear) /
//     272 B (LSTM, two gate rows at a time) of loads in flight per lane.
#include "common.cuh"
#include <stdlib.h>

namespace {

constexpr int IM_THREADS = 512;
constexpr int MAP = 16384;    // floats per activation map buffer (16 x 32 x 32)
constexpr int BAND = 9248;    // conv-row band for the pooled stages: 9 rows x (64 x 16 | 32 x 32), 32 x 16 x 16; staging of a
                              // layer's weights (32 rows x 289)
constexpr int WMAX = 9216;    // largest conv weight block (32 x 32 x 3 x 3)

struct ConvP { int g, be, w, b, bm, bv, cin, cout; };
struct ImpalaP {
    ConvP feat[3];
    ConvP res[2][3][2];
    int fc_g, fc_be, fc_w, fc_b, fc_bm, fc_bv;
    int wih, whh, bih, bhh;
    int pol_g, pol_be, pol_w, pol_b, pol_bm, pol_bv;
    int A;
    int64_t P;
    int seq_w[16], seq_n[16];     // conv weight segments in execution order (L2 prefetch of the next layer)
};

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
__device__ void load_conv(const Ctx& c, const ConvP& p, float* wsm, float* stage, float* s_in, float* sh_in, float* bias) {
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
    for (int t = tid; t < n; t += IM_THREADS) {
        const int k = t >> lc, oc = t & (p.cout - 1);
        wsm[t] = stage[oc * ss + k];
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
                                                                       long long* __restrict__ prof) {
    extern __shared__ __align__(16) float sm[];
    float* bufA = sm;
    float* bufB = bufA + MAP;
    float* band = bufB + MAP;
    float* wsm = band + BAND;
    float* s_in = wsm + WMAX;      // 32
    float* sh_in = s_in + 32;      // 32
    float* bias = sh_in + 32;      // 32
    float* vec = bias + 32;        // 2048 + 257 + 256 + 1024 + 32 scratch for the dense tail

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int stamp_i = 0;
    // DFD_IMPALA_PROF=1: cycle stamps of CTA 7 at the phase boundaries (all stamps follow a CTA barrier)
    auto stamp = [&]() {
        if (prof != nullptr && blockIdx.x == 7 && tid == 0) prof[stamp_i] = clock64();
        ++stamp_i;
    };
    stamp();
    // consecutive CTAs take the two members of an antithetic pair ([plus | minus] batches: members j and j + M/2 share
    // their table row), so the pair streams the same eps row at the same time and HBM serves it once
    const int mb = blockIdx.x / E, env = blockIdx.x - mb * E;
    const int m = pair_order ? ((mb & 1) ? (n_members >> 1) + (mb >> 1) : (mb >> 1)) : mb;
    const int inst = m * E + env;           // (member, env)
    Ctx c;
    c.theta = theta;
    c.bn = bnbuf;
    c.sg = sigma * (float)sign[m];
    c.row = table_row_ptr(replicas, stride, idx[m]);

    // frame / 255 -> bufA  (impala.py:142); 24 loads per thread in three rounds of 8
    const float* fr = frame + (int64_t)inst * 12288;
    for (int t0 = tid; t0 < 12288; t0 += IM_THREADS * 8) {
        float f[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) f[u] = fr[t0 + u * IM_THREADS];
#pragma unroll
        for (int u = 0; u < 8; ++u) bufA[t0 + u * IM_THREADS] = f[u] / 255.0f;
    }

    __syncthreads();
    stamp();            // 1: frame in shared memory
    float* x = bufA;    // current map
    float* t = bufB;    // the other buffer
    int H = 64;
    int li = 0;         // position in the execution-order list of conv layers
    auto next_layer = [&](const ConvP& p) {
        load_conv(c, p, wsm, band, s_in, sh_in, bias);
        ++li;
        if (li < 15) prefetch_l2(c.row + L.seq_w[li], L.seq_n[li]);
        else prefetch_l2(c.row + L.fc_w, 65536);      // first 32 rows of the Linear
    };
    for (int s = 0; s < 3; ++s) {
        const ConvP& fp = L.feat[s];
        __syncthreads();
        next_layer(fp);
        __syncthreads();
        // conv (BN on the input, no ReLU) + maxpool 3x3 stride 2 pad 1 (-inf padding)
        const int W = H, Ho = H / 2, Wo = W / 2;
        const int lwo = 31 - __clz(Wo);
        if (s < 2) {
            // bands of 4 pooled rows: 8 NEW conv rows per band go into a 9-row circular band (conv row r lives in slot
            // r % 9), the row above them is still there from the previous band
            for (int py0 = 0; py0 < Ho; py0 += 4) {
                const int cr0 = 2 * py0;
                conv3x3<4, false, false>(x, fp.cin, H, W, wsm, fp.cout, s_in, sh_in, bias, cr0, cr0 + 8, band, 9 * W, cr0 % 9, 9);
                __syncthreads();
                for (int o = tid; o < fp.cout * 4 * Wo; o += IM_THREADS) {
                    const int oc = o >> (lwo + 2), rem = o & (4 * Wo - 1);     // Wo is a power of two
                    const int py = py0 + (rem >> lwo), px = rem & (Wo - 1);
                    float mx = -INFINITY;
#pragma unroll
                    for (int dy = -1; dy <= 1; ++dy) {
                        const int yy = 2 * py + dy;
                        if (yy < 0 || yy >= H) continue;
                        const float* brow = band + oc * 9 * W + (yy % 9) * W;
#pragma unroll
                        for (int dx = -1; dx <= 1; ++dx) {
                            const int xx = 2 * px + dx;
                            if (xx < 0 || xx >= W) continue;
                            mx = fmaxf(mx, brow[xx]);
                        }
                    }
                    t[oc * Ho * Wo + py * Wo + px] = mx;
                }
                __syncthreads();
            }
        } else {
            // 16 x 16: the whole conv output (32 x 16 x 16) fits the band buffer
            conv3x3<4, false, false>(x, fp.cin, H, W, wsm, fp.cout, s_in, sh_in, bias, 0, H, band, H * W, 0, H);
            __syncthreads();
            for (int o = tid; o < fp.cout * Ho * Wo; o += IM_THREADS) {
                const int oc = o >> (2 * lwo), rem = o & (Ho * Wo - 1);
                const int py = rem >> lwo, px = rem & (Wo - 1);
                float mx = -INFINITY;
#pragma unroll
                for (int dy = -1; dy <= 1; ++dy) {
                    const int yy = 2 * py + dy;
                    if (yy < 0 || yy >= H) continue;
#pragma unroll
                    for (int dx = -1; dx <= 1; ++dx) {
                        const int xx = 2 * px + dx;
                        if (xx < 0 || xx >= W) continue;
                        mx = fmaxf(mx, band[oc * H * W + yy * W + xx]);
                    }
                }
                t[o] = mx;
            }
            __syncthreads();
        }
        { float* tmp = x; x = t; t = tmp; }
        H = Ho;
        stamp();        // 2, 5, 8: stage conv + pool done
        // two residual blocks at this resolution: x += conv_b(relu(BN_b(conv_a(relu(BN_a(x))))))
        for (int blk = 0; blk < 2; ++blk) {
            const ConvP& pa = L.res[blk][s][0];
            const ConvP& pb = L.res[blk][s][1];
            next_layer(pa);
            __syncthreads();
            if (s == 0) conv3x3<8, true, false>(x, pa.cin, H, H, wsm, pa.cout, s_in, sh_in, bias, 0, H, t, H * H, 0, H);
            else if (s == 1) conv3x3<4, true, false>(x, pa.cin, H, H, wsm, pa.cout, s_in, sh_in, bias, 0, H, t, H * H, 0, H);
            else conv3x3<2, true, false>(x, pa.cin, H, H, wsm, pa.cout, s_in, sh_in, bias, 0, H, t, H * H, 0, H);
            __syncthreads();
            next_layer(pb);
            __syncthreads();
            if (s == 0) conv3x3<8, true, true>(t, pb.cin, H, H, wsm, pb.cout, s_in, sh_in, bias, 0, H, x, H * H, 0, H);
            else if (s == 1) conv3x3<4, true, true>(t, pb.cin, H, H, wsm, pb.cout, s_in, sh_in, bias, 0, H, x, H * H, 0, H);
            else conv3x3<2, true, true>(t, pb.cin, H, H, wsm, pb.cout, s_in, sh_in, bias, 0, H, x, H * H, 0, H);
            __syncthreads();
            stamp();    // residual block done
        }
    }
    // x: [32][8][8].  relu -> flatten (C,H,W) -> BN1d(2048) -> vec[0..2048)
    float* fcin = vec;             // 2048
    float* core = vec + 2048;      // 257: relu(fc) | clamped reward
    float* hst = core + 260;       // 256: h0
    float* gates = hst + 256;      // 1024
    float* hn = gates + 1024;      // 256: BN'd new h for the policy head
    float* lg = hn + 256;          // 32 logits
    for (int k = tid; k < 2048; k += IM_THREADS) {
        const float inv = 1.0f / sqrtf(bnbuf[L.fc_bv + k] + 1e-5f);
        const float s = c.par(L.fc_g + k) * inv;
        fcin[k] = fmaf(fmaxf(x[k], 0.f), s, c.par(L.fc_be + k) - bnbuf[L.fc_bm + k] * s);
    }
    const bool dn = done[inst] != 0;
    for (int k = tid; k < 256; k += IM_THREADS) hst[k] = dn ? 0.f : h_in[(int64_t)inst * 256 + k];
    __syncthreads();
    // Linear 2048 -> 256 (+ReLU): warp per output row, 8-byte loads (row starts are 8-byte aligned only), 16 + 16 loads
    // (256 B) in flight per lane
    for (int o = warp; o < 256; o += IM_THREADS / 32) {
        const int64_t base = L.fc_w + (int64_t)o * 2048;
        float acc = 0.f;
        for (int v0 = 0; v0 < 1024; v0 += 512) {
            float2 tw[16], ew[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                const int v = v0 + u * 32 + lane;
                tw[u] = *reinterpret_cast<const float2*>(theta + base + 2 * v);
                ew[u] = *reinterpret_cast<const float2*>(c.row + base + 2 * v);
            }
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                const int v = v0 + u * 32 + lane;
                const float2 f = *reinterpret_cast<const float2*>(fcin + 2 * v);
                acc = fmaf(perturb1(tw[u].x, c.sg, ew[u].x), f.x, acc);
                acc = fmaf(perturb1(tw[u].y, c.sg, ew[u].y), f.y, acc);
            }
        }
        acc = warp_sum(acc);
        if (lane == 0) core[o] = fmaxf(acc + c.par(L.fc_b + o), 0.f);
    }
    if (tid == 0) core[256] = fminf(fmaxf(reward[inst], -1.f), 1.f);   // clamp(reward, -1, 1), impala.py:158
    __syncthreads();
    stamp();            // 11: Linear 2048 -> 256 done
    // LSTM gates = W_ih [x;r] + b_ih + W_hh h0 + b_hh   (rows: i | f | g | o, 256 each).  A warp takes two gate rows at a
    // time and issues all their loads first: W_ih rows (257 floats, 4-byte aligned only) lane-strided, W_hh rows (256
    // floats, 8-byte aligned) as float2 - 52 loads = 272 B in flight per lane
    for (int o0 = 2 * warp; o0 < 1024; o0 += 2 * (IM_THREADS / 32)) {
        float ti[2][9], ei[2][9];
        float2 th[2][4], eh[2][4];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int64_t bi = L.wih + (int64_t)(o0 + q) * 257, bh = L.whh + (int64_t)(o0 + q) * 256;
#pragma unroll
            for (int u = 0; u < 9; ++u) {
                const int k = lane + 32 * u;
                ti[q][u] = 0.f;
                ei[q][u] = 0.f;
                if (k < 257) {
                    ti[q][u] = theta[bi + k];
                    ei[q][u] = c.row[bi + k];
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int v = lane + 32 * u;
                th[q][u] = *reinterpret_cast<const float2*>(theta + bh + 2 * v);
                eh[q][u] = *reinterpret_cast<const float2*>(c.row + bh + 2 * v);
            }
        }
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            float acc = 0.f;
#pragma unroll
            for (int u = 0; u < 9; ++u) {
                const int k = lane + 32 * u;
                if (k < 257) acc = fmaf(perturb1(ti[q][u], c.sg, ei[q][u]), core[k], acc);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int v = lane + 32 * u;
                const float2 hv = *reinterpret_cast<const float2*>(hst + 2 * v);
                acc = fmaf(perturb1(th[q][u].x, c.sg, eh[q][u].x), hv.x, acc);
                acc = fmaf(perturb1(th[q][u].y, c.sg, eh[q][u].y), hv.y, acc);
            }
            acc = warp_sum(acc);
            if (lane == 0) gates[o0 + q] = acc + c.par(L.bih + o0 + q) + c.par(L.bhh + o0 + q);
        }
    }
    __syncthreads();
    stamp();            // 12: LSTM gates done
    for (int k = tid; k < 256; k += IM_THREADS) {
        const float c0 = dn ? 0.f : c_in[(int64_t)inst * 256 + k];
        const float ig = sigmoidf_(gates[k]), fg = sigmoidf_(gates[256 + k]);
        const float gg = tanhf(gates[512 + k]), og = sigmoidf_(gates[768 + k]);
        const float c1 = fg * c0 + ig * gg;
        const float h1 = og * tanhf(c1);
        c_out[(int64_t)inst * 256 + k] = c1;
        h_out[(int64_t)inst * 256 + k] = h1;
        const float inv = 1.0f / sqrtf(bnbuf[L.pol_bv + k] + 1e-5f);
        const float s = c.par(L.pol_g + k) * inv;
        hn[k] = fmaf(h1, s, c.par(L.pol_be + k) - bnbuf[L.pol_bm + k] * s);
    }
    __syncthreads();
    for (int a = warp; a < L.A; a += IM_THREADS / 32) {
        float acc = 0.f;
        for (int k = lane; k < 256; k += 32) acc = fmaf(c.par(L.pol_w + a * 256 + k), hn[k], acc);
        acc = warp_sum(acc);
        if (lane == 0) lg[a] = acc + c.par(L.pol_b + a);
    }
    __syncthreads();
    if (tid == 0) {
        float mx = -INFINITY;
        for (int a = 0; a < L.A; ++a) mx = fmaxf(mx, lg[a]);
        float ssum = 0.f;
        for (int a = 0; a < L.A; ++a) ssum += expf(lg[a] - mx);
        const float inv = 1.0f / ssum;
        for (int a = 0; a < L.A; ++a) probs[(int64_t)inst * L.A + a] = expf(lg[a] - mx) * inv;
    }
    stamp();            // 13: end
}

ImpalaP make_impala(int A) {
    ImpalaP L = {};
    int off = 0, boff = 0;
    const int cin_[3] = {3, 16, 32}, cout_[3] = {16, 32, 32};
    auto conv = [&](ConvP& p, int cin, int cout) {
        p.cin = cin; p.cout = cout;
        p.g = off; off += cin;
        p.be = off; off += cin;
        p.bm = boff; boff += cin;
        p.bv = boff; boff += cin;
        boff += 1;   // num_batches_tracked
        p.w = off; off += cout * cin * 9;
        p.b = off; off += cout;
    };
    for (int s = 0; s < 3; ++s) conv(L.feat[s], cin_[s], cout_[s]);
    for (int blk = 0; blk < 2; ++blk)
        for (int s = 0; s < 3; ++s) {
            conv(L.res[blk][s][0], cout_[s], cout_[s]);
            conv(L.res[blk][s][1], cout_[s], cout_[s]);
        }
    L.fc_g = off; off += 2048;
    L.fc_be = off; off += 2048;
    L.fc_bm = boff; boff += 2048;
    L.fc_bv = boff; boff += 2048;
    boff += 1;
    L.fc_w = off; off += 256 * 2048;
    L.fc_b = off; off += 256;
    L.wih = off; off += 1024 * 257;
    L.whh = off; off += 1024 * 256;
    L.bih = off; off += 1024;
    L.bhh = off; off += 1024;
    L.pol_g = off; off += 256;
    L.pol_be = off; off += 256;
    L.pol_bm = boff; boff += 256;
    L.pol_bv = boff; boff += 256;
    boff += 1;
    L.pol_w = off; off += A * 256;
    L.pol_b = off; off += A;
    L.A = A;
    L.P = off;
    int n = 0;
    auto seq = [&](const ConvP& p) { L.seq_w[n] = p.w; L.seq_n[n] = p.cout * p.cin * 9; ++n; };
    for (int s = 0; s < 3; ++s) {
        seq(L.feat[s]);
        for (int blk = 0; blk < 2; ++blk) { seq(L.res[blk][s][0]); seq(L.res[blk][s][1]); }
    }
    L.seq_w[15] = 0; L.seq_n[15] = 0;
    return L;
}

}  // namespace

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
    const size_t smem = (size_t)(2 * MAP + BAND + WMAX + 96 + 2048 + 260 + 256 + 1024 + 256 + 32) * sizeof(float);
    DFD_CUDA(cudaFuncSetAttribute(impala_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long* prof = nullptr;
    if (getenv("DFD_IMPALA_PROF")) { cudaMalloc(&prof, 32 * 8); cudaMemset(prof, 0, 32 * 8); }
    impala_forward_kernel<<<n_members * obs_per_member, IM_THREADS, smem, (cudaStream_t)stream>>>(
        L, table->replicas, table->replica_stride, theta, bn_buffers, idx, sign, sigma, frame, reward, done, h_in, c_in,
        obs_per_member, probs, h_out, c_out, n_members, (n_members % 2 == 0) ? 1 : 0, prof);
    DFD_LAUNCHED(ctx);
    if (prof) {
        cudaStreamSynchronize((cudaStream_t)stream);
        long long h[32];
        cudaMemcpy(h, prof, sizeof(h), cudaMemcpyDeviceToHost);
        fprintf(stderr, "[impala timeline] CTA 7, cycles per phase: frame %lld | s0 conv+pool %lld res %lld %lld | s1 conv+pool %lld res %lld %lld | "
                        "s2 conv+pool %lld res %lld %lld | fc %lld | lstm %lld | head %lld | total %lld\n",
                h[1] - h[0], h[2] - h[1], h[3] - h[2], h[4] - h[3], h[5] - h[4], h[6] - h[5], h[7] - h[6], h[8] - h[7], h[9] - h[8],
                h[10] - h[9], h[11] - h[10], h[12] - h[11], h[13] - h[12], h[13] - h[0]);
        cudaFree(prof);
    }
    return 0;
}
