// IMPALA CNN + LSTM perturbed forward (policies/impala.py:136-186), one CTA per (member, environment).
//   x = frame/255; 3 stages {BN -> conv3x3 -> maxpool(3,2,1); 2 x [x += conv(relu(BN(conv(relu(BN(x))))))]};
//   relu -> flatten(2048) -> relu(Linear(BN1d(x))) -> concat clamp(reward,-1,1) -> LSTM cell (gates i,f,g,o,
//   state zeroed where done) -> Linear(BN1d(h)) -> softmax.
// BN layers are eval-mode with shared running statistics and per-member (perturbed) gamma / beta; every
// parameter is theta + sign*sigma*eps generated in-kernel (worker/worker.py:28).  The whole conv trunk keeps its
// activations in shared memory (two ping-pong maps + a conv-row band for the pooled stages); only the
// carried LSTM state and the action probabilities touch HBM.  Round-1 version: exact fp32 on CUDA cores.
#include "common.cuh"

namespace {

constexpr int IM_THREADS = 512;
constexpr int MAP = 16384;    // floats per activation map buffer (16 x 32 x 32)
constexpr int BAND = 9216;    // conv-row band for the pooled stages: 9 rows x (64 x 16 | 32 x 32 | 16 x 32 (+pad))
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
};

struct Ctx {
    const float* theta;
    const float* row;
    const float* bn;
    float sg;
    __device__ __forceinline__ float par(int p) const { return perturb1(theta[p], sg, row[p]); }
};

// weights -> wsm[(ci*9+tap)*cout + oc]; input-side BN folded to per-input-channel scale/shift; conv bias
__device__ void load_conv(const Ctx& c, const ConvP& p, float* wsm, float* s_in, float* sh_in, float* bias) {
    const int n = p.cout * p.cin * 9;
    // all of a thread's loads of a batch are issued before the first shared-memory store (the compiler cannot hoist
    // global loads over stores it cannot prove disjoint, which would expose one memory latency per element)
    constexpr int U = 8;
    for (int t0 = threadIdx.x; t0 < n; t0 += IM_THREADS * U) {
        float a[U], e[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int t = t0 + u * IM_THREADS;
            a[u] = 0.f;
            e[u] = 0.f;
            if (t < n) {
                a[u] = c.theta[p.w + t];
                e[u] = c.row[p.w + t];
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int t = t0 + u * IM_THREADS;
            if (t < n) {
                const int oc = t / (p.cin * 9), k = t - oc * (p.cin * 9);
                wsm[k * p.cout + oc] = perturb1(a[u], c.sg, e[u]);
            }
        }
    }
    for (int t = threadIdx.x; t < p.cin; t += IM_THREADS) {
        const float inv = 1.0f / sqrtf(c.bn[p.bv + t] + 1e-5f);
        const float s = c.par(p.g + t) * inv;
        s_in[t] = s;
        sh_in[t] = c.par(p.be + t) - c.bn[p.bm + t] * s;
    }
    for (int t = threadIdx.x; t < p.cout; t += IM_THREADS) bias[t] = c.par(p.b + t);
}

// conv3x3 pad 1 over output rows [r0, r1) of an H x W map.  in: [cin][H][W]; the input is BN'd (scale/shift)
// and optionally ReLU'd on the fly, zero padding applies AFTER that (torch pads the BN output).
// dst element (oc, r, x) at dst[oc*dst_cs + (r - r0 + dst_r0)*W + x]; accumulate adds to dst (residual).
template <bool RELU_IN, bool ACCUM>
__device__ void conv3x3(const float* __restrict__ in, int cin, int H, int W, const float* __restrict__ wsm, int cout,
                        const float* __restrict__ s_in, const float* __restrict__ sh_in,
                        const float* __restrict__ bias, int r0, int r1, float* __restrict__ dst, int dst_cs, int dst_r0) {
    const int npix = (r1 - r0) * W, ngrp = cout >> 3;
    for (int item = threadIdx.x; item < npix * ngrp; item += IM_THREADS) {
        const int og = item / npix, pix = item - og * npix;
        const int r = r0 + pix / W, x = pix % W;
        float acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = bias[og * 8 + i];
        for (int ci = 0; ci < cin; ++ci) {
            const float s = s_in[ci], sh = sh_in[ci];
            const float* ip = in + ci * H * W;
            const float* wp = wsm + ci * 9 * cout + og * 8;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                const int yy = r + ky - 1;
                if (yy < 0 || yy >= H) continue;
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const int xx = x + kx - 1;
                    if (xx < 0 || xx >= W) continue;
                    float v = fmaf(ip[yy * W + xx], s, sh);
                    if (RELU_IN) v = fmaxf(v, 0.f);
                    const float4 wa = *reinterpret_cast<const float4*>(wp + (ky * 3 + kx) * cout);
                    const float4 wb = *reinterpret_cast<const float4*>(wp + (ky * 3 + kx) * cout + 4);
                    acc[0] = fmaf(wa.x, v, acc[0]);
                    acc[1] = fmaf(wa.y, v, acc[1]);
                    acc[2] = fmaf(wa.z, v, acc[2]);
                    acc[3] = fmaf(wa.w, v, acc[3]);
                    acc[4] = fmaf(wb.x, v, acc[4]);
                    acc[5] = fmaf(wb.y, v, acc[5]);
                    acc[6] = fmaf(wb.z, v, acc[6]);
                    acc[7] = fmaf(wb.w, v, acc[7]);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float* d = dst + (og * 8 + i) * dst_cs + (r - r0 + dst_r0) * W + x;
            if (ACCUM) *d += acc[i]; else *d = acc[i];
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
                                                                       float* __restrict__ c_out, int n_members, int pair_order) {
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

    // frame / 255 -> bufA  (impala.py:142)
    const float* fr = frame + (int64_t)inst * 12288;
    for (int t = tid; t < 12288; t += IM_THREADS) bufA[t] = fr[t] / 255.0f;

    float* x = bufA;    // current map
    float* t = bufB;    // the other buffer
    int H = 64;
    for (int s = 0; s < 3; ++s) {
        const ConvP& fp = L.feat[s];
        __syncthreads();
        load_conv(c, fp, wsm, s_in, sh_in, bias);
        __syncthreads();
        // conv (BN on the input, no ReLU) + maxpool 3x3 stride 2 pad 1 (-inf padding), in bands of 4 pooled rows
        const int W = H, Ho = H / 2, Wo = W / 2;
        for (int py0 = 0; py0 < Ho; py0 += 4) {
            const int cr0 = max(2 * py0 - 1, 0), cr1 = min(2 * py0 + 8, H);   // conv rows needed by pooled rows py0..py0+3
            conv3x3<false, false>(x, fp.cin, H, W, wsm, fp.cout, s_in, sh_in, bias, cr0, cr1, band, 9 * W, 0);
            __syncthreads();
            for (int o = tid; o < fp.cout * 4 * Wo; o += IM_THREADS) {
                const int oc = o / (4 * Wo), rem = o - oc * 4 * Wo;
                const int py = py0 + rem / Wo, px = rem % Wo;
                float mx = -INFINITY;
#pragma unroll
                for (int dy = -1; dy <= 1; ++dy) {
                    const int yy = 2 * py + dy;
                    if (yy < 0 || yy >= H) continue;
#pragma unroll
                    for (int dx = -1; dx <= 1; ++dx) {
                        const int xx = 2 * px + dx;
                        if (xx < 0 || xx >= W) continue;
                        mx = fmaxf(mx, band[oc * 9 * W + (yy - cr0) * W + xx]);
                    }
                }
                t[oc * Ho * Wo + py * Wo + px] = mx;
            }
            __syncthreads();
        }
        { float* tmp = x; x = t; t = tmp; }
        H = Ho;
        // two residual blocks at this resolution: x += conv_b(relu(BN_b(conv_a(relu(BN_a(x))))))
        for (int blk = 0; blk < 2; ++blk) {
            const ConvP& pa = L.res[blk][s][0];
            const ConvP& pb = L.res[blk][s][1];
            load_conv(c, pa, wsm, s_in, sh_in, bias);
            __syncthreads();
            conv3x3<true, false>(x, pa.cin, H, H, wsm, pa.cout, s_in, sh_in, bias, 0, H, t, H * H, 0);
            __syncthreads();
            load_conv(c, pb, wsm, s_in, sh_in, bias);
            __syncthreads();
            conv3x3<true, true>(t, pb.cin, H, H, wsm, pb.cout, s_in, sh_in, bias, 0, H, x, H * H, 0);
            __syncthreads();
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
    // Linear 2048 -> 256 (+ReLU): warp per output row, 8-byte loads (row starts are 8-byte aligned only)
    for (int o = warp; o < 256; o += IM_THREADS / 32) {
        const int64_t base = L.fc_w + (int64_t)o * 2048;
        float acc = 0.f;
        for (int v0 = 0; v0 < 1024; v0 += 128) {
            float2 tw[4], ew[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int v = v0 + u * 32 + lane;
                tw[u] = *reinterpret_cast<const float2*>(theta + base + 2 * v);
                ew[u] = *reinterpret_cast<const float2*>(c.row + base + 2 * v);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int v = v0 + u * 32 + lane;
                acc = fmaf(perturb1(tw[u].x, c.sg, ew[u].x), fcin[2 * v], acc);
                acc = fmaf(perturb1(tw[u].y, c.sg, ew[u].y), fcin[2 * v + 1], acc);
            }
        }
        acc = warp_sum(acc);
        if (lane == 0) core[o] = fmaxf(acc + c.par(L.fc_b + o), 0.f);
    }
    if (tid == 0) core[256] = fminf(fmaxf(reward[inst], -1.f), 1.f);   // clamp(reward, -1, 1), impala.py:158
    __syncthreads();
    // LSTM gates = W_ih [x;r] + b_ih + W_hh h0 + b_hh   (rows: i | f | g | o, 256 each)
    for (int o = warp; o < 1024; o += IM_THREADS / 32) {
        const int64_t bi = L.wih + (int64_t)o * 257, bh = L.whh + (int64_t)o * 256;
        float acc = 0.f;
        for (int k = lane; k < 257; k += 32) acc = fmaf(c.par((int)(bi + k)), core[k], acc);
        for (int k = lane; k < 256; k += 32) acc = fmaf(c.par((int)(bh + k)), hst[k], acc);
        acc = warp_sum(acc);
        if (lane == 0) gates[o] = acc + c.par(L.bih + o) + c.par(L.bhh + o);
    }
    __syncthreads();
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
    impala_forward_kernel<<<n_members * obs_per_member, IM_THREADS, smem, (cudaStream_t)stream>>>(
        L, table->replicas, table->replica_stride, theta, bn_buffers, idx, sign, sigma, frame, reward, done, h_in, c_in,
        obs_per_member, probs, h_out, c_out, n_members, (n_members % 2 == 0) ? 1 : 0);
    DFD_LAUNCHED(ctx);
    return 0;
}
