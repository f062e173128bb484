// Atari perturbed forward (policies/atari.py:35-51): Conv(4->16,k8,s4) BN ReLU -> Conv(16->32,k4,s2) BN ReLU
// -> flatten(C,H,W) -> Linear(2592->256) BN ReLU -> Linear(256->A) -> softmax.  Eval-mode BN with shared
// running statistics and per-member (perturbed) gamma / beta.  Member m evaluates
// theta + sign[m]*sigma*table[idx[m]:idx[m]+P] (worker/worker.py:28); the perturbed weights only ever exist in
// shared memory / registers.  98 % of the parameters are the first Linear (663 552 weights), so at small E the
// kernel is a stream of that weight block through the SM: HBM-bound on pairs*P*4 bytes.
// The first Linear streams through shared-memory slots fed by cp.async.bulk (TMA bulk copies, half a theta row and
// half an eps row of 5 184 bytes each per item, completion on mbarriers) issued by a producer warp: 14 consumer warps
// own one slot each (items w, w + 14, ...), the slots live in the frame / conv buffers that are dead by then
// (145 KB in flight per SM, no L1 line reservations - register-staged loads were capped by the ~28 KB of L1 left
// beside 200 KB of shared memory).
// Pair mode (E <= 2, even member count): one CTA evaluates members j and j + M/2 - the two members of an antithetic
// pair in [plus | minus] batches -, runs the conv stack for each and then streams theta and the SHARED eps row once for
// both (if the two indices differ, the first Linear simply runs one pass per member).  Otherwise consecutive CTAs take
// the two members of a pair, so HBM at least serves their shared eps row once.
#include "common.cuh"
#include <stdlib.h>

namespace {

constexpr int AT_THREADS = 512;       // worker threads; one more warp feeds the first Linear's ring
constexpr int AT_LAUNCH = AT_THREADS + 32;
constexpr int AT_HALF = 1296;         // half an FC1 weight row
constexpr int AT_SLOT = 2 * AT_HALF;  // floats per ring slot: theta half-row | eps half-row (10 368 B)
constexpr int AT_NSLOT = 14;          // one slot per consumer warp; 14 * 10 368 B + partial sums <= frame + w0 + sc0 + a0 (155 008 B)
constexpr int AT_ET = 4;              // observations per CTA pass (FC1 weights are streamed once per pass)
constexpr int FRAME = 4 * 84 * 84;    // 28224
constexpr int A0N = 16 * 20 * 20;     // 6400
constexpr int A1N = 32 * 9 * 9;       // 2592

// flat parameter offsets and BN buffer offsets (state_dict order)
constexpr int O_W0 = 0, O_B0 = 4096, O_G1 = 4112, O_BE1 = 4128, O_W3 = 4144, O_B3 = 12336, O_G4 = 12368,
              O_BE4 = 12400, O_W7 = 12432, O_B7 = 675984, O_G8 = 676240, O_BE8 = 676496, O_W10 = 676752;
constexpr int BU_M1 = 0, BU_V1 = 16, BU_M4 = 33, BU_V4 = 65, BU_M8 = 98, BU_V8 = 354;

__global__ void __launch_bounds__(AT_LAUNCH, 1) atari_forward_kernel(const float* __restrict__ replicas, int64_t stride,
                                                                      const float* __restrict__ theta,
                                                                      const float* __restrict__ bnbuf,
                                                                      const int64_t* __restrict__ idx,
                                                                      const int8_t* __restrict__ sign, float sigma,
                                                                      const float* __restrict__ obs, int E, int tiles,
                                                                      int A, float* __restrict__ out, int n_members,
                                                                      int pair_order) {
    extern __shared__ __align__(16) float sm[];
    float* frame = sm;                     // 28224; conv3 weights alias it once conv0 is done with the frame
    float* w0 = frame + FRAME;             // [256 k][16 oc]
    float* sc0 = w0 + 4096;                // 16 scale | 16 shift (conv bias and BN folded)
    float* a0 = sc0 + 32;                  // [16][20][20]
    float* a1 = a0 + A0N;                  // [ET][2592]
    float* a2 = a1 + AT_ET * A1N;          // [ET][256]
    float* lg = a2 + AT_ET * 256;          // [ET][32] logits
    float* w3 = frame;                     // [256 k][32 oc]
    float* sc3 = frame + 8192;             // 32 scale | 32 shift

    __shared__ __align__(8) uint64_t bars[2 * AT_NSLOT];      // full[slot], empty[slot]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool worker = tid < AT_THREADS;                    // warp 16 only feeds the ring (and joins the barriers)
    const int nmem = pair_order == 2 ? 2 : 1;                // members evaluated by this CTA
    const int mb = pair_order == 2 ? (int)blockIdx.x : blockIdx.x / tiles, tile = pair_order == 2 ? 0 : blockIdx.x % tiles;
    const int m0 = pair_order == 1 ? ((mb & 1) ? (n_members >> 1) + (mb >> 1) : (mb >> 1)) : mb;
    const int ms[2] = {m0, pair_order == 2 ? m0 + (n_members >> 1) : m0};
    const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(&bars[0]);
    if (tid == 0) {
        for (int i = 0; i < AT_NSLOT; ++i) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8u * i));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8u * (AT_NSLOT + i)));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const int e0 = tile * AT_ET;
    const int ne = min(AT_ET, E - e0);                       // pair mode: ne <= 2, activation slot = member * 2 + e
    const float sgs[2] = {sigma * (float)sign[ms[0]], sigma * (float)sign[ms[1]]};
    const float* rows[2] = {table_row_ptr(replicas, stride, idx[ms[0]]), table_row_ptr(replicas, stride, idx[ms[1]])};
    float sg = sgs[0];
    const float* row = rows[0];
    int m = ms[0];
    auto par = [&](int p) { return perturb1(theta[p], sg, row[p]); };
    auto slot_of = [&](int mem, int e) { return nmem == 2 ? mem * 2 + e : e; };

    for (int mem = 0; mem < nmem; ++mem) {
    sg = sgs[mem];
    row = rows[mem];
    m = ms[mem];
    __syncthreads();       // the previous member's conv stack is done with w0 / sc0 / frame / a0

    // conv0 weights -> [k][oc]; bias + BN1 folded into per-channel scale / shift
    for (int t = tid; t < 4096 && worker; t += AT_THREADS) {
        const int oc = t >> 8, k = t & 255;
        w0[k * 16 + oc] = par(O_W0 + t);
    }
    if (tid < 16) {
        const float inv = 1.0f / sqrtf(bnbuf[BU_V1 + tid] + 1e-5f);
        const float s = par(O_G1 + tid) * inv;
        sc0[tid] = s;
        sc0[16 + tid] = (par(O_B0 + tid) - bnbuf[BU_M1 + tid]) * s + par(O_BE1 + tid);
    }

    for (int e = 0; e < ne; ++e) {
        __syncthreads();   // previous observation is done with frame / w3 / a0
        const float* fr = obs + ((int64_t)m * E + e0 + e) * FRAME;
        for (int t = tid; t < FRAME / 4 && worker; t += AT_THREADS)
            reinterpret_cast<float4*>(frame)[t] = ldg_stream_f4(fr + 4 * t);
        __syncthreads();
        // ---- conv0: one output pixel x 16 channels per thread (400 threads)
        if (tid < 400) {
            const int oy = tid / 20, ox = tid - oy * 20;
            float acc[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = 0.f;
            for (int c = 0; c < 4; ++c) {
                for (int ky = 0; ky < 8; ++ky) {
                    const float* xr = frame + c * 7056 + (4 * oy + ky) * 84 + 4 * ox;
                    const float4 x0 = *reinterpret_cast<const float4*>(xr);
                    const float4 x1 = *reinterpret_cast<const float4*>(xr + 4);
                    const float xs[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
                    const float* wk = w0 + (c * 64 + ky * 8) * 16;
#pragma unroll
                    for (int kx = 0; kx < 8; ++kx) {
                        const float4* w4 = reinterpret_cast<const float4*>(wk + kx * 16);
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float4 w = w4[q];
                            acc[4 * q + 0] = fmaf(w.x, xs[kx], acc[4 * q + 0]);
                            acc[4 * q + 1] = fmaf(w.y, xs[kx], acc[4 * q + 1]);
                            acc[4 * q + 2] = fmaf(w.z, xs[kx], acc[4 * q + 2]);
                            acc[4 * q + 3] = fmaf(w.w, xs[kx], acc[4 * q + 3]);
                        }
                    }
                }
            }
#pragma unroll
            for (int oc = 0; oc < 16; ++oc) a0[oc * 400 + tid] = fmaxf(fmaf(acc[oc], sc0[oc], sc0[16 + oc]), 0.f);
        }
        __syncthreads();
        // ---- conv3 weights -> [k][oc] over the (now free) frame region; bias + BN4 folded
        for (int t = tid; t < 8192 && worker; t += AT_THREADS) {
            const int oc = t >> 8, k = t & 255;
            w3[k * 32 + oc] = par(O_W3 + t);
        }
        if (tid < 32) {
            const float inv = 1.0f / sqrtf(bnbuf[BU_V4 + tid] + 1e-5f);
            const float s = par(O_G4 + tid) * inv;
            sc3[tid] = s;
            sc3[32 + tid] = (par(O_B3 + tid) - bnbuf[BU_M4 + tid]) * s + par(O_BE4 + tid);
        }
        __syncthreads();
        // ---- conv3: one output pixel x 8 channels per thread (81 pixels x 4 channel groups = 324 threads)
        if (tid < 324) {
            const int pix = tid % 81, og = tid / 81;
            const int oy = pix / 9, ox = pix - oy * 9;
            float acc[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = 0.f;
            for (int c = 0; c < 16; ++c) {
#pragma unroll
                for (int ky = 0; ky < 4; ++ky) {
                    const float* xr = a0 + c * 400 + (2 * oy + ky) * 20 + 2 * ox;
                    const float2 xa = *reinterpret_cast<const float2*>(xr);
                    const float2 xb = *reinterpret_cast<const float2*>(xr + 2);
                    const float xs[4] = {xa.x, xa.y, xb.x, xb.y};
                    const float* wk = w3 + (c * 16 + ky * 4) * 32 + og * 8;
#pragma unroll
                    for (int kx = 0; kx < 4; ++kx) {
                        const float4 wa = *reinterpret_cast<const float4*>(wk + kx * 32);
                        const float4 wb = *reinterpret_cast<const float4*>(wk + kx * 32 + 4);
                        acc[0] = fmaf(wa.x, xs[kx], acc[0]);
                        acc[1] = fmaf(wa.y, xs[kx], acc[1]);
                        acc[2] = fmaf(wa.z, xs[kx], acc[2]);
                        acc[3] = fmaf(wa.w, xs[kx], acc[3]);
                        acc[4] = fmaf(wb.x, xs[kx], acc[4]);
                        acc[5] = fmaf(wb.y, xs[kx], acc[5]);
                        acc[6] = fmaf(wb.z, xs[kx], acc[6]);
                        acc[7] = fmaf(wb.w, xs[kx], acc[7]);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int oc = og * 8 + i;
                a1[slot_of(mem, e) * A1N + oc * 81 + pix] = fmaxf(fmaf(acc[i], sc3[oc], sc3[32 + oc]), 0.f);   // flatten (C,H,W)
            }
        }
    }
    }   // members
    __syncthreads();       // a1 complete; frame / w0 / a0 are dead from here on: they become the weight ring
    bool act[AT_ET];                                         // which activation slots hold an observation
#pragma unroll
    for (int sl = 0; sl < AT_ET; ++sl) act[sl] = nmem == 2 ? ((sl & 1) < ne) : (sl < ne);
    const bool shared_row = nmem == 2 && rows[0] == rows[1];
    const int npass = (nmem == 2 && !shared_row) ? 2 : 1;    // pair with one table row: ONE pass feeds both members

    // ---- Linear 2592 -> 256: item h = (neuron h >> 1, half h & 1) -> slot h % 14, owned by consumer warp h % 14 (one
    //      waiter per barrier, phases strictly in order); the producer warp issues one theta and one eps bulk copy per
    //      item; AT_ET observations share every weight; the two halves meet in shared memory, then bias + BN8 + ReLU
    float* ring = sm;
    float* pacc = sm + AT_NSLOT * AT_SLOT;     // [AT_ET][512] partial dot products
    auto wait_bar = [&](uint32_t addr, uint32_t parity) {
        uint32_t ok, spins = 0;
        do {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                         : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
            if (!ok && ++spins > (1u << 26)) __trap();
        } while (!ok);
    };
    if (!worker) {
        if (lane == 0) {
            // the ring region was last written through the generic proxy (frame, conv weights, activations)
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            for (int hg = 0; hg < npass * 512; ++hg) {
                const int h = hg & 511, pass = hg >> 9;
                const float* row = rows[pass];               // pass 0: member 0's row (= member 1's in a true pair)
                const int slot = hg % AT_NSLOT, it = hg / AT_NSLOT;
                wait_bar(bar0 + 8u * (AT_NSLOT + slot), (uint32_t)((it & 1) ^ 1));
                const uint32_t dst = (uint32_t)__cvta_generic_to_shared(ring + slot * AT_SLOT);
                const uint32_t fb = bar0 + 8u * slot;
                const int64_t base = O_W7 + (int64_t)(h >> 1) * A1N + (h & 1) * AT_HALF;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fb), "r"(2u * AT_HALF * 4u) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                             "l"(theta + base), "r"((uint32_t)(AT_HALF * 4)), "r"(fb) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 dst + AT_HALF * 4u), "l"(row + base), "r"((uint32_t)(AT_HALF * 4)), "r"(fb) : "memory");
            }
        }
    } else if (warp < AT_NSLOT) {
        int it = 0;
        for (int hg = warp; hg < npass * 512; hg += AT_NSLOT, ++it) {
            const int h = hg & 511, pass = hg >> 9;
            wait_bar(bar0 + 8u * warp, (uint32_t)(it & 1));
            const float4* tw4 = reinterpret_cast<const float4*>(ring + warp * AT_SLOT);
            const float4* ew4 = tw4 + AT_HALF / 4;
            const float* xa = a1 + (h & 1) * AT_HALF;
            // which slots this pass serves and with which signed sigma: single member: all; true pair: both members
            // from the one row; two unrelated members: the pass's own member only
            bool on[AT_ET];
            float sgl[AT_ET];
#pragma unroll
            for (int sl = 0; sl < AT_ET; ++sl) {
                const int mem = nmem == 2 ? (sl >> 1) : 0;
                on[sl] = act[sl] && (npass == 1 || mem == pass);
                sgl[sl] = sgs[mem];
            }
            float acc[AT_ET];
#pragma unroll
            for (int e = 0; e < AT_ET; ++e) acc[e] = 0.f;
            for (int v = lane; v < AT_HALF / 4; v += 32) {
                const float4 t4 = tw4[v], e4 = ew4[v];
#pragma unroll
                for (int e = 0; e < AT_ET; ++e) {
                    if (on[e]) {
                        const float4 w = make_float4(perturb1(t4.x, sgl[e], e4.x), perturb1(t4.y, sgl[e], e4.y),
                                                     perturb1(t4.z, sgl[e], e4.z), perturb1(t4.w, sgl[e], e4.w));
                        const float4 x = *reinterpret_cast<const float4*>(xa + e * A1N + 4 * v);
                        acc[e] = fmaf(w.x, x.x, acc[e]);
                        acc[e] = fmaf(w.y, x.y, acc[e]);
                        acc[e] = fmaf(w.z, x.z, acc[e]);
                        acc[e] = fmaf(w.w, x.w, acc[e]);
                    }
                }
            }
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar0 + 8u * (AT_NSLOT + warp)) : "memory");
#pragma unroll
            for (int e = 0; e < AT_ET; ++e) acc[e] = warp_sum(acc[e]);
            if (lane == 0) {
#pragma unroll
                for (int e = 0; e < AT_ET; ++e)
                    if (on[e]) pacc[e * 512 + h] = acc[e];
            }
        }
    }
    __syncthreads();
    if (tid < 256) {
        const int o = tid;
        const float inv = 1.0f / sqrtf(bnbuf[BU_V8 + o] + 1e-5f);
        for (int mem = 0; mem < nmem; ++mem) {
            sg = sgs[mem];
            row = rows[mem];
            const float s = par(O_G8 + o) * inv;
            const float sh = (par(O_B7 + o) - bnbuf[BU_M8 + o]) * s + par(O_BE8 + o);
#pragma unroll
            for (int sl = 0; sl < AT_ET; ++sl) {
                const bool mine = nmem == 2 ? (sl >> 1) == mem : true;
                if (mine && act[sl]) a2[sl * 256 + o] = fmaxf(fmaf(pacc[sl * 512 + 2 * o] + pacc[sl * 512 + 2 * o + 1], s, sh), 0.f);
            }
        }
    }
    __syncthreads();
    // ---- Linear 256 -> A (warp per action), softmax per observation
    for (int a = warp; a < A && worker; a += AT_THREADS / 32) {
        for (int mem = 0; mem < nmem; ++mem) {
            sg = sgs[mem];
            row = rows[mem];
            float w[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) w[j] = par(O_W10 + a * 256 + lane + 32 * j);
            const float b = par(O_W10 + A * 256 + a);
            for (int e = 0; e < ne; ++e) {
                const int sl = slot_of(mem, e);
                float s = 0.f;
#pragma unroll
                for (int j = 0; j < 8; ++j) s = fmaf(w[j], a2[sl * 256 + lane + 32 * j], s);
                s = warp_sum(s);
                if (lane == 0) lg[sl * 32 + a] = s + b;
            }
        }
    }
    __syncthreads();
    if (tid < AT_ET && act[tid]) {
        const int mem = nmem == 2 ? (tid >> 1) : 0, e = nmem == 2 ? (tid & 1) : tid;
        const float* l = lg + tid * 32;
        float mx = -INFINITY;
        for (int a = 0; a < A; ++a) mx = fmaxf(mx, l[a]);
        float s = 0.f;
        for (int a = 0; a < A; ++a) s += expf(l[a] - mx);
        const float inv = 1.0f / s;
        float* o = out + ((int64_t)ms[mem] * E + e0 + e) * A;
        for (int a = 0; a < A; ++a) o[a] = expf(l[a] - mx) * inv;
    }
}

}  // namespace

int dfd_atari_forward_tc_impl(dfd_ctx* ctx, const dfd_policy_desc* desc, const dfd_table* table, const float* theta,
                              const float* bn_buffers, const int64_t* idx, const int8_t* sign, int n_members, float sigma,
                              const float* obs, int obs_per_member, float* out, cudaStream_t st);

int dfd_atari_forward_impl(dfd_ctx* ctx, const dfd_policy_desc* desc, const dfd_table* table, const float* theta,
                           const float* bn_buffers, const int64_t* idx, const int8_t* sign, int n_members, float sigma,
                           const float* obs, int obs_per_member, float* out, cudaStream_t st) {
    DFD_CHECK_ARG(bn_buffers, "dfd_policy_forward: Atari needs bn_buffers");
    DFD_CHECK_ARG(desc->n_act >= 1 && desc->n_act <= 32, "dfd_policy_forward: Atari n_act %d out of range (1..32)", desc->n_act);
    DFD_CHECK_ARG((((uintptr_t)obs) & 15) == 0 && (((uintptr_t)theta) & 15) == 0,
                  "dfd_policy_forward: Atari obs / theta must be 16-byte aligned");
    DFD_CHECK_ARG(dfd_policy_num_params(desc) < table->size, "dfd_policy_forward: num_params >= table size");
    if (desc->precision >= 1) {       // tensor path (csrc/cnn_forward_tc.cu) when the scaled table mirror is registered
        const int rt = dfd_atari_forward_tc_impl(ctx, desc, table, theta, bn_buffers, idx, sign, n_members, sigma, obs,
                                                 obs_per_member, out, st);
        if (rt >= 0) return rt;
    }
    const int tiles = (obs_per_member + AT_ET - 1) / AT_ET;
    DFD_CHECK_ARG((int64_t)n_members * tiles < 2147483647LL, "dfd_policy_forward: grid too large");
    const size_t smem = (size_t)(FRAME + 4096 + 32 + A0N + AT_ET * A1N + AT_ET * 256 + AT_ET * 32) * sizeof(float);
    DFD_CUDA(cudaFuncSetAttribute(atari_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // 2: one CTA per pair of members (E <= 2); 1: one CTA per member, pair-adjacent order; 0: plain order
    int mode = (n_members % 2 == 0) ? ((obs_per_member <= 2 && !getenv("DFD_ATARI_NO_PAIR")) ? 2 : 1) : 0;
    const int grid = mode == 2 ? n_members / 2 : n_members * tiles;
    atari_forward_kernel<<<grid, AT_LAUNCH, smem, st>>>(table->replicas, table->replica_stride, theta, bn_buffers, idx, sign,
                                                        sigma, obs, obs_per_member, tiles, desc->n_act, out, n_members, mode);
    DFD_LAUNCHED(ctx);
    return 0;
}
