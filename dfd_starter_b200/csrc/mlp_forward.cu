// Per-member perturbed MLP forward, exact fp32 path (CUDA cores).
//   MuJoCo   policies/mujoco.py:35-41 + utils/torch_helpers.py:20-25
//   Discrete policies/discrete.py:37-48 (eval-mode BN, per-member gamma/beta)
// Member m evaluates theta + sign[m]*sigma*table[idx[m]:idx[m]+P] (worker/worker.py:28);
// the perturbed weights are produced in shared memory from the theta tile and the
// table-row tile and are never written to HBM.
#include "common.cuh"
#include <stdlib.h>

struct MlpLayout {
    int kind, K, h1, h2, nout, A;
    // offsets into the flat parameter vector (parameters() order)
    int bn_g[3], bn_b[3];  // discrete only
    int w[3], b[3];
    int in[3], out[3];
    // offsets into the BN buffer vector (state_dict order: mean, var, num_batches_tracked per BN)
    int bn_mean[3], bn_var[3];
    int64_t P;
};

static MlpLayout make_layout(const dfd_policy_desc* d) {
    MlpLayout L = {};
    L.kind = d->kind;
    L.K = d->n_in;
    L.h1 = d->h1;
    L.h2 = d->h2;
    L.A = d->n_act;
    L.nout = d->kind == DFD_POLICY_MUJOCO ? 2 * d->n_act : d->n_act;
    L.in[0] = L.K;  L.out[0] = L.h1;
    L.in[1] = L.h1; L.out[1] = L.h2;
    L.in[2] = L.h2; L.out[2] = L.nout;
    int off = 0, boff = 0;
    for (int l = 0; l < 3; ++l) {
        if (d->kind == DFD_POLICY_DISCRETE) {
            L.bn_g[l] = off; off += L.in[l];
            L.bn_b[l] = off; off += L.in[l];
            L.bn_mean[l] = boff; boff += L.in[l];
            L.bn_var[l] = boff; boff += L.in[l];
            boff += 1;  // num_batches_tracked
        }
        L.w[l] = off; off += L.in[l] * L.out[l];
        L.b[l] = off; off += L.out[l];
    }
    L.P = off;
    return L;
}

extern "C" int64_t dfd_policy_num_params(const dfd_policy_desc* d) {
    if (!d) return -1;
    switch (d->kind) {
        case DFD_POLICY_MUJOCO:
        case DFD_POLICY_DISCRETE: return make_layout(d).P;
        case DFD_POLICY_ATARI: return 12432 + 663552 + 256 + 512 + (int64_t)d->n_act * 256 + d->n_act;
        case DFD_POLICY_IMPALA: return 1154854 + (int64_t)d->n_act * 256 + d->n_act;
        default: return -1;
    }
}

extern "C" int64_t dfd_policy_num_buffers(const dfd_policy_desc* d) {
    if (!d) return -1;
    switch (d->kind) {
        case DFD_POLICY_MUJOCO: return 0;
        case DFD_POLICY_DISCRETE: return 2 * (d->n_in + d->h1 + d->h2) + 3;
        case DFD_POLICY_ATARI: return 2 * (16 + 32 + 256) + 3;
        case DFD_POLICY_IMPALA: return 5367;
        default: return -1;
    }
}

extern "C" int64_t dfd_policy_out_width(const dfd_policy_desc* d) {
    if (!d) return -1;
    return d->kind == DFD_POLICY_MUJOCO ? 2 * d->n_act : d->n_act;
}

static const int MLP_THREADS = 256;
static const int MLP_WBUF = 12288;  // floats of staged perturbed weights per chunk (48 KB)

// act buffers are k-major: act[k*ET + e]
template <int ET>
__global__ void __launch_bounds__(MLP_THREADS) mlp_forward_fp32_kernel(MlpLayout L, const float* __restrict__ replicas,
                                                                       int64_t stride, const float* __restrict__ theta,
                                                                       const float* __restrict__ bnbuf,
                                                                       const int64_t* __restrict__ idx,
                                                                       const int8_t* __restrict__ sign, float sigma,
                                                                       const float* __restrict__ obs, int E,
                                                                       float* __restrict__ out, int maxdim) {
    extern __shared__ __align__(16) float smem[];
    float* actA = smem;
    float* actB = actA + (size_t)maxdim * ET;
    float* wbuf = actB + (size_t)maxdim * ET;
    float* bbuf = wbuf + MLP_WBUF;      // biases of the chunk (<= maxdim)
    float* bnS = bbuf + maxdim;         // BN scale / shift of the current layer input
    float* bnB = bnS + maxdim;

    const int m = blockIdx.x;
    const int e0 = blockIdx.y * ET;
    const int ne = min(ET, E - e0);
    const int tid = threadIdx.x;
    const float sg = sigma * (float)sign[m];
    const float* row = table_row_ptr(replicas, stride, idx[m]);
    auto par = [&](int p) { return perturb1(theta[p], sg, row[p]); };

    // observations -> actA (k-major), zero for the padded observations of the tile
    const float* ob = obs + ((int64_t)m * E + e0) * L.K;
    for (int t = tid; t < L.K * ET; t += MLP_THREADS) {
        const int e = t / L.K, k = t - e * L.K;
        actA[k * ET + e] = e < ne ? ob[(int64_t)e * L.K + k] : 0.f;
    }
    float* ain = actA;
    float* aout = actB;
    __syncthreads();

    for (int l = 0; l < 3; ++l) {
        const int in = L.in[l], no = L.out[l];
        if (L.kind == DFD_POLICY_DISCRETE) {
            // eval-mode BatchNorm1d on the layer input, gamma/beta perturbed per member, running stats shared
            for (int k = tid; k < in; k += MLP_THREADS) {
                const float invstd = 1.0f / sqrtf(bnbuf[L.bn_var[l] + k] + 1e-5f);
                const float a = par(L.bn_g[l] + k) * invstd;
                bnS[k] = a;
                bnB[k] = par(L.bn_b[l] + k) - bnbuf[L.bn_mean[l] + k] * a;
            }
            __syncthreads();
            for (int t = tid; t < in * ET; t += MLP_THREADS) {
                const int k = t / ET;
                ain[t] = fmaf(ain[t], bnS[k], bnB[k]);
            }
            __syncthreads();
        }
        const int ldw = in + 1;  // +1: consecutive rows land in different banks
        int rc = MLP_WBUF / ldw;
        if (rc > no) rc = no;
        for (int o0 = 0; o0 < no; o0 += rc) {
            const int nr = min(rc, no - o0);
            const int wbase = L.w[l] + o0 * in;
            for (int t = tid; t < nr * in; t += MLP_THREADS) {
                const int r = t / in, k = t - r * in;
                wbuf[r * ldw + k] = par(wbase + t);
            }
            for (int r = tid; r < nr; r += MLP_THREADS) bbuf[r] = par(L.b[l] + o0 + r);
            __syncthreads();
            for (int item = tid; item < nr * (ET / 4); item += MLP_THREADS) {
                const int r = item % nr, eg = item / nr;
                const float* wr = wbuf + r * ldw;
                const float* xa = ain + eg * 4;
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
                for (int k = 0; k < in; ++k) {
                    const float wv = wr[k];
                    const float4 x = *reinterpret_cast<const float4*>(xa + k * ET);
                    acc.x = fmaf(wv, x.x, acc.x);
                    acc.y = fmaf(wv, x.y, acc.y);
                    acc.z = fmaf(wv, x.z, acc.z);
                    acc.w = fmaf(wv, x.w, acc.w);
                }
                const float bv = bbuf[r];
                float y[4] = {acc.x + bv, acc.y + bv, acc.z + bv, acc.w + bv};
                if (L.kind == DFD_POLICY_MUJOCO) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) y[q] = tanhf(y[q]);
                } else if (l < 2) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) y[q] = fmaxf(y[q], 0.f);
                }
                *reinterpret_cast<float4*>(aout + (o0 + r) * ET + eg * 4) = make_float4(y[0], y[1], y[2], y[3]);
            }
            __syncthreads();
        }
        float* t = ain; ain = aout; aout = t;
    }
    // ain now holds the head outputs [nout][ET]
    float* o = out + ((int64_t)m * E + e0) * L.nout;
    if (L.kind == DFD_POLICY_MUJOCO) {
        // MapContinuousToAction (torch_helpers.py:20-25): mean = y[:A], std = 0.55 + 0.45*y[A:]
        for (int t = tid; t < ne * L.nout; t += MLP_THREADS) {
            const int e = t / L.nout, j = t - e * L.nout;
            const float y = ain[j * ET + e];
            o[t] = j < L.A ? y : 0.55f + 0.45f * y;
        }
    } else {
        for (int e = tid; e < ne; e += MLP_THREADS) {
            float mx = -INFINITY;
            for (int j = 0; j < L.nout; ++j) mx = fmaxf(mx, ain[j * ET + e]);
            float s = 0.f;
            for (int j = 0; j < L.nout; ++j) s += expf(ain[j * ET + e] - mx);
            const float inv = 1.0f / s;
            for (int j = 0; j < L.nout; ++j) o[(int64_t)e * L.nout + j] = expf(ain[j * ET + e] - mx) * inv;
        }
    }
}


// ---------------------------------------------------------------------------------------------------------------
// Small-batch exact path (E <= 8 observations per member, P <= 12 288): the reference-faithful operating point -
// one observation per policy call (policies/*.py get_action) - is a per-member GEMV whose cost is the member's
// eps row.  One CTA per (member, observation tile): ALL 16-byte loads of theta and eps are issued before the first
// use - two cp.async.bulk (TMA bulk) copies per member, theta and the eps row, straight into shared memory: one
// exposed memory latency per member instead of one per layer, and no L1 line reservations (register-staged loads
// were capped by the ~28 KB of L1 left beside shared memory: 53 -> 28 us; bulk copies: see DESIGN.md).  The perturbed
// vector is built in place (two roundings, bit-identical to worker/worker.py:28), then the three layers run from shared memory
// with 4 lanes per output neuron (k interleaved, rotated by the neuron index so weight and activation reads are
// bank-conflict-free for the 64-wide layers) and a 2-step shuffle reduction.
// ---------------------------------------------------------------------------------------------------------------
template <int ET>
__global__ void __launch_bounds__(MLP_THREADS) mlp_forward_small_kernel(MlpLayout L, const float* __restrict__ replicas,
                                                                        int64_t stride, const float* __restrict__ theta,
                                                                        const float* __restrict__ bnbuf,
                                                                        const int64_t* __restrict__ idx,
                                                                        const int8_t* __restrict__ sign, float sigma,
                                                                        const float* __restrict__ obs, int E,
                                                                        float* __restrict__ out, int maxdim, int P4) {
    extern __shared__ __align__(16) float smem[];
    float* W = smem;                         // [P4] theta, then the member's perturbed flat parameter vector
    float* EPS = W + P4;                     // [P4] the member's table row
    float* actA = EPS + P4;                  // [ET][maxdim]
    float* actB = actA + ET * maxdim;
    float* bnS = actB + ET * maxdim;         // BN scale / shift of the current layer input (Discrete)
    float* bnB = bnS + maxdim;

    const int m = blockIdx.x, e0 = blockIdx.y * ET, ne = min(ET, E - e0);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float sg = sigma * (float)sign[m];
    const float* row = table_row_ptr(replicas, stride, idx[m]);
    const int P = (int)L.P, nv = P >> 2;
    // ---- the whole parameter vector: two bulk copies, completion on an mbarrier -------------------------------
    __shared__ __align__(8) uint64_t bar;
    const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&bar);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const uint32_t bytes = (uint32_t)nv * 16u;
        if (bytes) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(2u * bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             (uint32_t)__cvta_generic_to_shared(W)), "l"(theta), "r"(bytes), "r"(bar_a) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             (uint32_t)__cvta_generic_to_shared(EPS)), "l"(row), "r"(bytes), "r"(bar_a) : "memory");
        } else {
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_a) : "memory");
        }
    }
    // observations (row-major [e][k]) -> actA, zero rows for the padding of the tile
    const float* ob = obs + ((int64_t)m * E + e0) * L.K;
    for (int t = tid; t < ET * L.K; t += MLP_THREADS) {
        const int e = t / L.K, k = t - e * L.K;
        actA[e * maxdim + k] = e < ne ? ob[t] : 0.f;
    }
    float* ain = actA;
    float* aout = actB;
    __syncthreads();           // barrier initialised (and observations staged) before anyone waits on it
    {
        uint32_t ok, spins = 0;
        do {
            asm volatile(
                "{\n\t"
                ".reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t"
                "}\n"
                : "=r"(ok)
                : "r"(bar_a)
                : "memory");
            if (!ok && ++spins > (1u << 26)) __trap();
        } while (!ok);
    }
    for (int i = tid; i < nv; i += MLP_THREADS) {          // perturb in place: theta + sg * eps, two roundings
        const float4 a = reinterpret_cast<const float4*>(W)[i];
        const float4 e = reinterpret_cast<const float4*>(EPS)[i];
        reinterpret_cast<float4*>(W)[i] = make_float4(perturb1(a.x, sg, e.x), perturb1(a.y, sg, e.y), perturb1(a.z, sg, e.z),
                                                      perturb1(a.w, sg, e.w));
    }
    if (tid < P - 4 * nv) W[4 * nv + tid] = perturb1(theta[4 * nv + tid], sg, row[4 * nv + tid]);
    __syncthreads();

    for (int l = 0; l < 3; ++l) {
        const int in = L.in[l], no = L.out[l];
        if (L.kind == DFD_POLICY_DISCRETE) {
            // eval-mode BatchNorm1d on the layer input, gamma/beta perturbed per member, running stats shared
            for (int k = tid; k < in; k += MLP_THREADS) {
                const float invstd = 1.0f / sqrtf(bnbuf[L.bn_var[l] + k] + 1e-5f);
                const float a = W[L.bn_g[l] + k] * invstd;
                bnS[k] = a;
                bnB[k] = W[L.bn_b[l] + k] - bnbuf[L.bn_mean[l] + k] * a;
            }
            __syncthreads();
            for (int t = tid; t < in * ET; t += MLP_THREADS) {
                const int e = t / in, k = t - e * in;
                ain[e * maxdim + k] = fmaf(ain[e * maxdim + k], bnS[k], bnB[k]);
            }
            __syncthreads();
        }
        const int ks = lane & 3, o8 = lane >> 2;
        const int rot = (in & 31) == 0 ? 4 * o8 : 0;      // bank rotation for the 32-multiple widths
        const int kiter = (in + 3) >> 2;
        for (int o = warp * 8 + o8; o < ((no + 7) & ~7); o += (MLP_THREADS / 32) * 8) {
            float acc[ET];
#pragma unroll
            for (int e = 0; e < ET; ++e) acc[e] = 0.f;
            if (o < no) {
                const float* wr = W + L.w[l] + o * in;
                for (int j = 0; j < kiter; ++j) {
                    int k = ks + 4 * j + rot;
                    if (k >= in) k -= in;                  // rot < in and ks + 4j < in + 3
                    if (ks + 4 * j < in) {
                        const float wv = wr[k];
#pragma unroll
                        for (int e = 0; e < ET; ++e) acc[e] = fmaf(wv, ain[e * maxdim + k], acc[e]);
                    }
                }
            }
#pragma unroll
            for (int e = 0; e < ET; ++e) {
                acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 1);
                acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 2);
            }
            if (ks == 0 && o < no) {
                const float bv = W[L.b[l] + o];
#pragma unroll
                for (int e = 0; e < ET; ++e) {
                    float y = acc[e] + bv;
                    if (L.kind == DFD_POLICY_MUJOCO) y = tanhf(y);
                    else if (l < 2) y = fmaxf(y, 0.f);
                    aout[e * maxdim + o] = y;
                }
            }
        }
        __syncthreads();
        float* t = ain; ain = aout; aout = t;
    }
    // ain now holds the head outputs [e][nout]
    float* o = out + ((int64_t)m * E + e0) * L.nout;
    if (L.kind == DFD_POLICY_MUJOCO) {
        // MapContinuousToAction (torch_helpers.py:20-25): mean = y[:A], std = 0.55 + 0.45*y[A:]
        for (int t = tid; t < ne * L.nout; t += MLP_THREADS) {
            const int e = t / L.nout, j = t - e * L.nout;
            const float y = ain[e * maxdim + j];
            o[t] = j < L.A ? y : 0.55f + 0.45f * y;
        }
    } else {
        for (int e = tid; e < ne; e += MLP_THREADS) {
            float mx = -INFINITY;
            for (int j = 0; j < L.nout; ++j) mx = fmaxf(mx, ain[e * maxdim + j]);
            float s = 0.f;
            for (int j = 0; j < L.nout; ++j) s += expf(ain[e * maxdim + j] - mx);
            const float inv = 1.0f / s;
            for (int j = 0; j < L.nout; ++j) o[(int64_t)e * L.nout + j] = expf(ain[e * maxdim + j] - mx) * inv;
        }
    }
}

int dfd_atari_forward_impl(dfd_ctx* ctx, const dfd_policy_desc* desc, const dfd_table* table, const float* theta,
                           const float* bn_buffers, const int64_t* idx, const int8_t* sign, int n_members, float sigma,
                           const float* obs, int obs_per_member, float* out, cudaStream_t st);
int dfd_mlp_forward_tc_impl(dfd_ctx* ctx, const dfd_policy_desc* desc, const dfd_table* table, const float* theta,
                            const int64_t* idx, const int8_t* sign, int n_members, float sigma, const float* obs,
                            int obs_per_member, float* out, cudaStream_t st);

int dfd_mlp_forward_ws_impl(dfd_ctx* ctx, const dfd_policy_desc* desc, const dfd_table* table, const float* theta,
                            const int64_t* idx, const int8_t* sign, int n_members, float sigma, const float* obs,
                            int obs_per_member, float* out, int approx_tanh, cudaStream_t st);

int dfd_mlp_forward_direct_impl(dfd_ctx* ctx, const dfd_policy_desc* desc, const dfd_table* table, const float* theta,
                                const int64_t* idx, const int8_t* sign, int n_members, float sigma, const float* obs,
                                int obs_per_member, float* out, int approx_tanh, cudaStream_t st);
int dfd_mlp_forward_ws16_impl(dfd_ctx* ctx, const dfd_policy_desc* desc, const dfd_table* table, const float* theta,
                              const int64_t* idx, const int8_t* sign, int n_members, float sigma, const float* obs,
                              int obs_per_member, float* out, int approx_tanh, cudaStream_t st);
int dfd_mlp_forward_stream_impl(dfd_ctx* ctx, const dfd_policy_desc* desc, const dfd_table* table, const float* theta,
                                const int64_t* idx, const int8_t* sign, int n_members, float sigma, const float* obs,
                                int obs_per_member, float* out, int approx_tanh, cudaStream_t st);

extern "C" int dfd_policy_forward(dfd_ctx* ctx, const dfd_policy_desc* desc, const dfd_table* table,
                                  const float* theta, const float* bn_buffers, const int64_t* idx, const int8_t* sign,
                                  int n_members, float sigma, const float* obs, int obs_per_member, float* out,
                                  dfd_stream stream) {
    DFD_CHECK_ARG(ctx && desc && table && theta && idx && sign && obs && out, "dfd_policy_forward: NULL argument");
    if (n_members == 0 || obs_per_member == 0) return 0;
    DFD_CHECK_ARG(n_members > 0 && obs_per_member > 0, "dfd_policy_forward: negative sizes");
    cudaStream_t st = (cudaStream_t)stream;
    if (desc->kind == DFD_POLICY_ATARI)
        return dfd_atari_forward_impl(ctx, desc, table, theta, bn_buffers, idx, sign, n_members, sigma, obs,
                                      obs_per_member, out, st);
    DFD_CHECK_ARG(desc->kind == DFD_POLICY_MUJOCO || desc->kind == DFD_POLICY_DISCRETE,
                  "dfd_policy_forward: kind %d is not served by this entry point (IMPALA: dfd_impala_forward)", desc->kind);
    DFD_CHECK_ARG(desc->kind == DFD_POLICY_MUJOCO || bn_buffers, "dfd_policy_forward: Discrete needs bn_buffers");
    DFD_CHECK_ARG(desc->n_in > 0 && desc->h1 > 0 && desc->h2 > 0 && desc->n_act > 0, "dfd_policy_forward: bad dims");
    const MlpLayout L = make_layout(desc);
    DFD_CHECK_ARG(L.P < table->size, "dfd_policy_forward: num_params %lld >= table size", (long long)L.P);
    if (desc->precision >= 1 && desc->kind == DFD_POLICY_MUJOCO) {
        // 64x64 nets: warp-specialised resident-weight kernel; wide nets: warp-specialised streaming kernel;
        // anything else: the generic tcgen05 kernel
        // 64x64 nets with the table mirror registered: resident theta image + eps tiles straight from the table by TMA
        const int r6 = dfd_mlp_forward_ws16_impl(ctx, desc, table, theta, idx, sign, n_members, sigma, obs, obs_per_member, out,
                                                 desc->precision == 2 ? 1 : 0, st);
        if (r6 >= 0) return r6;
        const int rc = dfd_mlp_forward_ws_impl(ctx, desc, table, theta, idx, sign, n_members, sigma, obs, obs_per_member, out,
                                               desc->precision == 2 ? 1 : 0, st);
        if (rc >= 0) return rc;
        // wide nets with a sigma-scaled fp16 mirror of the table registered: weights straight from the table by TMA
        const int rd = dfd_mlp_forward_direct_impl(ctx, desc, table, theta, idx, sign, n_members, sigma, obs, obs_per_member,
                                                   out, desc->precision == 2 ? 1 : 0, st);
        if (rd >= 0) return rd;
        const int rs = dfd_mlp_forward_stream_impl(ctx, desc, table, theta, idx, sign, n_members, sigma, obs, obs_per_member,
                                                   out, desc->precision == 2 ? 1 : 0, st);
        if (rs >= 0) return rs;
        return dfd_mlp_forward_tc_impl(ctx, desc, table, theta, idx, sign, n_members, sigma, obs, obs_per_member, out, st);
    }
    int maxdim = L.K;
    if (L.h1 > maxdim) maxdim = L.h1;
    if (L.h2 > maxdim) maxdim = L.h2;
    if (L.nout > maxdim) maxdim = L.nout;
    DFD_CHECK_ARG(maxdim + 1 <= MLP_WBUF, "dfd_policy_forward: layer width %d too large", maxdim);
    DFD_CHECK_ARG(n_members <= 2147483647 / 1 && (obs_per_member + 3) / 4 <= 65535, "dfd_policy_forward: grid too large");
    // small batches of observations: the one-latency GEMV kernel (whole perturbed vector resident in shared memory)
    if (obs_per_member <= 8 && L.P <= 12288 && (((uintptr_t)theta) & 15) == 0 && !getenv("DFD_MLP_NO_SMALL")) {
        const int P4 = (int)((L.P + 3) / 4 * 4);
        dim3 grid(n_members, obs_per_member <= 2 ? (obs_per_member + 1) / 2 : (obs_per_member + 7) / 8);
        if (obs_per_member <= 2) {
            const size_t sm = ((size_t)2 * P4 + 2 * 2 * maxdim + 2 * maxdim) * sizeof(float);
            DFD_CUDA(cudaFuncSetAttribute(mlp_forward_small_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
            mlp_forward_small_kernel<2><<<grid, MLP_THREADS, sm, st>>>(L, table->replicas, table->replica_stride, theta, bn_buffers,
                                                                       idx, sign, sigma, obs, obs_per_member, out, maxdim, P4);
        } else {
            const size_t sm = ((size_t)2 * P4 + 2 * 8 * maxdim + 2 * maxdim) * sizeof(float);
            DFD_CUDA(cudaFuncSetAttribute(mlp_forward_small_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
            mlp_forward_small_kernel<8><<<grid, MLP_THREADS, sm, st>>>(L, table->replicas, table->replica_stride, theta, bn_buffers,
                                                                       idx, sign, sigma, obs, obs_per_member, out, maxdim, P4);
        }
        DFD_LAUNCHED(ctx);
        return 0;
    }
    const int ET = obs_per_member <= 4 ? 4 : 16;
    const size_t smem = ((size_t)2 * maxdim * ET + MLP_WBUF + 3 * (size_t)maxdim) * sizeof(float);
    DFD_CHECK_ARG(smem <= 227 * 1024, "dfd_policy_forward: layer width %d needs %zu B of shared memory", maxdim, smem);
    dim3 grid(n_members, (obs_per_member + ET - 1) / ET);
    if (ET == 4) {
        DFD_CUDA(cudaFuncSetAttribute(mlp_forward_fp32_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        mlp_forward_fp32_kernel<4><<<grid, MLP_THREADS, smem, st>>>(L, table->replicas, table->replica_stride, theta,
                                                                    bn_buffers, idx, sign, sigma, obs, obs_per_member,
                                                                    out, maxdim);
    } else {
        DFD_CUDA(cudaFuncSetAttribute(mlp_forward_fp32_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        mlp_forward_fp32_kernel<16><<<grid, MLP_THREADS, smem, st>>>(L, table->replicas, table->replica_stride, theta,
                                                                     bn_buffers, idx, sign, sigma, obs, obs_per_member,
                                                                     out, maxdim);
    }
    DFD_LAUNCHED(ctx);
    return 0;
}
