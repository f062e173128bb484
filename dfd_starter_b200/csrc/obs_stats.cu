// Observation normalisation and per-member observation statistics (second half).
//
// Reference: worker/agent.py:37-41 - inside the rollout loop every observation is, with probability
// obs_stats_update_chance, folded into the agent's WelfordRunningStat (utils/math_helpers.py:29-39) and then
// normalised, `clip((obs - mean) / std, -10, 10)`, with the learner-wide statistics before it reaches the policy; the
// agent's statistics travel with the return (`obs_stats_update`, worker/worker.py:56) and the learner merges them
// (math_helpers.py:68-87, host scalar work on 2K+1 numbers).
//
// Both kernels are elementwise / per-(member, feature) sequential fp32 in the reference's operation order with no FMA
// contraction, so results are BIT-IDENTICAL to numpy's.  Bound: HBM (one read, one write of the observations).
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256) normalize_obs_kernel(const float* __restrict__ obs, int64_t n, int width,
                                                            const float* __restrict__ mean, const float* __restrict__ stdv,
                                                            float clip, float* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(i % width);
        float v = __fdiv_rn(__fsub_rn(obs[i], mean[k]), stdv[k]);      // np.subtract(obs, mean) / std   (agent.py:40)
        v = fminf(fmaxf(v, -clip), clip);                             // np.clip(obs, -10, 10)           (agent.py:41)
        out[i] = v;
    }
}

// The reference WORKER normalises with statistics it deserialised from the learner's FDState (worker/worker.py:43,47):
// WelfordRunningStat.deserialize rebuilds them from a Python list, i.e. as float64 arrays, so `(obs - mean) / std` runs in
// fp64 there and is rounded to fp32 once, when the observation enters the policy (policies/policy.py:28).
__global__ void __launch_bounds__(256) normalize_obs_f64_kernel(const float* __restrict__ obs, int64_t n, int width,
                                                                const double* __restrict__ mean,
                                                                const double* __restrict__ stdv, float clip,
                                                                float* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(i % width);
        double v = __ddiv_rn(__dsub_rn((double)obs[i], mean[k]), stdv[k]);
        v = fmin(fmax(v, -(double)clip), (double)clip);
        out[i] = (float)v;
    }
}

// one thread per (member, feature): the member's selected observations folded in order, exactly as
// WelfordRunningStat.update does it (math_helpers.py:29-39)
__global__ void __launch_bounds__(128) member_obs_stats_kernel(const float* __restrict__ obs, const uint8_t* __restrict__ select,
                                                               int n_members, int obs_per_member, int width,
                                                               float* __restrict__ out) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)n_members * width) return;
    const int m = (int)(t / width), k = (int)(t % width);
    const float* o = obs + (int64_t)m * obs_per_member * width + k;
    const uint8_t* s = select + (int64_t)m * obs_per_member;
    float mean = 0.f, var = 0.f;
    int count = 0;
    for (int e = 0; e < obs_per_member; ++e) {
        if (!s[e]) continue;
        const int current = count;
        count += 1;
        const float delta = __fsub_rn(o[(int64_t)e * width], mean);          // :34
        const float delta_n = __fdiv_rn(delta, (float)count);                  // :35
        mean = __fadd_rn(mean, delta_n);                                       // :37
        var = __fadd_rn(var, __fmul_rn(__fmul_rn(delta, delta_n), (float)current));   // :38
    }
    float* row = out + (int64_t)m * (2 * width + 1);                           // serialize(): mean | variance | count  (:89-90)
    row[k] = mean;
    row[width + k] = var;
    if (k == 0) row[2 * width] = (float)count;
}

}  // namespace

extern "C" int dfd_normalize_obs(dfd_ctx* ctx, const float* obs, int64_t n_rows, int width, const void* mean,
                                 const void* stdv, int stats_f64, float clip, float* out, dfd_stream stream) {
    DFD_CHECK_ARG(ctx && obs && mean && stdv && out, "dfd_normalize_obs: NULL argument");
    DFD_CHECK_ARG(width > 0 && n_rows >= 0, "dfd_normalize_obs: bad shape");
    const int64_t n = n_rows * width;
    if (n == 0) return 0;
    int64_t grid = (n + 255) / 256;
    const int64_t cap = (int64_t)ctx->sm_count * 16;
    if (grid > cap) grid = cap;
    if (stats_f64)
        normalize_obs_f64_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(obs, n, width, (const double*)mean,
                                                                                   (const double*)stdv, clip, out);
    else
        normalize_obs_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(obs, n, width, (const float*)mean,
                                                                               (const float*)stdv, clip, out);
    DFD_LAUNCHED(ctx);
    return 0;
}

extern "C" int dfd_member_obs_stats(dfd_ctx* ctx, const float* obs, const uint8_t* select, int n_members, int obs_per_member,
                                    int width, float* out, dfd_stream stream) {
    DFD_CHECK_ARG(ctx && obs && select && out, "dfd_member_obs_stats: NULL argument");
    DFD_CHECK_ARG(width > 0 && obs_per_member >= 0 && n_members >= 0, "dfd_member_obs_stats: bad shape");
    const int64_t threads = (int64_t)n_members * width;
    if (threads == 0) return 0;
    member_obs_stats_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, (cudaStream_t)stream>>>(obs, select, n_members,
                                                                                                 obs_per_member, width, out);
    DFD_LAUNCHED(ctx);
    return 0;
}
