// Strategy distances / novelty of many members at once.
//
// Reference: a member's "strategy" is its policy head evaluated on the zeta frames (`get_strategy`,
// policies/*.py), its novelty is the smallest distance to the strategies of the history
// (utils/math_helpers.py:147-155 `compute_strategy_novelty`, strategy/strategy_handler.py:25-30), the
// distance being one of utils/math_helpers.py:166-222.  The reference does this one member at a time in numpy;
// here one launch produces the [n_a, n_b] distance table (history vs history: sparse_history_manager.py:48-70)
// and / or the per-row minimum (novelty of every perturbed member).  The member strategies come from the batched
// perturbed forward (dfd_policy_forward with the zeta frames as observations).
//
// Bound: the tables are tiny (n_b * Z * W * 4 bytes stays in L2); the kernel is latency / issue bound and is not
// on the step's critical path.  Per-row terms follow the reference's fp32 operation order, the mean over the zeta
// frames is accumulated in fp64.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

// distance term of ONE zeta frame: a, b point at W contiguous floats
template <int KIND>
__device__ __forceinline__ float row_term(const float* __restrict__ a, const float* __restrict__ b, int W) {
    if (KIND == DFD_DIST_L2) {                       // math_helpers.py:166-170
        float s = 0.f;
        for (int w = 0; w < W; ++w) { float d = b[w] - a[w]; s += d * d; }
        return sqrtf(s);
    } else if (KIND == DFD_DIST_CATEGORICAL_TVD) {    // :218-221
        float s = 0.f;
        for (int w = 0; w < W; ++w) s += fabsf(a[w] - b[w]);
        return s;
    } else if (KIND == DFD_DIST_GAUSSIAN_WASSERSTEIN) {   // :200-216 (mean | std halves)
        const int n = W / 2;
        float m = 0.f, t = 0.f;
        for (int w = 0; w < n; ++w) { float d = a[w] - b[w]; m += d * d; }
        for (int w = n; w < 2 * n; ++w) t += a[w] + b[w] - 2.f * sqrtf(a[w] * b[w]);
        const float nrm = sqrtf(m);                  // np.square(np.linalg.norm(.)): sqrt, then square
        return nrm * nrm + t;
    } else if (KIND == DFD_DIST_CATEGORICAL_BHATTACHARYYA) {   // :194-197
        float bc = 0.f;
        for (int w = 0; w < W; ++w) bc += sqrtf(a[w] * b[w]);
        return -logf(bc + 1e-12f);
    } else {                                          // gaussian_bhattacharrya_dist :173-191 (as written there)
        const int n = W / 2;
        float mt = 0.f, d1 = 1.f, d2 = 1.f, d3 = 1.f;
        for (int w = 0; w < n; ++w) {
            const float s1 = a[n + w], s2 = b[n + w], s3 = (s1 + s2) / 2.f, d = a[w] - b[w];
            mt += d * d / s3;
            d1 *= s1; d2 *= s2; d3 *= s3;
        }
        return mt / 8.f + (d3 / sqrtf(d1 * d2)) / 4.f;
    }
}

template <int KIND>
__global__ void __launch_bounds__(kThreads) strategy_distance_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                                     int n_b, int Z, int W, double* __restrict__ dists,
                                                                     double* __restrict__ row_min, int exclude_diagonal) {
    extern __shared__ float a_sh[];                   // this row's [Z, W] strategy when it fits
    __shared__ double warp_min[kWarps];
    const int ia = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t per = (int64_t)Z * W;
    const float* a = A + ia * per;
    const bool staged = per * 4 <= 40 * 1024;
    if (staged) {
        for (int64_t i = threadIdx.x; i < per; i += kThreads) a_sh[i] = a[i];
        __syncthreads();
        a = a_sh;
    }
    double best = INFINITY;
    for (int ib = warp; ib < n_b; ib += kWarps) {
        const float* b = B + ib * per;
        double acc = 0.0;
        for (int z = lane; z < Z; z += 32) acc += (double)row_term<KIND>(a + (int64_t)z * W, b + (int64_t)z * W, W);
        acc = warp_sum(acc) / (double)Z;
        if (lane == 0 && dists) dists[(int64_t)ia * n_b + ib] = acc;
        if (!(exclude_diagonal && ib == ia)) best = acc < best ? acc : best;
    }
    if (lane == 0) warp_min[warp] = best;
    __syncthreads();
    if (threadIdx.x == 0 && row_min) {
        double m = warp_min[0];
        for (int w = 1; w < kWarps; ++w) m = warp_min[w] < m ? warp_min[w] : m;
        row_min[ia] = m;
    }
}

}  // namespace

extern "C" int dfd_strategy_distances(dfd_ctx* ctx, const float* a, int n_a, const float* b, int n_b, int n_frames,
                                      int width, int kind, double* dists, double* row_min, int exclude_diagonal,
                                      dfd_stream stream) {
    DFD_CHECK_ARG(ctx && a && b, "dfd_strategy_distances: NULL argument");
    DFD_CHECK_ARG(dists || row_min, "dfd_strategy_distances: neither dists nor row_min requested");
    DFD_CHECK_ARG(n_frames > 0 && width > 0 && n_b >= 0, "dfd_strategy_distances: bad shape");
    DFD_CHECK_ARG(kind >= DFD_DIST_L2 && kind <= DFD_DIST_GAUSSIAN_BHATTACHARYYA, "dfd_strategy_distances: unknown distance %d", kind);
    DFD_CHECK_ARG(!((kind == DFD_DIST_GAUSSIAN_WASSERSTEIN || kind == DFD_DIST_GAUSSIAN_BHATTACHARYYA) && (width & 1)),
                  "dfd_strategy_distances: gaussian strategies are mean | std halves, width %d is odd", width);
    if (n_a <= 0) return 0;
    const size_t per = (size_t)n_frames * width * 4;
    const size_t smem = per <= 40 * 1024 ? per : 0;
    cudaStream_t st = (cudaStream_t)stream;
#define DFD_LAUNCH_DIST(K)                                                                                             \
    case K:                                                                                                            \
        strategy_distance_kernel<K><<<n_a, kThreads, smem, st>>>(a, b, n_b, n_frames, width, dists, row_min, exclude_diagonal); \
        break;
    switch (kind) {
        DFD_LAUNCH_DIST(DFD_DIST_L2)
        DFD_LAUNCH_DIST(DFD_DIST_CATEGORICAL_TVD)
        DFD_LAUNCH_DIST(DFD_DIST_GAUSSIAN_WASSERSTEIN)
        DFD_LAUNCH_DIST(DFD_DIST_CATEGORICAL_BHATTACHARYYA)
        DFD_LAUNCH_DIST(DFD_DIST_GAUSSIAN_BHATTACHARYYA)
    }
#undef DFD_LAUNCH_DIST
    DFD_LAUNCHED(ctx);
    return 0;
}
