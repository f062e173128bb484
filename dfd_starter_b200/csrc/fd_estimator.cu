// The estimator on the device: learner/finite_differences.py:24-114 and
// dsgd/dynamic_sgd.py:18-39, as three hot calls (see include/dfd_b200.h).
//
//   g = sum_i w_i * lambda_i / ||lambda_i||^2,   lambda_i = s_i*sigma*eps_i + d_{e_i}
//
// is evaluated as ONE streaming pass  g[p] = sum_r coef_r * row_r[p]  over a row
// list (table rows + the few theta-history distance rows).  The row norms come
// from an fp64 prefix sum of squares of the table (two loads per row) plus, for
// returns from older epochs only, eps_i . d_e dot products (a first pass over
// just those rows).  The reduction is HBM-bound: rows*P*4 bytes in, P*4 out.
#include "common.cuh"
#include <cuda_fp16.h>
#include <stdlib.h>

// ---------------------------------------------------------------------------
// (1) prepare: dots for delayed rows, then coefficients + row list
// ---------------------------------------------------------------------------
static const int DOT_CHUNK = 8192;  // columns per CTA in the dots pass
static const int DOT_THREADS = 256;

// layout of the prepare scratch (doubles): dot[n_returns] | dd[n_hist] | q_hist[n_hist] | counter
extern "C" size_t dfd_fd_prepare_scratch_bytes(int n_returns, int n_hist) {
    return dfd_align_up((size_t)(n_returns + 2 * n_hist + 2) * sizeof(double), 256);
}

// blockIdx.y < RB: dots of one table row with the distance rows of its returns (skipped when hist_row < 0).  RB = n_returns,
// or, with antithetic pairs ([plus | minus] batches: returns y and y + R share their table row by contract), RB = R: the row
// is streamed from HBM ONCE for both members - when the two returns come from the same epoch (the usual case: a worker
// evaluates a pair against one FDState) the dot itself is shared too.  Four 16-byte row loads in flight per thread.
// blockIdx.y >= RB: ||dist row||^2
__global__ void __launch_bounds__(DOT_THREADS) fd_dots_kernel(const float* __restrict__ replicas, int64_t stride,
                                                              const int64_t* __restrict__ idx,
                                                              const int32_t* __restrict__ hist_row, int n_returns, int R_pairs,
                                                              const float* __restrict__ dist, int64_t dist_stride,
                                                              int64_t P, double* __restrict__ out,
                                                              const __half* __restrict__ mirror, int64_t mstride, double inv_sigma) {
    __shared__ double sh[2][DOT_THREADS / 32];
    const int y = blockIdx.y;
    const int RB = R_pairs > 0 ? R_pairs : n_returns;
    const float* a;
    const float* b0;
    const float* b1 = nullptr;
    int h[2] = {-1, -1};
    if (y < RB) {
        h[0] = hist_row[y];
        if (R_pairs > 0) h[1] = hist_row[y + R_pairs];
        const int ha = h[0] >= 0 ? h[0] : h[1];
        if (ha < 0) return;
        a = table_row_ptr(replicas, stride, idx[y]);
        b0 = dist + (int64_t)ha * dist_stride;
        if (h[1] >= 0 && h[1] != ha) b1 = dist + (int64_t)h[1] * dist_stride;
    } else {
        a = b0 = dist + (int64_t)(y - RB) * dist_stride;
    }
    const int64_t c0 = (int64_t)blockIdx.x * DOT_CHUNK;
    const int64_t c1 = min(c0 + (int64_t)DOT_CHUNK, P);
    double acc0 = 0.0, acc1 = 0.0;
    if (mirror != nullptr && y < RB) {
        // the table row from the sigma-scaled fp16 mirror (dfd_table_build_scaled16: half the bytes of the fp32 row; the dot
        // enters ||lambda||^2 as a ~1e-3 correction, so the 2^-11 operand rounding moves the coefficient by ~1e-7
        // relative): 8 halves per 16-byte load against two float4 of the distance row (L2-resident)
        const int64_t id = idx[y];
        const __half* a16 = mirror + (id & 7) * mstride + (id & ~(int64_t)7);
        const int64_t v8 = c0 + ((c1 - c0) & ~(int64_t)7);
        constexpr int U8 = 4;
        for (int64_t cb = c0 + 8 * threadIdx.x; cb < v8; cb += 8 * DOT_THREADS * U8) {
            uint4 x[U8];
#pragma unroll
            for (int u = 0; u < U8; ++u) {
                const int64_t c = cb + (int64_t)u * 8 * DOT_THREADS;
                x[u] = make_uint4(0, 0, 0, 0);
                if (c < v8) asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(x[u].x), "=r"(x[u].y), "=r"(x[u].z), "=r"(x[u].w) : "l"(a16 + c));
            }
#pragma unroll
            for (int u = 0; u < U8; ++u) {
                const int64_t c = cb + (int64_t)u * 8 * DOT_THREADS;
                if (c < v8) {
                    const float2 e0 = __half22float2(*reinterpret_cast<const __half2*>(&x[u].x)), e1 = __half22float2(*reinterpret_cast<const __half2*>(&x[u].y));
                    const float2 e2 = __half22float2(*reinterpret_cast<const __half2*>(&x[u].z)), e3 = __half22float2(*reinterpret_cast<const __half2*>(&x[u].w));
                    const float4 ya = *reinterpret_cast<const float4*>(b0 + c), yb = *reinterpret_cast<const float4*>(b0 + c + 4);
                    float s = e0.x * ya.x;
                    s = fmaf(e0.y, ya.y, s); s = fmaf(e1.x, ya.z, s); s = fmaf(e1.y, ya.w, s);
                    s = fmaf(e2.x, yb.x, s); s = fmaf(e2.y, yb.y, s); s = fmaf(e3.x, yb.z, s); s = fmaf(e3.y, yb.w, s);
                    acc0 += (double)s;
                    if (b1) {
                        const float4 za = *reinterpret_cast<const float4*>(b1 + c), zb = *reinterpret_cast<const float4*>(b1 + c + 4);
                        float s1 = e0.x * za.x;
                        s1 = fmaf(e0.y, za.y, s1); s1 = fmaf(e1.x, za.z, s1); s1 = fmaf(e1.y, za.w, s1);
                        s1 = fmaf(e2.x, zb.x, s1); s1 = fmaf(e2.y, zb.y, s1); s1 = fmaf(e3.x, zb.z, s1); s1 = fmaf(e3.y, zb.w, s1);
                        acc1 += (double)s1;
                    }
                }
            }
        }
        for (int64_t c = v8 + threadIdx.x; c < c1; c += DOT_THREADS) {
            const double e = (double)__half2float(a16[c]);
            acc0 += e * (double)b0[c];
            if (b1) acc1 += e * (double)b1[c];
        }
        acc0 *= inv_sigma;
        acc1 *= inv_sigma;
    } else {
    // all bases are 16-byte aligned (replica rows by construction, dist rows because dist_stride % 4 == 0)
    const int64_t v1 = c0 + ((c1 - c0) & ~(int64_t)3);
    constexpr int U = 4;
    for (int64_t cb = c0 + 4 * threadIdx.x; cb < v1; cb += 4 * DOT_THREADS * U) {
        float4 x[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t c = cb + (int64_t)u * 4 * DOT_THREADS;
            x[u] = c < v1 ? ldg_stream_f4(a + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t c = cb + (int64_t)u * 4 * DOT_THREADS;
            if (c < v1) {
                const float4 y0 = *reinterpret_cast<const float4*>(b0 + c);
                float s = x[u].x * y0.x;
                s = fmaf(x[u].y, y0.y, s);
                s = fmaf(x[u].z, y0.z, s);
                s = fmaf(x[u].w, y0.w, s);
                acc0 += (double)s;
                if (b1) {
                    const float4 y1 = *reinterpret_cast<const float4*>(b1 + c);
                    float s1 = x[u].x * y1.x;
                    s1 = fmaf(x[u].y, y1.y, s1);
                    s1 = fmaf(x[u].z, y1.z, s1);
                    s1 = fmaf(x[u].w, y1.w, s1);
                    acc1 += (double)s1;
                }
            }
        }
    }
    for (int64_t c = v1 + threadIdx.x; c < c1; c += DOT_THREADS) {
        acc0 += (double)a[c] * (double)b0[c];
        if (b1) acc1 += (double)a[c] * (double)b1[c];
    }
    }
    acc0 = warp_sum(acc0);
    acc1 = warp_sum(acc1);
    if ((threadIdx.x & 31) == 0) {
        sh[0][threadIdx.x >> 5] = acc0;
        sh[1][threadIdx.x >> 5] = acc1;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t0 = 0.0, t1 = 0.0;
        for (int i = 0; i < DOT_THREADS / 32; ++i) {
            t0 += sh[0][i];
            t1 += sh[1][i];
        }
        if (y >= RB) {
            atomicAdd(out + n_returns + (y - RB), t0);
        } else {
            const int ha = h[0] >= 0 ? h[0] : h[1];
            if (h[0] >= 0) atomicAdd(out + y, t0);                                  // h[0] >= 0 implies ha == h[0]
            if (h[1] >= 0) atomicAdd(out + y + R_pairs, h[1] == ha ? t0 : t1);
        }
    }
}

// Epoch-grouped form of the dots pass: CTA (column chunk, history row e) keeps the chunk of the distance row d_e in SHARED
// memory and streams the table rows of every return whose epoch maps to e against it - the distance rows are read from L2
// once per chunk instead of once per return (at C5 size: 2.1 GB of L2 reads for 466 delayed returns, which bounded the
// per-return form above at ~255 us), so the pass is bound by the table rows from HBM.  With the sigma-scaled fp16 mirror
// registered the rows are read from it (half the bytes; the dot enters ||lambda||^2 as a ~1e-3 correction).  Also emits
// ||d_e||^2 of the chunk (the CTA holds it anyway).  out: dot[n_returns] | dd[n_hist].
static const int DOTE_CHUNK = 4096;          // columns per CTA: 16 KB of d_e in shared memory, several CTAs per SM
__global__ void __launch_bounds__(DOT_THREADS, 4) fd_dots_epoch_kernel(const float* __restrict__ replicas, int64_t stride,
                                                                    const int64_t* __restrict__ idx, const int32_t* __restrict__ hist_row,
                                                                    int n_returns, int R_pairs, const float* __restrict__ dist,
                                                                    int64_t dist_stride, int64_t P, double* __restrict__ out,
                                                                    const __half* __restrict__ mirror, int64_t mstride, double inv_sigma) {
    // d_e chunk in two planes: floats [8v, 8v+4) of vector v in plane 0, [8v+4, 8v+8) in plane 1 (64 bytes further than half
    // the chunk, so the planes start 16 banks apart): a lane of the fp16 path reads ONE float4 from each plane for its 8
    // table entries and consecutive lanes read consecutive float4 (in column order a lane's two float4 are 32 bytes apart
    // from its neighbour's: two-way bank conflicts on every read, 47 % of the kernel's shared-memory wavefronts in
    // prof_dots_c5); the fp32 path (float4 per lane) alternates planes lane by lane and stays conflict-free
    constexpr int DS_P1 = DOTE_CHUNK / 2 + 16;
    __shared__ __align__(16) float d_s[DOTE_CHUNK + 16];
    __shared__ double sh[DOT_THREADS / 32];
    auto ds_at = [](int c) { return ((c >> 2) & 1) * DS_P1 + (c >> 3) * 4 + (c & 3); };
    const int e = blockIdx.y;
    const int RB = R_pairs > 0 ? R_pairs : n_returns;
    const int64_t c0 = (int64_t)blockIdx.x * DOTE_CHUNK;
    const int nc = (int)min((int64_t)DOTE_CHUNK, P - c0);                 // columns of this chunk
    const int nc8 = nc & ~7;
    const float* de = dist + (int64_t)e * dist_stride + c0;               // 16-byte aligned: dist_stride % 4 == 0, c0 % 4096 == 0
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double dd = 0.0;
    for (int c = 4 * threadIdx.x; c < DOTE_CHUNK; c += 4 * DOT_THREADS) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c + 3 < nc) v = *reinterpret_cast<const float4*>(de + c);
        else {
            float t[4] = {0.f, 0.f, 0.f, 0.f};
            for (int k = 0; k < 4 && c + k < nc; ++k) t[k] = de[c + k];
            v = make_float4(t[0], t[1], t[2], t[3]);
        }
        *reinterpret_cast<float4*>(d_s + ds_at(c)) = v;
        dd += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
    }
    dd = warp_sum(dd);
    if (lane == 0) sh[warp] = dd;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < DOT_THREADS / 32; ++i) t += sh[i];
        atomicAdd(out + n_returns + e, t);
    }
    // the rows whose return(s) come from epoch row e are compacted into shared memory first (blocks of 1024 rows: one
    // coalesced pass over hist_row / idx instead of dependent global loads per candidate row), then the warps take the
    // matching rows round-robin; a warp covers the whole chunk of a row: 4096 columns = 32 lanes x 16 vectors of 8 elements,
    // all loads of a row in flight before the first use
    __shared__ int64_t l_idx[1024];
    __shared__ int l_row[1024];          // row | (h0 == e) << 30 | (h1 == e) << 31
    __shared__ int l_n;
    for (int yb = 0; yb < RB; yb += 1024) {
    __syncthreads();
    if (threadIdx.x == 0) l_n = 0;
    __syncthreads();
    for (int y = yb + threadIdx.x; y < min(yb + 1024, RB); y += DOT_THREADS) {
        const int h0 = hist_row[y], h1 = R_pairs > 0 ? hist_row[y + R_pairs] : -1;
        if (h0 == e || h1 == e) {
            const int slot = atomicAdd(&l_n, 1);
            l_row[slot] = y | (h0 == e ? (1 << 30) : 0) | (h1 == e ? (1 << 31) : 0);
            l_idx[slot] = idx[y];
        }
    }
    __syncthreads();
    const int ln = l_n;
    for (int li = warp; li < ln; li += DOT_THREADS / 32) {
        const int lr = l_row[li], y = lr & 0x3fffffff;
        const bool m0 = (lr >> 30) & 1, m1 = (lr >> 31) & 1;
        const int64_t id = l_idx[li];
        double acc = 0.0;
        if (mirror != nullptr) {
            const __half* a16 = mirror + (id & 7) * mstride + (id & ~(int64_t)7) + c0;      // c0 % 8 == 0: 16-byte aligned
#pragma unroll 1
            for (int part = 0; part < 2; ++part) {         // 8 loads of 16 bytes in flight per lane, twice (64 registers per thread cap: 4 CTAs per SM)
                uint4 x[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int c = 8 * (lane + 32 * (u + 8 * part));
                    x[u] = make_uint4(0, 0, 0, 0);
                    if (c < nc8) asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(x[u].x), "=r"(x[u].y), "=r"(x[u].z), "=r"(x[u].w) : "l"(a16 + c));
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int c = 8 * (lane + 32 * (u + 8 * part));
                    if (c < nc8) {
                        const float2 e0 = __half22float2(*reinterpret_cast<const __half2*>(&x[u].x)), e1 = __half22float2(*reinterpret_cast<const __half2*>(&x[u].y));
                        const float2 e2 = __half22float2(*reinterpret_cast<const __half2*>(&x[u].z)), e3 = __half22float2(*reinterpret_cast<const __half2*>(&x[u].w));
                        const float4 ya = *reinterpret_cast<const float4*>(d_s + (c >> 1)), yb = *reinterpret_cast<const float4*>(d_s + DS_P1 + (c >> 1));
                        float s = e0.x * ya.x;
                        s = fmaf(e0.y, ya.y, s); s = fmaf(e1.x, ya.z, s); s = fmaf(e1.y, ya.w, s);
                        s = fmaf(e2.x, yb.x, s); s = fmaf(e2.y, yb.y, s); s = fmaf(e3.x, yb.z, s); s = fmaf(e3.y, yb.w, s);
                        acc += (double)s;
                    }
                }
            }
            for (int c = nc8 + lane; c < nc; c += 32) acc += (double)__half2float(a16[c]) * (double)d_s[ds_at(c)];
            acc *= inv_sigma;
        } else {
            const float* a = table_row_ptr(replicas, stride, id) + c0;
            const int nc4 = nc & ~3;
#pragma unroll 1
            for (int half = 0; half < 4; ++half) {
                float4 x[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int c = 4 * (lane + 32 * (u + 8 * half));
                    x[u] = c < nc4 ? ldg_stream_f4(a + c) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int c = 4 * (lane + 32 * (u + 8 * half));
                    if (c < nc4) {
                        const float4 yv = *reinterpret_cast<const float4*>(d_s + ds_at(c));
                        float s = x[u].x * yv.x;
                        s = fmaf(x[u].y, yv.y, s); s = fmaf(x[u].z, yv.z, s); s = fmaf(x[u].w, yv.w, s);
                        acc += (double)s;
                    }
                }
            }
            for (int c = nc4 + lane; c < nc; c += 32) acc += (double)a[c] * (double)d_s[ds_at(c)];
        }
        acc = warp_sum(acc);
        if (lane == 0) {
            if (m0) atomicAdd(out + y, acc);
            if (m1) atomicAdd(out + y + R_pairs, acc);
        }
    }
    }
}

__device__ __forceinline__ double block_reduce_d(double v, double* sh, int op /*0 sum,1 min,2 max*/) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double y = __shfl_xor_sync(0xffffffffu, v, o);
        v = op == 0 ? v + y : (op == 1 ? fmin(v, y) : fmax(v, y));
    }
    __syncthreads();
    if (lane == 0) sh[w] = v;
    __syncthreads();
    double t = sh[0];
    for (int i = 1; i < nw; ++i) t = op == 0 ? t + sh[i] : (op == 1 ? fmin(t, sh[i]) : fmax(t, sh[i]));
    return t;
}

// Coefficients + row list.  One thread per table ROW (both members of an antithetic pair), a few CTAs; every
// CTA recomputes the batch statistics itself (N doubles from L2) so there is no grid-wide dependency:
//   standardize_arr (utils/math_helpers.py:127-134): population std, identity when std == 0;
//   ||lambda||^2 = sigma^2 (S[i+P]-S[i]) [+ 2 s sigma (eps.d) + ||d||^2 for returns from older epochs].
// The two dependent global round trips (rewards -> statistics, idx -> prefix entries) run concurrently.
static const int COEF_THREADS = 256;

__global__ void __launch_bounds__(COEF_THREADS) fd_coef_kernel(const float* __restrict__ replicas, int64_t stride,
                                                               const double* __restrict__ prefix, int64_t P,
                                                               const double* __restrict__ reward,
                                                               const int64_t* __restrict__ idx,
                                                               const int8_t* __restrict__ sign,
                                                               const int32_t* __restrict__ hist_row, int n, int paired,
                                                               double baseline, float sigma,
                                                               const float* __restrict__ dist, int64_t dist_stride,
                                                               int n_hist, const double* __restrict__ stats_reward,
                                                               int n_stats, const double* __restrict__ dots,
                                                               double* __restrict__ q_hist, unsigned* __restrict__ counter,
                                                               const float** row_ptr, float* __restrict__ row_coef,
                                                               double* __restrict__ stats_out) {
    __shared__ double sh[32];
    __shared__ unsigned ticket_s;
    const double* sr = stats_reward ? stats_reward : reward;
    const int ns = stats_reward ? n_stats : n;
    const double sig = (double)sigma;
    const double* dd = dots + n;
    const int R = paired ? n / 2 : n;
    const int r = blockIdx.x * COEF_THREADS + threadIdx.x;

    // this thread's row: issue its loads first, they fly while the statistics are reduced
    double x_[2] = {0.0, 0.0}, n2_[2] = {1.0, 1.0}, sg_[2] = {0.0, 0.0};
    int h_[2] = {-1, -1};
    int64_t id0 = 0;
    const int nk = paired ? 2 : 1;
    if (r < R) {
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            if (k < nk) {
                const int i = r + k * R;
                const int64_t id = idx[i];
                if (k == 0) id0 = id;
                x_[k] = reward[i] - baseline;
                sg_[k] = (double)sign[i];
                h_[k] = hist_row[i];
                n2_[k] = sig * sig * (prefix[id + P] - prefix[id]);
                if (h_[k] >= 0) n2_[k] += 2.0 * sg_[k] * sig * dots[i] + dd[h_[k]];
            }
        }
    }
    // finite_differences.py:40  rewards - policy_reward ; :43 standardize
    double s = 0.0, mn = 1e300, mx = -1e300;
    for (int i = threadIdx.x; i < ns; i += COEF_THREADS) {
        const double x = sr[i] - baseline;
        s += x;
        mn = fmin(mn, x);
        mx = fmax(mx, x);
    }
    const double mean = block_reduce_d(s, sh, 0) / (double)ns;
    mn = block_reduce_d(mn, sh, 1);
    mx = block_reduce_d(mx, sh, 2);
    const double sum_x = mean * (double)ns;
    double v = 0.0, v2 = 0.0;
    for (int i = threadIdx.x; i < ns; i += COEF_THREADS) {
        const double x = sr[i] - baseline;
        const double d = x - mean;
        v += d * d;
        v2 += x * x;
    }
    double sd = sqrt(block_reduce_d(v, sh, 0) / (double)ns);
    if (mn == mx) sd = 0.0;  // all rewards equal: numpy's std is exactly 0 and the array passes through
    double inv_sd = sd == 0.0 ? 1.0 : 1.0 / sd;
    // sharded population (dfd_fd_prepare_partial): this rank's statistics are published for the exchange step
    // and the coefficients stay un-standardised - in paired form the mean cancels and 1/std is applied after
    // the gradients of all ranks have been summed
    const bool deferred = stats_out != nullptr;
    if (deferred) {
        v2 = block_reduce_d(v2, sh, 0);
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            stats_out[0] = sum_x;
            stats_out[1] = v2;
            stats_out[2] = (double)ns;
            stats_out[3] = mn;
            stats_out[4] = mx;
        }
        sd = 0.0;          // w = x below
        inv_sd = 1.0;
    }

    if (r < R) {
        float c = 0.f;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            if (k < nk) {
                const double w = sd == 0.0 ? x_[k] : (x_[k] - mean) * inv_sd;
                const double winv = w / n2_[k];
                c += (float)(winv * sg_[k] * sig);
                if (h_[k] >= 0) atomicAdd(q_hist + h_[k], winv);
            }
        }
        row_ptr[r] = table_row_ptr(replicas, stride, id0);
        row_coef[r] = c;
    }
    if (n_hist == 0) return;
    // the last CTA to finish turns the per-epoch sums into the distance rows' coefficients
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) ticket_s = atomicAdd(counter, 1u);
    __syncthreads();
    if (ticket_s != gridDim.x - 1) return;
    __threadfence();
    for (int h = threadIdx.x; h < n_hist; h += COEF_THREADS) {
        row_ptr[R + h] = dist + (int64_t)h * dist_stride;
        row_coef[R + h] = (float)__ldcg(q_hist + h);
    }
    if (threadIdx.x == 0) *counter = 0;
}

extern "C" int dfd_fd_prepare(dfd_ctx* ctx, const dfd_table* table, int64_t n_params, const double* reward,
                              const int64_t* idx, const int8_t* sign, const int32_t* hist_row, int n_returns,
                              int paired, double baseline, float sigma, const float* dist, int64_t dist_stride,
                              int n_hist, const double* stats_reward, int n_stats, dfd_fd_rows* rows, void* scratch,
                              size_t scratch_bytes, dfd_stream stream) {
    DFD_CHECK_ARG(ctx && table && reward && idx && sign && hist_row && rows && scratch, "dfd_fd_prepare: NULL argument");
    DFD_CHECK_ARG(n_returns > 0, "dfd_fd_prepare: empty batch (the host returns 0 before calling, finite_differences.py:30-31)");
    DFD_CHECK_ARG(!paired || (n_returns % 2) == 0, "dfd_fd_prepare: paired mode needs an even number of returns");
    DFD_CHECK_ARG(n_hist >= 0 && n_hist <= 128, "dfd_fd_prepare: n_hist %d out of range (0..128)", n_hist);
    DFD_CHECK_ARG(n_hist == 0 || (dist && dist_stride >= n_params && dist_stride % 4 == 0 && ((uintptr_t)dist & 15) == 0),
                  "dfd_fd_prepare: dist must be 16-byte aligned with a stride that is a multiple of 4 floats");
    DFD_CHECK_ARG(n_params > 0 && n_params < table->size, "dfd_fd_prepare: n_params out of range");
    const int R = paired ? n_returns / 2 : n_returns;
    DFD_CHECK_ARG(rows->max_rows >= R + n_hist, "dfd_fd_prepare: row list too small (%d < %d)", rows->max_rows, R + n_hist);
    DFD_CHECK_ARG(scratch_bytes >= dfd_fd_prepare_scratch_bytes(n_returns, n_hist), "dfd_fd_prepare: scratch too small");
    cudaStream_t st = (cudaStream_t)stream;
    double* dots = (double*)scratch;
    double* q_hist = dots + n_returns + n_hist;
    unsigned* counter = (unsigned*)((double*)scratch + (dfd_fd_prepare_scratch_bytes(n_returns, n_hist) / sizeof(double) - 1));
    if (n_hist > 0) {
        DFD_CUDA(cudaMemsetAsync(dots, 0, (size_t)(n_returns + 2 * n_hist) * sizeof(double), st));
        dim3 grid((unsigned)((n_params + DOT_CHUNK - 1) / DOT_CHUNK), (unsigned)((paired ? R : n_returns) + n_hist));
        // with the sigma-scaled fp16 mirror of this table registered (dfd_table_build_scaled16) the pass reads half the bytes
        const bool m16 = ctx->scaled16 && ctx->scaled_src == table->replicas && ctx->scaled_sigma == sigma && sigma != 0.f &&
                         !getenv("DFD_DOTS_FP32");
        // long rows with many returns per epoch: the epoch-grouped form (distance-row chunk in shared memory, table rows
        // streamed against it); otherwise one CTA per (chunk, return)
        static const bool per_return = getenv("DFD_DOTS_PER_RETURN") != nullptr;
        if (!per_return && n_params >= 4 * DOTE_CHUNK && n_returns >= 4 * n_hist) {
            dim3 ge((unsigned)((n_params + DOTE_CHUNK - 1) / DOTE_CHUNK), (unsigned)n_hist);
            fd_dots_epoch_kernel<<<ge, DOT_THREADS, 0, st>>>(table->replicas, table->replica_stride, idx, hist_row, n_returns,
                                                             paired ? R : 0, dist, dist_stride, n_params, dots,
                                                             m16 ? (const __half*)ctx->scaled16 : nullptr, ctx->scaled16_stride, 1.0 / (double)sigma);
        } else
        fd_dots_kernel<<<grid, DOT_THREADS, 0, st>>>(table->replicas, table->replica_stride, idx, hist_row, n_returns,
                                                     paired ? R : 0, dist, dist_stride, n_params, dots,
                                                     m16 ? (const __half*)ctx->scaled16 : nullptr, ctx->scaled16_stride, 1.0 / (double)sigma);
        DFD_LAUNCHED(ctx);
    }
    fd_coef_kernel<<<(R + COEF_THREADS - 1) / COEF_THREADS, COEF_THREADS, 0, st>>>(
        table->replicas, table->replica_stride, table->prefix_sq, n_params, reward, idx, sign, hist_row, n_returns, paired,
        baseline, sigma, dist, dist_stride, n_hist, stats_reward, n_stats, dots, q_hist, counter, rows->row_ptr,
        rows->row_coef, nullptr);
    DFD_LAUNCHED(ctx);
    return 0;
}

extern "C" int dfd_fd_prepare_partial(dfd_ctx* ctx, const dfd_table* table, int64_t n_params, const double* reward,
                                      const int64_t* idx, const int8_t* sign, const int32_t* hist_row, int n_returns,
                                      double baseline, float sigma, dfd_fd_rows* rows, double* stats_out, void* scratch,
                                      size_t scratch_bytes, dfd_stream stream) {
    DFD_CHECK_ARG(ctx && table && reward && idx && sign && hist_row && rows && stats_out && scratch,
                  "dfd_fd_prepare_partial: NULL argument");
    DFD_CHECK_ARG(n_returns > 0 && (n_returns % 2) == 0, "dfd_fd_prepare_partial: needs a non-empty batch of antithetic pairs");
    DFD_CHECK_ARG(n_params > 0 && n_params < table->size, "dfd_fd_prepare_partial: n_params out of range");
    const int R = n_returns / 2;
    DFD_CHECK_ARG(rows->max_rows >= R, "dfd_fd_prepare_partial: row list too small (%d < %d)", rows->max_rows, R);
    DFD_CHECK_ARG(scratch_bytes >= dfd_fd_prepare_scratch_bytes(n_returns, 0), "dfd_fd_prepare_partial: scratch too small");
    double* dots = (double*)scratch;
    double* q_hist = dots + n_returns;
    unsigned* counter = (unsigned*)((double*)scratch + (dfd_fd_prepare_scratch_bytes(n_returns, 0) / sizeof(double) - 1));
    fd_coef_kernel<<<(R + COEF_THREADS - 1) / COEF_THREADS, COEF_THREADS, 0, (cudaStream_t)stream>>>(
        table->replicas, table->replica_stride, table->prefix_sq, n_params, reward, idx, sign, hist_row, n_returns, 1,
        baseline, sigma, nullptr, 0, 0, nullptr, 0, dots, q_hist, counter, rows->row_ptr, rows->row_coef, stats_out);
    DFD_LAUNCHED(ctx);
    return 0;
}

// ---------------------------------------------------------------------------
// (2) the streaming reduction  g[p] = sum_r coef[r] * row[r][p]
// ---------------------------------------------------------------------------
// CTA = RW warps; every warp owns the same 32*4*VPT-column tile and a different
// slice of the CTA's row range; lanes load 16-byte vectors (rows are 16-byte
// aligned by the replica construction), U rows in flight per lane.  Warps are
// combined through shared memory; row-split CTAs of a column tile are combined
// by the last CTA to finish (fixed order -> run-to-run deterministic).
static const int RED_WARPS = 8;
static const int RED_THREADS = RED_WARPS * 32;
static const int RED_U = 8;

template <int VPT>
__global__ void __launch_bounds__(RED_THREADS) fd_reduce_kernel(const float* const* __restrict__ row_ptr,
                                                                const float* __restrict__ row_coef, int n_rows,
                                                                int64_t P, int rows_per_cta, int n_splits,
                                                                float* __restrict__ partial, int64_t partial_stride,
                                                                unsigned* __restrict__ counters,
                                                                float* __restrict__ grad) {
    constexpr int TILE = 128 * VPT;
    __shared__ float4 sm[RED_WARPS][VPT][32];
    __shared__ unsigned ticket_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t col0 = (int64_t)blockIdx.x * TILE + 4 * lane;
    const int split = blockIdx.y;
    const int r_begin = split * rows_per_cta;
    const int r_end = min(r_begin + rows_per_cta, n_rows);

    float4 acc[VPT];
    bool active[VPT];
#pragma unroll
    for (int v = 0; v < VPT; ++v) {
        acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
        active[v] = (col0 + 128 * v) < P;
    }

    // the CTA's rows are split evenly over its warps; each warp walks its slice in batches of 32
    // (row pointer / coefficient fetched once per batch, one per lane, then broadcast by shuffle)
    const int rpw = (r_end - r_begin + RED_WARPS - 1) / RED_WARPS;
    const int w_begin = r_begin + warp * rpw;
    const int w_end = min(w_begin + rpw, r_end);
    for (int rb = w_begin; rb < w_end; rb += 32) {
        const int my = rb + lane;
        const float* pl = my < w_end ? row_ptr[my] : nullptr;
        const float cl = my < w_end ? row_coef[my] : 0.f;
        const int cnt = min(32, w_end - rb);
        for (int j0 = 0; j0 < cnt; j0 += RED_U) {
            float4 x[RED_U][VPT];
            float c[RED_U];
#pragma unroll
            for (int u = 0; u < RED_U; ++u) {
                const int j = j0 + u;
                const float* p = (const float*)__shfl_sync(0xffffffffu, (unsigned long long)pl, j & 31);
                c[u] = __shfl_sync(0xffffffffu, cl, j & 31);
                const bool ok = j < cnt;
#pragma unroll
                for (int v = 0; v < VPT; ++v) {
                    if (ok && active[v])
                        x[u][v] = ldg_stream_f4(p + col0 + 128 * v);
                    else
                        x[u][v] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
                if (!ok) c[u] = 0.f;
            }
#pragma unroll
            for (int u = 0; u < RED_U; ++u) {
#pragma unroll
                for (int v = 0; v < VPT; ++v) {
                    acc[v].x = fmaf(c[u], x[u][v].x, acc[v].x);
                    acc[v].y = fmaf(c[u], x[u][v].y, acc[v].y);
                    acc[v].z = fmaf(c[u], x[u][v].z, acc[v].z);
                    acc[v].w = fmaf(c[u], x[u][v].w, acc[v].w);
                }
            }
        }
    }
#pragma unroll
    for (int v = 0; v < VPT; ++v) sm[warp][v][lane] = acc[v];
    __syncthreads();
    // threads 0..32*VPT-1 combine the warps for one float4 each
    float4 tot = make_float4(0.f, 0.f, 0.f, 0.f);
    const int v_me = threadIdx.x >> 5, l_me = threadIdx.x & 31;
    const bool combiner = threadIdx.x < 32 * VPT;
    const int64_t colc = (int64_t)blockIdx.x * TILE + 128 * v_me + 4 * l_me;
    if (combiner) {
#pragma unroll
        for (int w = 0; w < RED_WARPS; ++w) {
            const float4 t = sm[w][v_me][l_me];
            tot.x += t.x;
            tot.y += t.y;
            tot.z += t.z;
            tot.w += t.w;
        }
    }
    if (n_splits == 1) {
        if (combiner && colc < P) {
            if (colc + 3 < P && (((uintptr_t)grad) & 15) == 0)
                *reinterpret_cast<float4*>(grad + colc) = tot;
            else {
                const float t[4] = {tot.x, tot.y, tot.z, tot.w};
                for (int k = 0; k < 4 && colc + k < P; ++k) grad[colc + k] = t[k];
            }
        }
        return;
    }
    // partial_stride is a multiple of 4 and covers whole tiles, so vector stores are always in range
    if (combiner) *reinterpret_cast<float4*>(partial + (int64_t)split * partial_stride + colc) = tot;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) ticket_s = atomicAdd(counters + blockIdx.x, 1u);
    __syncthreads();
    if (ticket_s != (unsigned)(n_splits - 1)) return;
    __threadfence();
    if (combiner) {
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int s = 0; s < n_splits; ++s) {
            const float4 t = __ldcg(reinterpret_cast<const float4*>(partial + (int64_t)s * partial_stride + colc));
            g.x += t.x;
            g.y += t.y;
            g.z += t.z;
            g.w += t.w;
        }
        const float t[4] = {g.x, g.y, g.z, g.w};
        for (int k = 0; k < 4 && colc + k < P; ++k) grad[colc + k] = t[k];
    }
    if (threadIdx.x == 0) counters[blockIdx.x] = 0;  // ready for the next launch
}

struct RedPlan {
    int vpt, tiles, splits, rows_per_cta;
    int64_t partial_stride;
};

static RedPlan red_plan(int sm_count, int64_t P, int n_rows) {
    RedPlan p;
    // wide rows: 256 columns per warp (1 KB contiguous per row per warp); narrow rows: 128
    p.vpt = (P >= (int64_t)sm_count * 8 * 256) ? 2 : 1;
    const int tile = 128 * p.vpt;
    p.tiles = (int)((P + tile - 1) / tile);
    // one full wave of resident CTAs (3 per SM at 69 registers): a second, ragged wave costs a whole CTA lifetime
    // (measured on B200: 25 MB launches are fastest in a single wave, larger ones with ~8 CTAs per SM queued)
    static const char* tgt = getenv("DFD_RED_TARGET");
    const int target = sm_count * (tgt ? atoi(tgt) : (p.tiles < sm_count ? 3 : 8));
    int splits = (target + p.tiles - 1) / p.tiles;
    const int max_splits = (n_rows + RED_WARPS * 4 - 1) / (RED_WARPS * 4);  // >= 4 rows per warp
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    p.rows_per_cta = (n_rows + splits - 1) / splits;
    p.splits = (n_rows + p.rows_per_cta - 1) / p.rows_per_cta;
    p.partial_stride = (int64_t)p.tiles * tile;
    return p;
}

size_t dfd_tma_scratch_bytes(int sm_count, int64_t P, int n_rows);
int dfd_fd_reduce_tma(dfd_ctx* ctx, const dfd_fd_rows* rows, int n_rows, int64_t P, float* grad, void* scratch,
                      cudaStream_t st);

// which implementation streams the rows: 1 = TMA bulk copies into a shared-memory ring (fd_reduce_tma.cu),
// 0 = register-staged 16-byte loads (below).  DFD_REDUCE_MODE=ldg|tma overrides for experiments.
static int reduce_mode(int64_t P, int n_rows) {
    static const char* e = getenv("DFD_REDUCE_MODE");
    if (e && e[0] == 'l') return 0;
    if (e && e[0] == 't') return 1;
    (void)n_rows;
    static const char* mp = getenv("DFD_TMA_MIN_P");
    return P >= (mp ? atoll(mp) : 16384) ? 1 : 0;   // measured on B200: the register-staged kernel wins below ~16 K columns
}

extern "C" size_t dfd_fd_reduce_scratch_bytes(const dfd_ctx* ctx, int64_t n_params, int n_rows) {
    if (!ctx || n_params <= 0 || n_rows <= 0) return 256;
    const RedPlan p = red_plan(ctx->sm_count, n_params, n_rows);
    const size_t counters = dfd_align_up((size_t)p.tiles * sizeof(unsigned), 256);
    const size_t partial = p.splits > 1 ? (size_t)p.splits * p.partial_stride * sizeof(float) : 0;
    const size_t a = counters + dfd_align_up(partial, 256) + 256;
    const size_t b = dfd_tma_scratch_bytes(ctx->sm_count, n_params, n_rows);
    return a > b ? a : b;   // either implementation may be chosen at launch
}

extern "C" int dfd_fd_reduce(dfd_ctx* ctx, const dfd_fd_rows* rows, int n_rows, int64_t n_params, float* grad,
                             void* scratch, size_t scratch_bytes, dfd_stream stream) {
    DFD_CHECK_ARG(ctx && rows && rows->row_ptr && rows->row_coef && grad && scratch, "dfd_fd_reduce: NULL argument");
    DFD_CHECK_ARG(n_rows > 0 && n_rows <= rows->max_rows, "dfd_fd_reduce: n_rows %d out of range", n_rows);
    DFD_CHECK_ARG(n_params > 0, "dfd_fd_reduce: n_params must be positive");
    DFD_CHECK_ARG(((uintptr_t)scratch & 255) == 0, "dfd_fd_reduce: scratch must be 256-byte aligned");
    DFD_CHECK_ARG(scratch_bytes >= dfd_fd_reduce_scratch_bytes(ctx, n_params, n_rows), "dfd_fd_reduce: scratch too small");
    if (reduce_mode(n_params, n_rows) == 1)
        return dfd_fd_reduce_tma(ctx, rows, n_rows, n_params, grad, scratch, (cudaStream_t)stream);
    const RedPlan p = red_plan(ctx->sm_count, n_params, n_rows);
    // the tile counters must be zero on entry; they are self-resetting, the caller zeroes scratch once at allocation
    unsigned* counters = (unsigned*)scratch;
    float* partial = (float*)((char*)scratch + dfd_align_up((size_t)p.tiles * sizeof(unsigned), 256));
    dim3 grid(p.tiles, p.splits);
    cudaStream_t st = (cudaStream_t)stream;
    if (p.vpt == 2)
        fd_reduce_kernel<2><<<grid, RED_THREADS, 0, st>>>(rows->row_ptr, rows->row_coef, n_rows, n_params,
                                                          p.rows_per_cta, p.splits, partial, p.partial_stride,
                                                          counters, grad);
    else
        fd_reduce_kernel<1><<<grid, RED_THREADS, 0, st>>>(rows->row_ptr, rows->row_coef, n_rows, n_params,
                                                          p.rows_per_cta, p.splits, partial, p.partial_stride,
                                                          counters, grad);
    DFD_LAUNCHED(ctx);
    return 0;
}

// ---------------------------------------------------------------------------
// (3) DSGD + theta-history / distance rows
// ---------------------------------------------------------------------------
static const int DSGD_THREADS = 256;
static const int DSGD_MAX_CTAS = 592;

extern "C" size_t dfd_dsgd_scratch_bytes(int64_t n_params) {
    (void)n_params;
    return dfd_align_up((size_t)(2 * DSGD_MAX_CTAS + 8) * sizeof(double), 256);
}

__global__ void __launch_bounds__(DSGD_THREADS) sumsq_partial_kernel(const float* __restrict__ g, int64_t P,
                                                                     double* __restrict__ partial) {
    __shared__ double sh[DSGD_THREADS / 32];
    double acc = 0.0;
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < P; p += (int64_t)gridDim.x * blockDim.x) {
        const double x = (double)g[p];
        acc += x * x;
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < DSGD_THREADS / 32; ++i) t += sh[i];
        partial[blockIdx.x] = t;
    }
}

// dynamic_sgd.py:27-37: the learner hands grad = -g, DSGD does p -= coef*grad, coef = lr*sqrt(P)*lr_scale/||grad||.
// theta_new = theta - fl(coef * (-g)) in fp32 like torch; dist rows and the ring slot are refreshed in the same pass.
__global__ void __launch_bounds__(DSGD_THREADS) dsgd_update_kernel(float* __restrict__ theta,
                                                                   const float* __restrict__ g, int64_t P, double step,
                                                                   const double* __restrict__ gnorm_partial,
                                                                   int n_partial, float* hist,
                                                                   float* dist, int64_t hist_stride,
                                                                   int n_hist_valid, int hist_write_row,
                                                                   double* __restrict__ upd_partial,
                                                                   unsigned* __restrict__ done_counter,
                                                                   float* __restrict__ update_size_out) {
    __shared__ double sh[DSGD_THREADS / 32];
    __shared__ float coef_s;
    __shared__ unsigned ticket_s;
    if (n_partial == 0) {
        // short gradients: every CTA sums g.g itself (same order in every CTA -> identical coefficient) instead of
        // paying a separate launch for the partial sums
        double a2 = 0.0;
        for (int64_t p = threadIdx.x; p < P; p += DSGD_THREADS) {
            const double x = (double)__ldcg(g + p);
            a2 += x * x;
        }
        a2 = warp_sum(a2);
        if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = a2;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int i = 0; i < DSGD_THREADS / 32; ++i) t += sh[i];
            const double norm = sqrt(t);
            coef_s = norm > 0.0 ? (float)(step / norm) : 0.f;
        }
    } else if (threadIdx.x < 32) {
        double t = 0.0;
        for (int i = threadIdx.x; i < n_partial; i += 32) t += gnorm_partial[i];
        t = warp_sum(t);
        if (threadIdx.x == 0) {
            const double norm = sqrt(t);
            // a zero gradient trips `assert norm > 0` in the reference; here the step degenerates to no update
            coef_s = norm > 0.0 ? (float)(step / norm) : 0.f;
        }
    }
    __syncthreads();
    const float coef = coef_s;
    double acc = 0.0;
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < P; p += (int64_t)gridDim.x * blockDim.x) {
        const float t_old = theta[p];
        const float t_new = __fsub_rn(t_old, __fmul_rn(coef, -g[p]));
        const float d = __fsub_rn(t_old, t_new);
        acc += (double)d * (double)d;
        // history rows: all loads of a batch are issued before its stores (hist is read AND written here)
        for (int r0 = 0; r0 < n_hist_valid; r0 += 8) {
            float h[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
                h[j] = (r0 + j < n_hist_valid) ? __ldcg(hist + (int64_t)(r0 + j) * hist_stride + p) : 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (r0 + j < n_hist_valid) __stcg(dist + (int64_t)(r0 + j) * hist_stride + p, __fsub_rn(h[j], t_new));
        }
        theta[p] = t_new;
        if (hist_write_row >= 0) __stcg(hist + (int64_t)hist_write_row * hist_stride + p, t_new);
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < DSGD_THREADS / 32; ++i) t += sh[i];
        upd_partial[blockIdx.x] = t;
        __threadfence();
        ticket_s = atomicAdd(done_counter, 1u);
    }
    __syncthreads();
    if (ticket_s != gridDim.x - 1) return;
    __threadfence();
    if (threadIdx.x < 32) {
        double t = 0.0;
        for (int i = threadIdx.x; i < (int)gridDim.x; i += 32) t += __ldcg(upd_partial + i);
        t = warp_sum(t);
        if (threadIdx.x == 0) {
            *update_size_out = (float)sqrt(t);
            *done_counter = 0;
        }
    }
}

extern "C" int dfd_dsgd_step(dfd_ctx* ctx, float* theta, const float* grad, int64_t n_params, double lr, double lr_scale,
                             float* hist, float* dist, int64_t hist_stride, int n_hist_valid, int hist_write_row,
                             float* update_size_out, void* scratch, size_t scratch_bytes, dfd_stream stream) {
    DFD_CHECK_ARG(ctx && theta && grad && update_size_out && scratch, "dfd_dsgd_step: NULL argument");
    DFD_CHECK_ARG(n_params > 0, "dfd_dsgd_step: n_params must be positive");
    DFD_CHECK_ARG(scratch_bytes >= dfd_dsgd_scratch_bytes(n_params), "dfd_dsgd_step: scratch too small");
    DFD_CHECK_ARG((n_hist_valid == 0 && hist_write_row < 0) || (hist && dist && hist_stride >= n_params),
                  "dfd_dsgd_step: history buffers missing");
    cudaStream_t st = (cudaStream_t)stream;
    int ctas = (int)((n_params + DSGD_THREADS - 1) / DSGD_THREADS);
    if (ctas > DSGD_MAX_CTAS) ctas = DSGD_MAX_CTAS;
    if (ctas < 1) ctas = 1;
    double* gpart = (double*)scratch;
    double* upart = gpart + DSGD_MAX_CTAS;
    unsigned* counter = (unsigned*)(upart + DSGD_MAX_CTAS);  // zero on entry (self-resetting)
    const bool fused_norm = n_params <= 32768;     // <= 128 KB of gradient: cheaper to re-sum per CTA than to launch
    if (!fused_norm) {
        sumsq_partial_kernel<<<ctas, DSGD_THREADS, 0, st>>>(grad, n_params, gpart);
        DFD_LAUNCHED(ctx);
    }
    // dynamic_sgd.py:30  coef = lr * sqrt(d) * lr_scale / norm  (python floats = fp64)
    const double step = lr * sqrt((double)n_params) * lr_scale;
    dsgd_update_kernel<<<ctas, DSGD_THREADS, 0, st>>>(theta, grad, n_params, step, gpart, fused_norm ? 0 : ctas, hist, dist, hist_stride,
                                                      n_hist_valid, hist_write_row, upart, counter, update_size_out);
    DFD_LAUNCHED(ctx);
    return 0;
}

// ---------------------------------------------------------------------------
// synthetic return (stand-in for the environment, which is outside this path):
// reward[m] = -mean_{e,j} (out[m,e,j] - target[j])^2, fp64, one CTA per member
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) synthetic_reward_kernel(const float* __restrict__ out, int per_member, int width,
                                                               const float* __restrict__ target,
                                                               double* __restrict__ reward) {
    __shared__ double sh[8];
    const float* o = out + (int64_t)blockIdx.x * per_member;
    float acc = 0.f;
    if ((((uintptr_t)o) & 15) == 0 && (width & 3) == 0) {
        for (int t = threadIdx.x; t < (per_member >> 2); t += 256) {
            const float4 v = ldg_stream_f4(o + 4 * t);
            const int j = (4 * t) % width;
            const float d0 = v.x - target[j], d1 = v.y - target[j + 1], d2 = v.z - target[j + 2], d3 = v.w - target[j + 3];
            acc += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
        }
        for (int t = (per_member & ~3) + threadIdx.x; t < per_member; t += 256) {
            const float d = o[t] - target[t % width];
            acc += d * d;
        }
    } else {
        for (int t = threadIdx.x; t < per_member; t += 256) {
            const float d = o[t] - target[t % width];
            acc += d * d;
        }
    }
    double a = warp_sum((double)acc);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < 8; ++i) t += sh[i];
        reward[blockIdx.x] = -t / (double)per_member;
    }
}

// short per-member outputs (<= 16 KB, 16-byte aligned, width a multiple of 4): one WARP per member, every lane issues
// all of its 16-byte loads before the first use -> one exposed memory latency, eight members per CTA
__global__ void __launch_bounds__(256) synthetic_reward_warp_kernel(const float* __restrict__ out, int per_member, int width,
                                                                    const float* __restrict__ target,
                                                                    double* __restrict__ reward, int n_members) {
    const int lane = threadIdx.x & 31, m = blockIdx.x * 8 + (threadIdx.x >> 5);
    dfd_grid_dependency_wait();        // launched with programmatic stream serialisation: `out` comes from the forward
    if (m >= n_members) return;
    const float* o = out + (int64_t)m * per_member;
    const int nq = per_member >> 2;          // <= 1024 quads: at most 32 per lane
    float acc = 0.f;
    for (int q0 = 0; q0 < nq; q0 += 32 * 16) {
        float4 v[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            const int q = q0 + u * 32 + lane;
            v[u] = q < nq ? ldg_stream_f4(o + 4 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            const int q = q0 + u * 32 + lane;
            if (q < nq) {
                const int j = (4 * q) % width;
                const float d0 = v[u].x - target[j], d1 = v[u].y - target[j + 1], d2 = v[u].z - target[j + 2],
                            d3 = v[u].w - target[j + 3];
                acc += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
            }
        }
    }
    const double a = warp_sum((double)acc);
    if (lane == 0) reward[m] = -a / (double)per_member;
}

extern "C" int dfd_synthetic_reward(dfd_ctx* ctx, const float* out, int n_members, int obs_per_member, int out_width,
                                    const float* target, double* reward, dfd_stream stream) {
    DFD_CHECK_ARG(ctx && out && target && reward, "dfd_synthetic_reward: NULL argument");
    if (n_members <= 0) return 0;
    const int64_t per_member = (int64_t)obs_per_member * out_width;
    if (per_member <= 4096 && (per_member & 3) == 0 && (out_width & 3) == 0 && (((uintptr_t)out) & 15) == 0) {
        DFD_CUDA(dfd_launch_pdl(synthetic_reward_warp_kernel, dim3((n_members + 7) / 8), dim3(256), 0, (cudaStream_t)stream, out,
                                (int)per_member, out_width, target, reward, n_members));
        DFD_LAUNCHED(ctx);
        return 0;
    }
    synthetic_reward_kernel<<<n_members, 256, 0, (cudaStream_t)stream>>>(out, obs_per_member * out_width, out_width,
                                                                         target, reward);
    DFD_LAUNCHED(ctx);
    return 0;
}
