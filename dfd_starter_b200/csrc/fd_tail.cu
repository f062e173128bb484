// The whole learner step after the returns are known, as ONE kernel for short parameter vectors
// (fd_return mode, P <= 32 768): learner/finite_differences.py:40-49 (baseline, standardise, norms, g = sum w_i v_i),
// the sharded population's gradient exchange and dsgd/dynamic_sgd.py:18-39 + finite_differences.py:54-78
// (DSGD update, ||dtheta||, theta-history ring, distance rows).
//
// At C2 size (1024 rows x 6 092 columns, 25 MB that the forward just pulled through L2) the separate kernels
// prepare -> reduce -> [exchange] -> norm -> update are each launch / latency bound; chained inside one grid of
// co-resident CTAs the dependent latencies overlap and three to four launches disappear:
//   A  every CTA reduces the batch statistics itself (N doubles from L2) while the loads of ITS rows' indices, rewards
//      and prefix-sum entries are in flight, and computes the coefficients of its rows into shared memory;
//   B  the streaming reduction of its (column tile x row range) block, exactly like fd_reduce_kernel;
//   C  the last CTA of a column tile ("finisher") sums the row-range partials in fixed order;
//      world > 1: it pushes the tile into every peer's mailbox over NVLink, the last finisher publishes the step
//      flag, every finisher waits for all peers, sums the world slots in rank order and applies 1/std of all ranks'
//      rewards (same protocol and mailbox layout as xchg_allreduce.cu);
//   D  the finishers meet on a counter (they are resident: the grid never exceeds the resident capacity), read the
//      per-tile sums of squares in fixed order -> ||g||, and each updates its own 128 columns of theta, the ring
//      slot and the distance rows; the last one writes ||dtheta||.
// Everything is summed in a fixed order: run-to-run deterministic, bitwise identical on every rank.
#include "common.cuh"
#include <stdlib.h>

namespace {

constexpr int TL_THREADS = 256;
constexpr int TL_WARPS = TL_THREADS / 32;
constexpr int TL_U = 8;
constexpr int TL_MAXROWS = 256;        // rows per CTA (one coefficient per thread)
constexpr int TL_MAXTILES = 256;       // 128-column tiles: P <= 32 768
constexpr int TL_MAX_WORLD = 16;
constexpr size_t TL_XHDR = 256, TL_XFLAGS = 2 * TL_MAX_WORLD * 8;   // mailbox layout of xchg_allreduce.cu

__host__ __device__ inline size_t tl_slot_bytes(int64_t P) { return (size_t)((P + 3) / 4 * 4) * 4 + 256; }
// low-latency region of the mailbox (after the flag-protocol slots): every value travels as an 8-byte packet
// {fp32 bits | step number}, so a packet is its own "ready" flag: no system-scope fence, no separate flag store, no grid-wide
// ticket on the exchange path - one NVLink store latency between a finished tile and the peers that combine it
__host__ __device__ inline size_t tl_ll_slot_bytes(int64_t P) { return P <= (int64_t)TL_MAXTILES * 128 ? (size_t)((P + 3) / 4 * 4 + 16) * 8 : 0; }
__host__ __device__ inline size_t tl_ll_base(int64_t P, int world) { return TL_XHDR + TL_XFLAGS + 2 * (size_t)world * tl_slot_bytes(P); }
__device__ __forceinline__ void tl_st_ll2(char* p, uint32_t a, uint32_t b, uint32_t flag) {
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(flag), "r"(b), "r"(flag) : "memory");
}
__device__ __forceinline__ void tl_ld_ll2(const char* p, uint32_t& a, uint32_t& fa, uint32_t& b, uint32_t& fb) {
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(fa), "=r"(b), "=r"(fb) : "l"(p) : "memory");
}

struct TailParams {
    const float* replicas; int64_t stride; const double* prefix; int64_t P;
    const double* reward; const int64_t* idx; const int8_t* sign; int n, paired; double baseline; float sigma;
    int tiles, splits, rows_per_cta, R;
    float* grad; float* theta; float* hist; float* dist; int64_t hist_stride; int n_hist_valid, write_row; double step;
    float* update_size;
    float* partial; int64_t partial_stride; unsigned* tile_ctr; unsigned* glob; double* gpart; double* upart;
    char* const* mailboxes; int rank, world, ll;
};

__device__ __forceinline__ double tl_block_reduce(double v, double* sh, int op /*0 sum,1 min,2 max*/) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double y = __shfl_xor_sync(0xffffffffu, v, o);
        v = op == 0 ? v + y : (op == 1 ? fmin(v, y) : fmax(v, y));
    }
    __syncthreads();
    if (lane == 0) sh[w] = v;
    __syncthreads();
    double t = sh[0];
    for (int i = 1; i < TL_WARPS; ++i) t = op == 0 ? t + sh[i] : (op == 1 ? fmin(t, sh[i]) : fmax(t, sh[i]));
    return t;
}
__device__ __forceinline__ void tl_st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long tl_ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned tl_ld_volatile(const unsigned* p) {
    unsigned v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(TL_THREADS, 3) fd_tail_kernel(const TailParams p) {
    __shared__ float4 sm[TL_WARPS][32];
    __shared__ float4 xsm[TL_MAX_WORLD][32];      // sharded: the world's packets of this tile, one row per rank
    __shared__ const float* ptr_s[TL_MAXROWS];
    __shared__ float coef_s[TL_MAXROWS];
    __shared__ double sh[TL_WARPS];
    __shared__ double stats_s[5];
    __shared__ unsigned ticket_s;
    __shared__ float scal_s;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x, split = blockIdx.y;
    const int r_begin = split * p.rows_per_cta;
    const int r_end = min(r_begin + p.rows_per_cta, p.R);
    const int nrows = r_end - r_begin;
    const double sig = (double)p.sigma;
    const int nk = p.paired ? 2 : 1;
    dfd_grid_dependency_wait();        // launched with programmatic stream serialisation: rewards come from the predecessor
    unsigned long long xstep = 0;
    if (p.world > 1) xstep = *reinterpret_cast<volatile unsigned long long*>(p.mailboxes[p.rank] + 8);

    // ---- A: this thread's row (loads first: they fly while the statistics are reduced) ------------------------
    double x_[2] = {0.0, 0.0}, sg_[2] = {0.0, 0.0}, n2 = 1.0;
    int64_t id0 = 0;
    if (tid < nrows) {
        const int r = r_begin + tid;
        id0 = p.idx[r];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            if (k < nk) {
                x_[k] = p.reward[r + k * p.R] - p.baseline;
                sg_[k] = (double)p.sign[r + k * p.R];
            }
        }
        n2 = sig * sig * (p.prefix[id0 + p.P] - p.prefix[id0]);     // ||sigma*eps||^2 from the prefix sum of squares
    }
    // finite_differences.py:40,43: rewards - policy_reward, standardize_arr (population std, identity when std == 0)
    double s = 0.0, mn = 1e300, mx = -1e300;
    for (int i = tid; i < p.n; i += TL_THREADS) {
        const double x = p.reward[i] - p.baseline;
        s += x;
        mn = fmin(mn, x);
        mx = fmax(mx, x);
    }
    s = tl_block_reduce(s, sh, 0);
    mn = tl_block_reduce(mn, sh, 1);
    mx = tl_block_reduce(mx, sh, 2);
    const double mean = s / (double)p.n;
    double v = 0.0, v2 = 0.0;
    for (int i = tid; i < p.n; i += TL_THREADS) {
        const double x = p.reward[i] - p.baseline;
        v += (x - mean) * (x - mean);
        v2 += x * x;
    }
    v = tl_block_reduce(v, sh, 0);
    double sd = sqrt(v / (double)p.n);
    if (mn == mx) sd = 0.0;
    const bool deferred = p.world > 1;          // sharded: coefficients un-standardised, 1/std applied after the exchange
    if (deferred) {
        v2 = tl_block_reduce(v2, sh, 0);
        if (tid == 0) { stats_s[0] = s; stats_s[1] = v2; stats_s[2] = (double)p.n; stats_s[3] = mn; stats_s[4] = mx; }
        sd = 0.0;
    }
    const double inv_sd = sd == 0.0 ? 1.0 : 1.0 / sd;
    if (tid < nrows) {
        float c = 0.f;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            if (k < nk) {
                const double w = sd == 0.0 ? x_[k] : (x_[k] - mean) * inv_sd;
                c += (float)(w / n2 * sg_[k] * sig);
            }
        }
        ptr_s[tid] = table_row_ptr(p.replicas, p.stride, id0);
        coef_s[tid] = c;
    }
    __syncthreads();

    // ---- B: streaming reduction of this CTA's rows over its 128-column tile -------------------------------------
    const int64_t col0 = (int64_t)tile * 128 + 4 * lane;
    const bool active = col0 < p.P;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const int rpw = (nrows + TL_WARPS - 1) / TL_WARPS;
    const int w_begin = warp * rpw, w_end = min(w_begin + rpw, nrows);
    for (int j0 = w_begin; j0 < w_end; j0 += TL_U) {
        float4 x[TL_U];
        float c[TL_U];
#pragma unroll
        for (int u = 0; u < TL_U; ++u) {
            const int j = j0 + u;
            const bool ok = j < w_end;
            c[u] = ok ? coef_s[j] : 0.f;
            x[u] = (ok && active) ? ldg_stream_f4(ptr_s[j] + col0) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < TL_U; ++u) {
            acc.x = fmaf(c[u], x[u].x, acc.x);
            acc.y = fmaf(c[u], x[u].y, acc.y);
            acc.z = fmaf(c[u], x[u].z, acc.z);
            acc.w = fmaf(c[u], x[u].w, acc.w);
        }
    }
    sm[warp][lane] = acc;
    __syncthreads();
    float4 tot = make_float4(0.f, 0.f, 0.f, 0.f);
    const int64_t colc = (int64_t)tile * 128 + 4 * tid;     // valid for tid < 32
    if (tid < 32) {
#pragma unroll
        for (int w = 0; w < TL_WARPS; ++w) {
            const float4 t = sm[w][tid];
            tot.x += t.x; tot.y += t.y; tot.z += t.z; tot.w += t.w;
        }
        *reinterpret_cast<float4*>(p.partial + (int64_t)split * p.partial_stride + colc) = tot;
    }
    // ---- C: the last CTA of the tile becomes its finisher ------------------------------------------------------------
    __threadfence();
    __syncthreads();
    if (tid == 0) ticket_s = atomicAdd(p.tile_ctr + tile, 1u);
    __syncthreads();
    if (ticket_s != (unsigned)(p.splits - 1)) return;
    __threadfence();
    if (tid == 0) p.tile_ctr[tile] = 0;          // ready for the next launch
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tid < 32) {
        for (int sp = 0; sp < p.splits; ++sp) {
            const float4 t = __ldcg(reinterpret_cast<const float4*>(p.partial + (int64_t)sp * p.partial_stride + colc));
            g.x += t.x; g.y += t.y; g.z += t.z; g.w += t.w;
        }
    }
    if (p.world > 1 && !p.ll) {
        // flag protocol: push, one system fence per CTA, grid ticket, one flag per rank (the default beyond two ranks)
        // ---- exchange: push the tile (and, from tile 0, the statistics) to every peer, publish, wait, combine ---
        const int par = (int)(xstep & 1ull);
        const size_t slot = tl_slot_bytes(p.P);
        const int64_t n4 = (p.P + 3) / 4;
        const size_t my_off = TL_XHDR + TL_XFLAGS + ((size_t)par * p.world + p.rank) * slot;
        if (tid < 32 && colc < p.P) {
            for (int w = 0; w < p.world; ++w) {
                const int dst = (p.rank + w) % p.world;
                *reinterpret_cast<float4*>(p.mailboxes[dst] + my_off + 4 * (size_t)colc) = g;
            }
        }
        if (tile == 0 && tid >= 32 && tid < 37) {
            const double sv = stats_s[tid - 32];
            for (int w = 0; w < p.world; ++w)
                *reinterpret_cast<double*>(p.mailboxes[w] + my_off + 16 * (size_t)n4 + 8 * (tid - 32)) = sv;
        }
        __syncthreads();
        if (tid == 0) {
            __threadfence_system();
            ticket_s = atomicAdd(p.glob + 0, 1u);
        }
        __syncthreads();
        if (ticket_s == (unsigned)(p.tiles - 1)) {       // last finisher to have pushed: publish this rank's step
            if (tid == 0) __threadfence_system();
            __syncthreads();
            if (tid < p.world) {
                unsigned long long* f = reinterpret_cast<unsigned long long*>(p.mailboxes[tid] + TL_XHDR) + par * TL_MAX_WORLD + p.rank;
                tl_st_release_sys(f, xstep + 1ull);
            }
            if (tid == 0) {
                p.glob[0] = 0u;
                *reinterpret_cast<unsigned long long*>(p.mailboxes[p.rank] + 8) = xstep + 1ull;
            }
        }
        char* const mine = p.mailboxes[p.rank];
        if (tid < p.world) {
            const unsigned long long* f = reinterpret_cast<const unsigned long long*>(mine + TL_XHDR) + par * TL_MAX_WORLD + tid;
            unsigned spins = 0;
            while (tl_ld_acquire_sys(f) < xstep + 1ull)
                if (++spins > (1u << 26)) __trap();      // a missing peer must fault, not hang the GPU
        }
        __syncthreads();
        const size_t slots0 = TL_XHDR + TL_XFLAGS + (size_t)par * p.world * slot;
        if (tid < 32) {
            double sv[5] = {0.0, 0.0, 0.0, 1e300, -1e300};
            if (tid < p.world) {
                const double* st = reinterpret_cast<const double*>(mine + slots0 + (size_t)tid * slot + 16 * (size_t)n4);
#pragma unroll
                for (int k = 0; k < 5; ++k) sv[k] = __ldcg(st + k);
            }
            double ts = 0.0, tss = 0.0, tn = 0.0, tmn = 1e300, tmx = -1e300;
            for (int w = 0; w < p.world; ++w) {
                const double a0 = __shfl_sync(0xffffffffu, sv[0], w), a1 = __shfl_sync(0xffffffffu, sv[1], w);
                const double a2 = __shfl_sync(0xffffffffu, sv[2], w), a3 = __shfl_sync(0xffffffffu, sv[3], w);
                const double a4 = __shfl_sync(0xffffffffu, sv[4], w);
                if (a2 > 0.0) { ts += a0; tss += a1; tn += a2; tmn = fmin(tmn, a3); tmx = fmax(tmx, a4); }
            }
            double inv = 1.0;
            if (tn > 0.0 && tmn != tmx) {
                const double m2 = ts / tn;
                const double var = fmax(tss / tn - m2 * m2, 0.0);
                if (var > 0.0) inv = 1.0 / sqrt(var);
            }
            const float invf = (float)inv;
            g = make_float4(0.f, 0.f, 0.f, 0.f);
            if (colc < p.P) {
                for (int w = 0; w < p.world; ++w) {
                    const float4 t = __ldcg(reinterpret_cast<const float4*>(mine + slots0 + (size_t)w * slot + 4 * (size_t)colc));
                    g.x += t.x; g.y += t.y; g.z += t.z; g.w += t.w;
                }
            }
            g.x *= invf; g.y *= invf; g.z *= invf; g.w *= invf;
        }
    } else if (p.world > 1) {
        // ---- exchange over the low-latency packets: push the tile (and, from tile 0, the statistics) to every peer,
        //      poll the world slots of this tile in rank order, combine ---
        const int par = (int)(xstep & 1ull);
        const uint32_t flag = (uint32_t)(xstep + 1ull);
        const int64_t n4 = (p.P + 3) / 4;
        const size_t ll_slot = tl_ll_slot_bytes(p.P);
        const size_t ll0 = tl_ll_base(p.P, p.world) + (size_t)par * p.world * ll_slot;
        const size_t my_off = ll0 + (size_t)p.rank * ll_slot;
        if (tid < 32 && colc < p.P) {
            for (int w = 0; w < p.world; ++w) {
                char* dst = p.mailboxes[(p.rank + w) % p.world] + my_off + 8 * (size_t)colc;
                tl_st_ll2(dst, __float_as_uint(g.x), __float_as_uint(g.y), flag);
                tl_st_ll2(dst + 16, __float_as_uint(g.z), __float_as_uint(g.w), flag);
            }
        }
        if (tile == 0 && tid >= 32 && tid < 37) {
            const unsigned long long sv = (unsigned long long)__double_as_longlong(stats_s[tid - 32]);
            for (int w = 0; w < p.world; ++w)
                tl_st_ll2(p.mailboxes[w] + my_off + 8 * (size_t)(4 * n4) + 16 * (tid - 32), (uint32_t)sv, (uint32_t)(sv >> 32), flag);
        }
        const char* mine = p.mailboxes[p.rank];
        // every warp polls the packets of its own ranks (w, w + 8, ...), so the peers' latencies overlap instead of adding up;
        // the sum below still runs in rank order
        {
            const int64_t cl = (int64_t)tile * 128 + 4 * lane;
            for (int w = warp; w < p.world; w += TL_WARPS) {
                float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
                if (cl < p.P) {
                    const char* src = mine + ll0 + (size_t)w * ll_slot + 8 * (size_t)cl;
                    uint32_t a, fa, b2, fb, c, fc, d, fd;
                    unsigned spins = 0;
                    do {
                        tl_ld_ll2(src, a, fa, b2, fb);
                        tl_ld_ll2(src + 16, c, fc, d, fd);
                        if (++spins > (1u << 26)) __trap();          // a missing peer must fault, not hang the GPU
                    } while (fa != flag || fb != flag || fc != flag || fd != flag);
                    t = make_float4(__uint_as_float(a), __uint_as_float(b2), __uint_as_float(c), __uint_as_float(d));
                }
                xsm[w][lane] = t;
            }
        }
        __syncthreads();
        if (tid < 32) {
            double sv[5] = {0.0, 0.0, 0.0, 1e300, -1e300};
            if (tid < p.world) {
                const char* st = mine + ll0 + (size_t)tid * ll_slot + 8 * (size_t)(4 * n4);
#pragma unroll
                for (int k = 0; k < 5; ++k) {
                    uint32_t lo, flo, hi, fhi;
                    unsigned spins = 0;
                    do {
                        tl_ld_ll2(st + 16 * k, lo, flo, hi, fhi);
                        if (++spins > (1u << 26)) __trap();          // a missing peer must fault, not hang the GPU
                    } while (flo != flag || fhi != flag);
                    sv[k] = __longlong_as_double((long long)(((unsigned long long)hi << 32) | lo));
                }
            }
            double ts = 0.0, tss = 0.0, tn = 0.0, tmn = 1e300, tmx = -1e300;
            for (int w = 0; w < p.world; ++w) {
                const double a0 = __shfl_sync(0xffffffffu, sv[0], w), a1 = __shfl_sync(0xffffffffu, sv[1], w);
                const double a2 = __shfl_sync(0xffffffffu, sv[2], w), a3 = __shfl_sync(0xffffffffu, sv[3], w);
                const double a4 = __shfl_sync(0xffffffffu, sv[4], w);
                if (a2 > 0.0) { ts += a0; tss += a1; tn += a2; tmn = fmin(tmn, a3); tmx = fmax(tmx, a4); }
            }
            double inv = 1.0;
            if (tn > 0.0 && tmn != tmx) {
                const double m2 = ts / tn;
                const double var = fmax(tss / tn - m2 * m2, 0.0);
                if (var > 0.0) inv = 1.0 / sqrt(var);
            }
            const float invf = (float)inv;
            g = make_float4(0.f, 0.f, 0.f, 0.f);
            if (colc < p.P) {
                for (int w = 0; w < p.world; ++w) {          // rank order: bitwise identical sums on every rank
                    const float4 t = xsm[w][tid];
                    g.x += t.x; g.y += t.y; g.z += t.z; g.w += t.w;
                }
            }
            g.x *= invf; g.y *= invf; g.z *= invf; g.w *= invf;
        }
    }
    // ---- the tile's gradient + its sum of squares --------------------------------------------------------------------
    if (tid < 32) {
        const float gv[4] = {g.x, g.y, g.z, g.w};
        double q = 0.0;
        for (int k = 0; k < 4; ++k) {
            if (colc + k < p.P) {
                p.grad[colc + k] = gv[k];
                q += (double)gv[k] * (double)gv[k];
            }
        }
        q = warp_sum(q);
        if (tid == 0) p.gpart[tile] = q;
        sm[0][tid] = g;                     // the update below re-reads the tile from shared memory
    }
    // ---- D: finishers meet, ||g||, DSGD update of this tile's columns ---------------------------------------------------
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        atomicAdd(p.glob + 1, 1u);
        unsigned spins = 0;
        while (tl_ld_volatile(p.glob + 1) < (unsigned)p.tiles)
            if (++spins > (1u << 26)) __trap();
        __threadfence();
    }
    __syncthreads();
    if (tid < 32) {
        double t = 0.0;
        for (int i = tid; i < p.tiles; i += 32) t += __ldcg(p.gpart + i);
        t = warp_sum(t);
        if (tid == 0) {
            const double norm = sqrt(t);
            // a zero gradient trips `assert norm > 0` in the reference; here the step degenerates to no update
            scal_s = norm > 0.0 ? (float)(p.step / norm) : 0.f;
        }
    }
    __syncthreads();
    const float coef = scal_s;
    double dacc = 0.0;
    if (tid < 128) {
        const int64_t q = (int64_t)tile * 128 + tid;
        if (q < p.P) {
            const float gq = reinterpret_cast<const float*>(&sm[0][0])[tid];
            const float t_old = p.theta[q];
            // dynamic_sgd.py:27-37: the learner hands grad = -g, DSGD does p -= coef*grad (fp32 like torch)
            const float t_new = __fsub_rn(t_old, __fmul_rn(coef, -gq));
            const float d = __fsub_rn(t_old, t_new);
            dacc = (double)d * (double)d;
            for (int r0 = 0; r0 < p.n_hist_valid; r0 += 8) {
                float h[8];
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    h[j] = (r0 + j < p.n_hist_valid) ? __ldcg(p.hist + (int64_t)(r0 + j) * p.hist_stride + q) : 0.f;
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (r0 + j < p.n_hist_valid) __stcg(p.dist + (int64_t)(r0 + j) * p.hist_stride + q, __fsub_rn(h[j], t_new));
            }
            p.theta[q] = t_new;
            if (p.write_row >= 0) __stcg(p.hist + (int64_t)p.write_row * p.hist_stride + q, t_new);
        }
    }
    dacc = tl_block_reduce(dacc, sh, 0);
    if (tid == 0) {
        p.upart[tile] = dacc;
        __threadfence();
        ticket_s = atomicAdd(p.glob + 2, 1u);
    }
    __syncthreads();
    if (ticket_s != (unsigned)(p.tiles - 1)) return;
    __threadfence();
    if (tid < 32) {
        double t = 0.0;
        for (int i = tid; i < p.tiles; i += 32) t += __ldcg(p.upart + i);
        t = warp_sum(t);
        if (tid == 0) {
            *p.update_size = (float)sqrt(t);
            p.glob[1] = 0u;      // every finisher is past the meeting point (it arrived here after it)
            p.glob[2] = 0u;
            if (p.world > 1 && p.ll) *reinterpret_cast<unsigned long long*>(p.mailboxes[p.rank] + 8) = xstep + 1ull;   // one exchange step done
        }
    }
}

struct TailPlan {
    int tiles, splits, rows_per_cta, R;
    int64_t partial_stride;
};

bool tail_plan(const dfd_ctx* ctx, int64_t P, int n, int paired, TailPlan* out) {
    static int per_sm = -1;
    if (per_sm < 0) {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fd_tail_kernel, TL_THREADS, 0) != cudaSuccess) per_sm = 0;
    }
    if (per_sm < 1 || P <= 0 || P > (int64_t)TL_MAXTILES * 128 || n <= 0 || (paired && (n & 1))) return false;
    TailPlan t;
    t.R = paired ? n / 2 : n;
    t.tiles = (int)((P + 127) / 128);
    const int capacity = ctx->sm_count * per_sm;          // every CTA must be resident: finishers wait inside the kernel
    int splits = capacity / t.tiles;
    const int max_splits = (t.R + TL_WARPS * 4 - 1) / (TL_WARPS * 4);   // >= 4 rows per warp
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) return false;
    t.rows_per_cta = (t.R + splits - 1) / splits;
    if (t.rows_per_cta > TL_MAXROWS) return false;
    t.splits = (t.R + t.rows_per_cta - 1) / t.rows_per_cta;
    t.partial_stride = (int64_t)t.tiles * 128;
    *out = t;
    return true;
}

}  // namespace

// scratch: tile counters | 4 global counters | gpart | upart | partial sums
extern "C" size_t dfd_fd_step_fused_scratch_bytes(const dfd_ctx* ctx, int64_t n_params, int n_returns, int paired) {
    TailPlan t;
    if (!ctx || !tail_plan(ctx, n_params, n_returns, paired, &t)) return 0;       // 0: this shape is not served
    return dfd_align_up((size_t)(t.tiles + 8) * sizeof(unsigned), 256) + dfd_align_up((size_t)2 * t.tiles * sizeof(double), 256) +
           dfd_align_up((size_t)t.splits * t.partial_stride * sizeof(float), 256) + 256;
}

extern "C" int dfd_fd_step_fused(dfd_ctx* ctx, const dfd_table* table, int64_t n_params, const double* reward,
                                 const int64_t* idx, const int8_t* sign, int n_returns, int paired, double baseline,
                                 float sigma, float* theta, float* grad, double lr, double lr_scale, float* hist, float* dist,
                                 int64_t hist_stride, int n_hist_valid, int hist_write_row, float* update_size_out,
                                 void* const* mailboxes, int rank, int world, void* scratch, size_t scratch_bytes,
                                 dfd_stream stream) {
    DFD_CHECK_ARG(ctx && table && reward && idx && sign && theta && grad && update_size_out && scratch, "dfd_fd_step_fused: NULL argument");
    DFD_CHECK_ARG(n_params > 0 && n_params < table->size, "dfd_fd_step_fused: n_params out of range");
    DFD_CHECK_ARG(((uintptr_t)scratch & 255) == 0, "dfd_fd_step_fused: scratch must be 256-byte aligned");
    DFD_CHECK_ARG((n_hist_valid == 0 && hist_write_row < 0) || (hist && dist && hist_stride >= n_params), "dfd_fd_step_fused: history buffers missing");
    DFD_CHECK_ARG(world >= 1 && world <= TL_MAX_WORLD && rank >= 0 && rank < world, "dfd_fd_step_fused: rank %d / world %d", rank, world);
    DFD_CHECK_ARG(world == 1 || (mailboxes && paired), "dfd_fd_step_fused: the sharded form needs mailboxes and antithetic pairs");
    TailPlan t;
    DFD_CHECK_ARG(tail_plan(ctx, n_params, n_returns, paired, &t), "dfd_fd_step_fused: shape not served (P %lld, n %d); use the three-call path",
                  (long long)n_params, n_returns);
    DFD_CHECK_ARG(scratch_bytes >= dfd_fd_step_fused_scratch_bytes(ctx, n_params, n_returns, paired), "dfd_fd_step_fused: scratch too small");
    TailParams p;
    p.replicas = table->replicas; p.stride = table->replica_stride; p.prefix = table->prefix_sq; p.P = n_params;
    p.reward = reward; p.idx = idx; p.sign = sign; p.n = n_returns; p.paired = paired; p.baseline = baseline; p.sigma = sigma;
    p.tiles = t.tiles; p.splits = t.splits; p.rows_per_cta = t.rows_per_cta; p.R = t.R;
    p.grad = grad; p.theta = theta; p.hist = hist; p.dist = dist; p.hist_stride = hist_stride; p.n_hist_valid = n_hist_valid;
    p.write_row = hist_write_row;
    p.step = lr * sqrt((double)n_params) * lr_scale;      // dynamic_sgd.py:30 (python floats = fp64)
    p.update_size = update_size_out;
    char* s = (char*)scratch;
    p.tile_ctr = (unsigned*)s;
    p.glob = p.tile_ctr + t.tiles;
    s += dfd_align_up((size_t)(t.tiles + 8) * sizeof(unsigned), 256);
    p.gpart = (double*)s;
    p.upart = p.gpart + t.tiles;
    s += dfd_align_up((size_t)2 * t.tiles * sizeof(double), 256);
    p.partial = (float*)s;
    p.partial_stride = t.partial_stride;
    p.mailboxes = (char* const*)mailboxes; p.rank = rank; p.world = world;
    // measured on B200 (C2, graph replay): packets 60.7 us vs flags 62.0 us per step at 2 ranks, 71.7 vs 65.4 us at 8 ranks
    // (every finisher warp polling its ranks' packets loads the L2 the incoming stores land in) - packets for a pair of GPUs,
    // the flag protocol beyond; DFD_TAIL_FLAGS=1 / DFD_TAIL_LL=1 force one of them
    static const bool force_flags = getenv("DFD_TAIL_FLAGS") != nullptr, force_ll = getenv("DFD_TAIL_LL") != nullptr;
    p.ll = force_ll ? 1 : (force_flags ? 0 : (world <= 2 ? 1 : 0));
    dim3 grid(t.tiles, t.splits);
    DFD_CUDA(dfd_launch_pdl(fd_tail_kernel, grid, dim3(TL_THREADS), 0, (cudaStream_t)stream, p));
    DFD_LAUNCHED(ctx);
    return 0;
}
