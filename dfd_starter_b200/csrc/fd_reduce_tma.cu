// The streaming reduction g[p] = sum_r coef[r] * row[r][p] with TMA bulk copies and shared-memory staging.
//
// CTA = 8 consumer warps + 1 producer warp.  The CTA owns a TILE-column slice of a range of rows.  One
// elected producer thread streams each row's slice (TILE*4 bytes, 16-byte aligned thanks to the table
// replicas) into a ring of shared-memory stages with cp.async.bulk (UBLKCP) completing on a "full" mbarrier;
// the 256 consumer threads each own 4 (or 1) columns, read them from the stage, FMA with the row's
// coefficient and release the stage on an "empty" mbarrier.  The ring keeps STAGES*TILE*4 bytes in flight
// per CTA without holding them in registers, so HBM latency is covered by resident bytes, not by warps.
// Row-split CTAs of a column slice are combined by the last CTA to finish, in fixed order (deterministic).
#include "common.cuh"

namespace {

constexpr int CONS_WARPS = 8;
constexpr int TMA_THREADS = (CONS_WARPS + 1) * 32;

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void bar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok, spins = 0;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        if (!ok && ++spins > (1u << 26)) __trap();   // a lost arrival must fault, not hang the GPU
    } while (!ok);
}

template <int VEC, int RB>   // VEC floats per consumer thread: 8 / 4 (TILE = 2048 / 1024) or 1 (TILE = 256); RB rows per ring stage
__global__ void __launch_bounds__(TMA_THREADS) fd_reduce_tma_kernel(const float* const* __restrict__ row_ptr,
                                                                    const float* __restrict__ row_coef, int n_rows,
                                                                    int64_t P, int rows_per_cta, int n_splits, int stages,
                                                                    float* __restrict__ partial, int64_t partial_stride,
                                                                    unsigned* __restrict__ counters,
                                                                    float* __restrict__ grad) {
    constexpr int TILE = 256 * VEC;
    constexpr int NV4 = VEC >= 4 ? VEC / 4 : 1;
    extern __shared__ __align__(128) unsigned char smraw[];
    float* stage = reinterpret_cast<float*>(smraw);                                       // [stages][RB][TILE]
    uint64_t* full = reinterpret_cast<uint64_t*>(smraw + (size_t)stages * RB * TILE * 4);  // [stages]
    uint64_t* empty = full + stages;                                                       // [stages]
    const float** rp_s = reinterpret_cast<const float**>(empty + stages);                  // [rows_per_cta]
    float* cf_s = reinterpret_cast<float*>(rp_s + rows_per_cta);                           // [rows_per_cta]
    __shared__ unsigned ticket_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t col0 = (int64_t)blockIdx.x * TILE;
    const int split = blockIdx.y;
    const int r_begin = split * rows_per_cta;
    const int nr = min(rows_per_cta, n_rows - r_begin);
    const int cols = (int)min((int64_t)TILE, P - col0);
    const uint32_t bytes = (uint32_t)((cols * 4 + 15) & ~15);
    const int n_blocks = (nr + RB - 1) / RB;

    for (int i = tid; i < nr; i += TMA_THREADS) {
        rp_s[i] = row_ptr[r_begin + i];
        cf_s[i] = row_coef[r_begin + i];
    }
    if (tid == 0) {
        for (int s = 0; s < stages; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s_u32(full + s)));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(empty + s)), "r"(CONS_WARPS));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    float acc[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[v] = 0.f;

    if (warp == CONS_WARPS) {
        // ---- producer: one thread issues every bulk copy of this CTA; a stage = RB rows under ONE transaction count
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;                       // parity of the ring pass the producer is filling
            for (int b = 0; b < n_blocks; ++b) {
                if (b >= stages) bar_wait(s_u32(empty + s), ph ^ 1u);
                const int r0 = b * RB;
                const int nb = min(RB, nr - r0);
                const uint32_t fb = s_u32(full + s);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fb), "r"(bytes * (uint32_t)nb) : "memory");
#pragma unroll
                for (int j = 0; j < RB; ++j) {
                    if (j < nb)
                        asm volatile(
                            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                s_u32(stage + ((size_t)s * RB + j) * TILE)),
                            "l"(rp_s[r0 + j] + col0), "r"(bytes), "r"(fb)
                            : "memory");
                }
                if (++s == stages) {
                    s = 0;
                    ph ^= 1u;
                }
            }
        }
    } else {
        // ---- consumers: thread t owns columns 4t..4t+3 (+1024 per further float4), rows in order -> the same sum order
        //      whatever RB / the ring depth
        int s = 0;
        uint32_t ph = 0;
        for (int b = 0; b < n_blocks; ++b) {
            bar_wait(s_u32(full + s), ph);
            const int r0 = b * RB;
#pragma unroll
            for (int j = 0; j < RB; ++j) {
                if (r0 + j < nr) {
                    const float c = cf_s[r0 + j];
                    const float* st = stage + ((size_t)s * RB + j) * TILE;
                    if (VEC >= 4) {
#pragma unroll
                        for (int q = 0; q < NV4; ++q) {
                            const float4 x = *reinterpret_cast<const float4*>(st + 1024 * q + 4 * tid);
                            acc[(4 * q + 0) % VEC] = fmaf(c, x.x, acc[(4 * q + 0) % VEC]);
                            acc[(4 * q + 1) % VEC] = fmaf(c, x.y, acc[(4 * q + 1) % VEC]);
                            acc[(4 * q + 2) % VEC] = fmaf(c, x.z, acc[(4 * q + 2) % VEC]);
                            acc[(4 * q + 3) % VEC] = fmaf(c, x.w, acc[(4 * q + 3) % VEC]);
                        }
                    } else {
                        acc[0] = fmaf(c, st[tid], acc[0]);
                    }
                }
            }
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s_u32(empty + s)) : "memory");
            if (++s == stages) {
                s = 0;
                ph ^= 1u;
            }
        }
    }

    // thread t's accumulators: columns col0 + 1024*q + 4*t + {0..3}  (VEC = 1: col0 + t)
    const bool cons = warp < CONS_WARPS;
    auto colof = [&](int v) -> int64_t { return VEC >= 4 ? col0 + 1024 * (v / 4) + 4 * tid + (v & 3) : col0 + tid; };
    if (n_splits == 1) {
        if (cons) {
#pragma unroll
            for (int v = 0; v < VEC; ++v)
                if (colof(v) < P) grad[colof(v)] = acc[v];
        }
        return;
    }
    if (cons) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) partial[(int64_t)split * partial_stride + colof(v)] = acc[v];
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) ticket_s = atomicAdd(counters + blockIdx.x, 1u);
    __syncthreads();
    if (ticket_s != (unsigned)(n_splits - 1)) return;
    __threadfence();
    if (cons) {
        float g[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) g[v] = 0.f;
        for (int s = 0; s < n_splits; ++s) {
#pragma unroll
            for (int v = 0; v < VEC; ++v) g[v] += __ldcg(partial + (int64_t)s * partial_stride + colof(v));
        }
#pragma unroll
        for (int v = 0; v < VEC; ++v)
            if (colof(v) < P) grad[colof(v)] = g[v];
    }
    if (tid == 0) counters[blockIdx.x] = 0;   // ready for the next launch
}

}  // namespace

struct TmaPlan {
    int vec, rb, stages, tiles, splits, rows_per_cta;
    int64_t partial_stride;
    size_t smem;
};

// DFD_TMA_PLAN="vec,rb,stages,ctas_per_sm,min_rows_per_cta" overrides the built-in choice (tuning sweeps: scripts/reduce_sweep.py)
static bool tma_plan_override(int* o) {
    static int cached[5];
    static int state = 0;       // 0 unknown, 1 none, 2 set
    if (state == 0) {
        const char* e = getenv("DFD_TMA_PLAN");
        state = (e && sscanf(e, "%d,%d,%d,%d,%d", cached, cached + 1, cached + 2, cached + 3, cached + 4) == 5) ? 2 : 1;
    }
    if (state != 2) return false;
    for (int i = 0; i < 5; ++i) o[i] = cached[i];
    return true;
}

TmaPlan dfd_tma_plan(int sm_count, int64_t P, int n_rows) {
    TmaPlan p;
    int o[5];
    int per_sm, min_rows;
    if (P >= 8192) {
        p.vec = 4; p.rb = 1; p.stages = 10; per_sm = 4; min_rows = 32;
    } else {
        p.vec = 1; p.rb = 1; p.stages = 16; per_sm = 8; min_rows = 32;
    }
    if (P >= 8192 && tma_plan_override(o)) {
        p.vec = o[0]; p.rb = o[1]; p.stages = o[2]; per_sm = o[3]; min_rows = o[4];
    }
    const int tile = 256 * p.vec;
    p.tiles = (int)((P + tile - 1) / tile);
    const int wave = sm_count * per_sm;              // CTAs resident at once
    // row splits: fill whole waves (a ragged last wave idles most of the chip for one CTA lifetime) while
    // keeping >= min_rows rows per CTA so the ring stays busy; among good candidates prefer the fewest splits
    const int max_splits = n_rows >= 2 * min_rows ? n_rows / min_rows : 1;
    int best = 1;
    double best_eff = 0.0;
    for (int sp = 1; sp <= max_splits && sp <= 64; ++sp) {
        const long total = (long)p.tiles * sp;
        const double eff = (double)total / (double)(((total + wave - 1) / wave) * wave);
        if (eff > best_eff + 0.03) {
            best_eff = eff;
            best = sp;
        }
    }
    int splits = best;
    p.rows_per_cta = (n_rows + splits - 1) / splits;
    p.splits = (n_rows + p.rows_per_cta - 1) / p.rows_per_cta;
    p.partial_stride = (int64_t)p.tiles * tile;
    p.smem = (size_t)p.stages * p.rb * tile * 4 + (size_t)p.stages * 16 + (size_t)p.rows_per_cta * 12 + 64;
    return p;
}

size_t dfd_tma_scratch_bytes(int sm_count, int64_t P, int n_rows) {
    const TmaPlan p = dfd_tma_plan(sm_count, P, n_rows);
    const size_t counters = dfd_align_up((size_t)p.tiles * sizeof(unsigned), 256);
    const size_t partial = p.splits > 1 ? (size_t)p.splits * p.partial_stride * sizeof(float) : 0;
    return counters + dfd_align_up(partial, 256) + 256;
}

template <int VEC, int RB>
static int launch_tma(dfd_ctx* ctx, const TmaPlan& p, const dfd_fd_rows* rows, int n_rows, int64_t P, float* grad,
                      float* partial, unsigned* counters, cudaStream_t st) {
    DFD_CUDA(cudaFuncSetAttribute(fd_reduce_tma_kernel<VEC, RB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
    dim3 grid(p.tiles, p.splits);
    fd_reduce_tma_kernel<VEC, RB><<<grid, TMA_THREADS, p.smem, st>>>(rows->row_ptr, rows->row_coef, n_rows, P, p.rows_per_cta,
                                                                     p.splits, p.stages, partial, p.partial_stride, counters,
                                                                     grad);
    DFD_LAUNCHED(ctx);
    return 0;
}

int dfd_fd_reduce_tma(dfd_ctx* ctx, const dfd_fd_rows* rows, int n_rows, int64_t P, float* grad, void* scratch,
                      cudaStream_t st) {
    const TmaPlan p = dfd_tma_plan(ctx->sm_count, P, n_rows);
    DFD_CHECK_ARG(p.smem <= 220 * 1024, "dfd_fd_reduce: %d rows per CTA do not fit the staging plan", p.rows_per_cta);
    unsigned* counters = (unsigned*)scratch;
    float* partial = (float*)((char*)scratch + dfd_align_up((size_t)p.tiles * sizeof(unsigned), 256));
#define DFD_TMA_CASE(V, B) \
    if (p.vec == V && p.rb == B) return launch_tma<V, B>(ctx, p, rows, n_rows, P, grad, partial, counters, st);
    DFD_TMA_CASE(1, 1) DFD_TMA_CASE(4, 1) DFD_TMA_CASE(4, 2) DFD_TMA_CASE(4, 4) DFD_TMA_CASE(8, 1) DFD_TMA_CASE(8, 2)
    DFD_TMA_CASE(8, 4) DFD_TMA_CASE(16, 1) DFD_TMA_CASE(16, 2)
#undef DFD_TMA_CASE
    DFD_CHECK_ARG(false, "dfd_fd_reduce: unsupported staging plan vec=%d rb=%d", p.vec, p.rb);
    return 0;
}
