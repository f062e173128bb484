// The one exchange step of the sharded learner: each GPU holds the partial gradient of its
// slice of the population; every rank needs the sum, scaled by 1 / std of ALL ranks' rewards
// (learner/finite_differences.py:43,49 over the whole batch).
//
// One process per GPU; every rank owns a MAILBOX in its HBM that all peers map through CUDA IPC.  ONE kernel
// per step and rank does the whole exchange over NVLink / NVSwitch peer stores - no NCCL call on this path:
//   1. push: the rank's partial gradient (P floats) and its reward statistics (sum, sum of squares, count,
//      min, max - 5 doubles) are written with 16-byte stores straight into slot [parity][rank] of EVERY
//      peer's mailbox (and its own);
//   2. publish: after a grid-wide arrival count the last CTA release-stores the step number into
//      flags[parity][rank] of every peer (system scope);
//   3. combine: every CTA acquire-polls its own flags until all ranks have published this step, then sums the
//      world slots IN RANK ORDER (bitwise identical result on every rank, run-to-run deterministic), applies
//      1 / std and writes the gradient the optimizer consumes.
// Slots and flags are double-buffered by step parity: a rank can run at most one step ahead of the slowest
// peer (it cannot publish step k+1 before it has combined step k), so parity k+2 never overwrites data a
// peer still reads.  The step counter lives in the mailbox, so the launch is CUDA-graph replayable.
#include "common.cuh"
#include <string.h>
#include <stdlib.h>

namespace {

constexpr int XCHG_THREADS = 256;
constexpr int XCHG_MAX_WORLD = 16;
constexpr size_t XCHG_HDR = 256;                       // bar_ctr (u32) | step_ctr (u64 at +8)
constexpr size_t XCHG_FLAGS = 2 * XCHG_MAX_WORLD * 8;  // u64 flags[2][16]

__host__ __device__ inline size_t xchg_slot_bytes(int64_t P) { return (size_t)((P + 3) / 4 * 4) * 4 + 256; }

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(XCHG_THREADS) xchg_allreduce_kernel(char* const* __restrict__ mailboxes, int rank,
                                                                      int world, int64_t P,
                                                                      const float* __restrict__ grad_partial,
                                                                      const double* __restrict__ stats5,
                                                                      float* __restrict__ grad_out,
                                                                      unsigned long long* __restrict__ prof) {
    // debugging aid (DFD_XCHG_PROF): accumulated nanoseconds of the four phases as seen by CTA 0
    unsigned long long t0 = 0, t1 = 0, t2 = 0, t3 = 0;
    auto now = []() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; };
    if (prof && blockIdx.x == 0 && threadIdx.x == 0) t0 = now();
    __shared__ unsigned ticket_s;
    __shared__ double inv_sd_s;
    char* const mine = mailboxes[rank];
    unsigned* bar_ctr = reinterpret_cast<unsigned*>(mine);
    unsigned long long* step_ctr = reinterpret_cast<unsigned long long*>(mine + 8);
    // every CTA reads the step before anyone can advance it (the advance happens after all CTAs arrived below)
    const unsigned long long step = *reinterpret_cast<volatile unsigned long long*>(step_ctr);
    const int par = (int)(step & 1ull);
    const size_t slot = xchg_slot_bytes(P);
    const int64_t n4 = (P + 3) / 4;
    const size_t my_slot_off = XCHG_HDR + XCHG_FLAGS + ((size_t)par * world + rank) * slot;

    // ---- 1. push ----------------------------------------------------------------------------------------
    for (int64_t i = (int64_t)blockIdx.x * XCHG_THREADS + threadIdx.x; i < n4; i += (int64_t)gridDim.x * XCHG_THREADS) {
        float4 v;
        if (4 * i + 3 < P) {
            v = *reinterpret_cast<const float4*>(grad_partial + 4 * i);
        } else {
            float t[4] = {0.f, 0.f, 0.f, 0.f};
            for (int k = 0; k < 4 && 4 * i + k < P; ++k) t[k] = grad_partial[4 * i + k];
            v = make_float4(t[0], t[1], t[2], t[3]);
        }
        for (int w = 0; w < world; ++w) {
            const int dst = (rank + w) % world;    // spread the peers over time: rank r starts at its own mailbox
            *reinterpret_cast<float4*>(mailboxes[dst] + my_slot_off + 16 * (size_t)i) = v;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x < 5) {
        const double s = stats5[threadIdx.x];
        for (int w = 0; w < world; ++w)
            *reinterpret_cast<double*>(mailboxes[w] + my_slot_off + 16 * (size_t)n4 + 8 * threadIdx.x) = s;
    }
    // ---- 2. publish --------------------------------------------------------------------------------------
    if (prof && blockIdx.x == 0 && threadIdx.x == 0) t1 = now();
    // the CTA barrier orders every thread's peer stores before thread 0's system-scope fence (fences are
    // cumulative), so ONE fence per CTA - not one per thread - makes the whole CTA's pushes visible before its
    // arrival is counted; the last CTA to arrive then publishes
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        ticket_s = atomicAdd(bar_ctr, 1u);
    }
    __syncthreads();
    if (ticket_s == gridDim.x - 1) {
        if (threadIdx.x == 0) __threadfence_system();
        __syncthreads();
        if (threadIdx.x < world) {
            unsigned long long* f = reinterpret_cast<unsigned long long*>(mailboxes[threadIdx.x] + XCHG_HDR) +
                                    par * XCHG_MAX_WORLD + rank;
            st_release_sys(f, step + 1ull);
        }
        if (threadIdx.x == 0) {
            *bar_ctr = 0u;
            *step_ctr = step + 1ull;
        }
    }
    // ---- 3. combine --------------------------------------------------------------------------------------
    if (prof && blockIdx.x == 0 && threadIdx.x == 0) t2 = now();
    if (threadIdx.x < world) {
        const unsigned long long* f = reinterpret_cast<const unsigned long long*>(mine + XCHG_HDR) + par * XCHG_MAX_WORLD +
                                      threadIdx.x;
        unsigned spins = 0;
        while (ld_acquire_sys(f) < step + 1ull) {
            if (++spins > (1u << 26)) __trap();    // a missing peer must fault, not hang the GPU
        }
    }
    __syncthreads();
    if (prof && blockIdx.x == 0 && threadIdx.x == 0) t3 = now();
    const size_t slots0 = XCHG_HDR + XCHG_FLAGS + (size_t)par * world * slot;
    if (threadIdx.x < 32) {
        // standardize_arr over the whole population (utils/math_helpers.py:127-134): population std,
        // identity when all rewards are equal.  Lane w loads rank w's five statistics (all loads in flight
        // together), lane 0 combines them in rank order.
        double v[5] = {0.0, 0.0, 0.0, 1e300, -1e300};
        if (threadIdx.x < world) {
            const double* st = reinterpret_cast<const double*>(mine + slots0 + (size_t)threadIdx.x * slot + 16 * (size_t)n4);
#pragma unroll
            for (int k = 0; k < 5; ++k) v[k] = __ldcg(st + k);
        }
        double s = 0.0, ss = 0.0, n = 0.0, mn = 1e300, mx = -1e300;
        for (int w = 0; w < world; ++w) {
            const double a0 = __shfl_sync(0xffffffffu, v[0], w), a1 = __shfl_sync(0xffffffffu, v[1], w);
            const double a2 = __shfl_sync(0xffffffffu, v[2], w), a3 = __shfl_sync(0xffffffffu, v[3], w);
            const double a4 = __shfl_sync(0xffffffffu, v[4], w);
            if (a2 > 0.0) {
                s += a0;
                ss += a1;
                n += a2;
                mn = fmin(mn, a3);
                mx = fmax(mx, a4);
            }
        }
        double inv = 1.0;
        if (n > 0.0 && mn != mx) {
            const double mean = s / n;
            const double var = fmax(ss / n - mean * mean, 0.0);
            if (var > 0.0) inv = 1.0 / sqrt(var);
        }
        if (threadIdx.x == 0) inv_sd_s = inv;
    }
    __syncthreads();
    const float inv_sd = (float)inv_sd_s;
    for (int64_t i = (int64_t)blockIdx.x * XCHG_THREADS + threadIdx.x; i < n4; i += (int64_t)gridDim.x * XCHG_THREADS) {
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int w = 0; w < world; ++w) {
            const float4 t = __ldcg(reinterpret_cast<const float4*>(mine + slots0 + (size_t)w * slot + 16 * (size_t)i));
            g.x += t.x;
            g.y += t.y;
            g.z += t.z;
            g.w += t.w;
        }
        const float t[4] = {g.x * inv_sd, g.y * inv_sd, g.z * inv_sd, g.w * inv_sd};
        for (int k = 0; k < 4 && 4 * i + k < P; ++k) grad_out[4 * i + k] = t[k];
    }
    if (prof && blockIdx.x == 0 && threadIdx.x == 0) {
        const unsigned long long t4 = now();
        prof[0] += t1 - t0; prof[1] += t2 - t1; prof[2] += t3 - t2; prof[3] += t4 - t3; prof[4] += 1;
    }
}

// Two-phase form for LONG parameter vectors on many ranks: the one-phase kernel above pushes every rank's whole partial to
// every peer (world * P floats leave each GPU - fine at 24 KB, 37 MB per rank at C5 size on 8 GPUs).  Here rank r OWNS slice
// r of the vector: (1) reduce-scatter - every rank pushes slice d of its partial to rank d only (and its 5 statistics to
// everybody), (2) the owner sums the world contributions of its slice in rank order, applies 1 / std and pushes the reduced
// slice to every peer's result region, (3) all-gather - every rank copies the other slices out of its own result region.
// ~2 P floats leave each GPU instead of world * P; two flag rounds instead of one.  Same determinism: every element is
// summed by exactly one rank, in rank order, and every rank receives those bits.
__global__ void __launch_bounds__(XCHG_THREADS) xchg_allreduce_rs_kernel(char* const* __restrict__ mailboxes, int rank, int world, int64_t P,
                                                                         const float* __restrict__ grad_partial,
                                                                         const double* __restrict__ stats5, float* __restrict__ grad_out,
                                                                         size_t off2) {
    __shared__ unsigned ticket_s;
    __shared__ double inv_sd_s;
    char* const mine = mailboxes[rank];
    unsigned* bar_ctr = reinterpret_cast<unsigned*>(mine);
    unsigned* bar_ctr2 = reinterpret_cast<unsigned*>(mine + 4);
    unsigned long long* step_ctr = reinterpret_cast<unsigned long long*>(mine + 8);
    const unsigned long long step = *reinterpret_cast<volatile unsigned long long*>(step_ctr);
    const int par = (int)(step & 1ull);
    const size_t slot = xchg_slot_bytes(P);
    const int64_t n4 = (P + 3) / 4;
    const int64_t per = (n4 + world - 1) / world;
    const size_t slots0 = XCHG_HDR + XCHG_FLAGS + (size_t)par * world * slot;
    const size_t my_slot_off = slots0 + (size_t)rank * slot;
    const size_t res_off = off2 + XCHG_FLAGS + (size_t)par * slot;
    const int64_t gstride = (int64_t)gridDim.x * XCHG_THREADS, g0 = (int64_t)blockIdx.x * XCHG_THREADS + threadIdx.x;
    // ---- 1. reduce-scatter push: element i goes to its owner only ----
    for (int64_t i = g0; i < n4; i += gstride) {
        float4 v;
        if (4 * i + 3 < P) v = *reinterpret_cast<const float4*>(grad_partial + 4 * i);
        else {
            float t[4] = {0.f, 0.f, 0.f, 0.f};
            for (int k = 0; k < 4 && 4 * i + k < P; ++k) t[k] = grad_partial[4 * i + k];
            v = make_float4(t[0], t[1], t[2], t[3]);
        }
        *reinterpret_cast<float4*>(mailboxes[(int)(i / per)] + my_slot_off + 16 * (size_t)i) = v;
    }
    if (blockIdx.x == 0 && threadIdx.x < 5) {
        const double sv = stats5[threadIdx.x];
        for (int w = 0; w < world; ++w) *reinterpret_cast<double*>(mailboxes[w] + my_slot_off + 16 * (size_t)n4 + 8 * threadIdx.x) = sv;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        ticket_s = atomicAdd(bar_ctr, 1u);
    }
    __syncthreads();
    if (ticket_s == gridDim.x - 1) {
        if (threadIdx.x == 0) __threadfence_system();
        __syncthreads();
        if (threadIdx.x < world)
            st_release_sys(reinterpret_cast<unsigned long long*>(mailboxes[threadIdx.x] + XCHG_HDR) + par * XCHG_MAX_WORLD + rank, step + 1ull);
        if (threadIdx.x == 0) *bar_ctr = 0u;
    }
    if (threadIdx.x < world) {
        const unsigned long long* f = reinterpret_cast<const unsigned long long*>(mine + XCHG_HDR) + par * XCHG_MAX_WORLD + threadIdx.x;
        unsigned spins = 0;
        while (ld_acquire_sys(f) < step + 1ull)
            if (++spins > (1u << 26)) __trap();    // a missing peer must fault, not hang the GPU
    }
    __syncthreads();
    if (threadIdx.x < 32) {        // standardize_arr over the whole population (utils/math_helpers.py:127-134), as in the one-phase kernel
        double v[5] = {0.0, 0.0, 0.0, 1e300, -1e300};
        if (threadIdx.x < world) {
            const double* st = reinterpret_cast<const double*>(mine + slots0 + (size_t)threadIdx.x * slot + 16 * (size_t)n4);
#pragma unroll
            for (int k = 0; k < 5; ++k) v[k] = __ldcg(st + k);
        }
        double s = 0.0, ss = 0.0, n = 0.0, mn = 1e300, mx = -1e300;
        for (int w = 0; w < world; ++w) {
            const double a0 = __shfl_sync(0xffffffffu, v[0], w), a1 = __shfl_sync(0xffffffffu, v[1], w);
            const double a2 = __shfl_sync(0xffffffffu, v[2], w), a3 = __shfl_sync(0xffffffffu, v[3], w);
            const double a4 = __shfl_sync(0xffffffffu, v[4], w);
            if (a2 > 0.0) { s += a0; ss += a1; n += a2; mn = fmin(mn, a3); mx = fmax(mx, a4); }
        }
        double inv = 1.0;
        if (n > 0.0 && mn != mx) {
            const double mean = s / n;
            const double var = fmax(ss / n - mean * mean, 0.0);
            if (var > 0.0) inv = 1.0 / sqrt(var);
        }
        if (threadIdx.x == 0) inv_sd_s = inv;
    }
    __syncthreads();
    const float inv_sd = (float)inv_sd_s;
    // ---- 2. the owner reduces its slice in rank order and hands it to every peer ----
    const int64_t lo = (int64_t)rank * per, hi = min(lo + per, n4);
    for (int64_t i = lo + g0; i < hi; i += gstride) {
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int w = 0; w < world; ++w) {
            const float4 t = __ldcg(reinterpret_cast<const float4*>(mine + slots0 + (size_t)w * slot + 16 * (size_t)i));
            g.x += t.x; g.y += t.y; g.z += t.z; g.w += t.w;
        }
        g.x *= inv_sd; g.y *= inv_sd; g.z *= inv_sd; g.w *= inv_sd;
        const float t[4] = {g.x, g.y, g.z, g.w};
        for (int k = 0; k < 4 && 4 * i + k < P; ++k) grad_out[4 * i + k] = t[k];
        for (int w = 1; w < world; ++w) *reinterpret_cast<float4*>(mailboxes[(rank + w) % world] + res_off + 16 * (size_t)i) = g;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        ticket_s = atomicAdd(bar_ctr2, 1u);
    }
    __syncthreads();
    if (ticket_s == gridDim.x - 1) {
        if (threadIdx.x == 0) __threadfence_system();
        __syncthreads();
        if (threadIdx.x < world)
            st_release_sys(reinterpret_cast<unsigned long long*>(mailboxes[threadIdx.x] + off2) + par * XCHG_MAX_WORLD + rank, step + 1ull);
        if (threadIdx.x == 0) {
            *bar_ctr2 = 0u;
            *step_ctr = step + 1ull;
        }
    }
    // ---- 3. all-gather: the other owners' slices out of this rank's result region ----
    if (threadIdx.x < world) {
        const unsigned long long* f = reinterpret_cast<const unsigned long long*>(mine + off2) + par * XCHG_MAX_WORLD + threadIdx.x;
        unsigned spins = 0;
        while (ld_acquire_sys(f) < step + 1ull)
            if (++spins > (1u << 26)) __trap();
    }
    __syncthreads();
    for (int64_t i = g0; i < n4; i += gstride) {
        if (i >= lo && i < hi) continue;
        const float4 g = __ldcg(reinterpret_cast<const float4*>(mine + res_off + 16 * (size_t)i));
        const float t[4] = {g.x, g.y, g.z, g.w};
        for (int k = 0; k < 4 && 4 * i + k < P; ++k) grad_out[4 * i + k] = t[k];
    }
}

// All-gather of a short fp64 vector (the rewards of every rank's returns: the standardisation of fd_state batches needs
// the mean / std over ALL ranks' returns before the coefficients can be formed, learner/finite_differences.py:40-43) with
// the same mailbox protocol and step counter: push n doubles into slot [parity][rank] of every peer, publish, wait for all
// ranks, copy the world slots in rank order to dst[world][n].  One CTA; n * 8 bytes must fit a slot.
__global__ void __launch_bounds__(XCHG_THREADS) xchg_gather_kernel(char* const* __restrict__ mailboxes, int rank, int world,
                                                                   int64_t P, const double* __restrict__ src, int n,
                                                                   double* __restrict__ dst) {
    char* const mine = mailboxes[rank];
    unsigned long long* step_ctr = reinterpret_cast<unsigned long long*>(mine + 8);
    const unsigned long long step = *reinterpret_cast<volatile unsigned long long*>(step_ctr);
    const int par = (int)(step & 1ull);
    const size_t slot = xchg_slot_bytes(P);
    const size_t slots0 = XCHG_HDR + XCHG_FLAGS + (size_t)par * world * slot;
    const size_t my_slot_off = slots0 + (size_t)rank * slot;
    for (int i = threadIdx.x; i < n; i += XCHG_THREADS) {
        const double v = src[i];
        for (int w = 0; w < world; ++w) {
            const int d = (rank + w) % world;
            *reinterpret_cast<double*>(mailboxes[d] + my_slot_off + 8 * (size_t)i) = v;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) __threadfence_system();
    __syncthreads();
    if (threadIdx.x < world) {
        unsigned long long* f = reinterpret_cast<unsigned long long*>(mailboxes[threadIdx.x] + XCHG_HDR) + par * XCHG_MAX_WORLD + rank;
        st_release_sys(f, step + 1ull);
    }
    if (threadIdx.x == 0) *step_ctr = step + 1ull;
    if (threadIdx.x < world) {
        const unsigned long long* f = reinterpret_cast<const unsigned long long*>(mine + XCHG_HDR) + par * XCHG_MAX_WORLD + threadIdx.x;
        unsigned spins = 0;
        while (ld_acquire_sys(f) < step + 1ull) {
            if (++spins > (1u << 26)) __trap();    // a missing peer must fault, not hang the GPU
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < world * n; i += XCHG_THREADS) {
        const int w = i / n, j = i - w * n;
        dst[i] = __ldcg(reinterpret_cast<const double*>(mine + slots0 + (size_t)w * slot) + j);
    }
}

}  // namespace

static unsigned long long* g_xchg_prof = nullptr;

extern "C" size_t dfd_xchg_mailbox_bytes(int64_t n_params, int world) {
    if (n_params <= 0 || world <= 0 || world > XCHG_MAX_WORLD) return 0;
    // flag-protocol slots, then (short parameter vectors only) the low-latency packet region of the one-kernel step
    // (csrc/fd_tail.cu: tl_ll_slot_bytes - 8 bytes per value + 16 packets of statistics)
    const size_t ll_slot = n_params <= 32768 ? (size_t)((n_params + 3) / 4 * 4 + 16) * 8 : 0;
    // long vectors: second flag array + result region (2 parities) of the two-phase kernel
    const size_t rs = n_params > 32768 ? XCHG_FLAGS + 2 * xchg_slot_bytes(n_params) : 0;
    return XCHG_HDR + XCHG_FLAGS + 2 * (size_t)world * xchg_slot_bytes(n_params) + 2 * (size_t)world * ll_slot + rs;
}

extern "C" int dfd_xchg_mailbox_create(dfd_ctx* ctx, size_t bytes, void** mailbox, unsigned char* ipc_handle64) {
    DFD_CHECK_ARG(ctx && mailbox && ipc_handle64 && bytes > 0, "dfd_xchg_mailbox_create: bad argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    void* p = nullptr;
    DFD_CUDA(cudaMalloc(&p, bytes));   // its own allocation: IPC handles cover whole cudaMalloc blocks
    DFD_CUDA(cudaMemset(p, 0, bytes));
    cudaIpcMemHandle_t h;
    DFD_CUDA(cudaIpcGetMemHandle(&h, p));
    memcpy(ipc_handle64, &h, 64);
    *mailbox = p;
    return 0;
}

extern "C" int dfd_xchg_mailbox_open(dfd_ctx* ctx, const unsigned char* ipc_handle64, void** peer_mailbox) {
    DFD_CHECK_ARG(ctx && ipc_handle64 && peer_mailbox, "dfd_xchg_mailbox_open: bad argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, ipc_handle64, 64);
    DFD_CUDA(cudaIpcOpenMemHandle(peer_mailbox, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}

extern "C" int dfd_xchg_mailbox_close(dfd_ctx* ctx, void* peer_mailbox) {
    DFD_CHECK_ARG(ctx && peer_mailbox, "dfd_xchg_mailbox_close: bad argument");
    DFD_CUDA(cudaIpcCloseMemHandle(peer_mailbox));
    return 0;
}

extern "C" int dfd_xchg_mailbox_destroy(dfd_ctx* ctx, void* mailbox) {
    DFD_CHECK_ARG(ctx && mailbox, "dfd_xchg_mailbox_destroy: bad argument");
    DFD_CUDA(cudaFree(mailbox));
    return 0;
}

extern "C" int dfd_xchg_allreduce(dfd_ctx* ctx, void* const* mailboxes, int rank, int world, int64_t n_params,
                                  const float* grad_partial, const double* stats5, float* grad_out, dfd_stream stream) {
    DFD_CHECK_ARG(ctx && mailboxes && grad_partial && stats5 && grad_out, "dfd_xchg_allreduce: NULL argument");
    DFD_CHECK_ARG(world >= 1 && world <= XCHG_MAX_WORLD && rank >= 0 && rank < world, "dfd_xchg_allreduce: rank %d / world %d", rank, world);
    DFD_CHECK_ARG(n_params > 0 && (((uintptr_t)grad_partial) & 15) == 0, "dfd_xchg_allreduce: grad_partial must be 16-byte aligned");
    const int64_t n4 = (n_params + 3) / 4;
    int grid = (int)((n4 + XCHG_THREADS - 1) / XCHG_THREADS);
    if (grid > ctx->sm_count) grid = ctx->sm_count;   // all CTAs co-resident: the in-kernel arrival count cannot deadlock
    static const bool want_prof = getenv("DFD_XCHG_PROF") != nullptr;
    if (want_prof && !g_xchg_prof) {
        cudaMalloc(&g_xchg_prof, 64);
        cudaMemset(g_xchg_prof, 0, 64);
    }
    unsigned long long* prof = g_xchg_prof;
    static const bool one_phase = getenv("DFD_XCHG_ONEPHASE") != nullptr;
    if (world >= 4 && n_params >= 65536 && !one_phase) {      // long vectors on many ranks: reduce-scatter + all-gather
        const size_t off2 = XCHG_HDR + XCHG_FLAGS + 2 * (size_t)world * xchg_slot_bytes(n_params);
        xchg_allreduce_rs_kernel<<<grid, XCHG_THREADS, 0, (cudaStream_t)stream>>>((char* const*)mailboxes, rank, world, n_params,
                                                                                  grad_partial, stats5, grad_out, off2);
        DFD_LAUNCHED(ctx);
        return 0;
    }
    xchg_allreduce_kernel<<<grid, XCHG_THREADS, 0, (cudaStream_t)stream>>>((char* const*)mailboxes, rank, world, n_params,
                                                                           grad_partial, stats5, grad_out, prof);
    DFD_LAUNCHED(ctx);
    return 0;
}

extern "C" int dfd_xchg_gather_f64(dfd_ctx* ctx, void* const* mailboxes, int rank, int world, int64_t n_params, const double* src,
                                   int n, double* dst, dfd_stream stream) {
    DFD_CHECK_ARG(ctx && mailboxes && src && dst, "dfd_xchg_gather_f64: NULL argument");
    DFD_CHECK_ARG(world >= 1 && world <= XCHG_MAX_WORLD && rank >= 0 && rank < world, "dfd_xchg_gather_f64: rank %d / world %d", rank, world);
    DFD_CHECK_ARG(n >= 0 && (size_t)n * 8 + 256 <= xchg_slot_bytes(n_params),
                  "dfd_xchg_gather_f64: %d doubles do not fit a mailbox slot of a %lld-parameter exchange", n, (long long)n_params);
    if (n == 0) return 0;
    xchg_gather_kernel<<<1, XCHG_THREADS, 0, (cudaStream_t)stream>>>((char* const*)mailboxes, rank, world, n_params, src, n, dst);
    DFD_LAUNCHED(ctx);
    return 0;
}

// debugging aid: mean nanoseconds per phase (push, publish, wait for peers, combine) accumulated so far
extern "C" int dfd_xchg_profile(double* out5) {
    if (!g_xchg_prof || !out5) return 1;
    unsigned long long h[5];
    cudaDeviceSynchronize();
    cudaMemcpy(h, g_xchg_prof, sizeof(h), cudaMemcpyDeviceToHost);
    for (int i = 0; i < 4; ++i) out5[i] = h[4] ? (double)h[i] / (double)h[4] : 0.0;
    out5[4] = (double)h[4];
    return 0;
}
