// rng_normal.cu - the reference's DEFAULT noise source drawn on the device (first half).
//
// utils/noise_sources.py:4-20 (RNGNoiseSource): sample() returns the PCG64 "state,inc" key and
// rng.standard_normal(n_params); decode(key) rewinds the generator to the key and redraws.  The reference worker
// (worker/worker.py:26-30) and learner (learner/finite_differences.py:87) call these once per member / per return on
// the host: 10 ns per normal, i.e. 3.5 s for the 2048 x 171 042 normals of one Humanoid-sized batch.  Here the rows of
// a whole batch are drawn by five launches, bit-identical to numpy (csrc/rng_normal_core.h has the algorithm and how
// parity is pinned), straight into the row table the forward / reduction kernels read:
//   rng_table_kernel    one thread per 32-word chunk: LCG jump-ahead to the chunk, path 0 of the chunk
//   rng_resolve_kernel  one thread per chunk: entry offset (speculated from the previous chunk's path-0 exit and
//                       verified), number of normals the chunk returns
//   rng_serial_kernel   one thread per stream, only for streams with a chunk that did not verify (never seen with
//                       32-word chunks; exercised by the tests through `force_serial`)
//   rng_blocksum_kernel / rng_scan_kernel   exclusive prefix of the per-chunk counts, per stream
//   rng_emit_kernel     one thread per chunk replays it from its entry; a CTA's normals are contiguous in the
//                       stream, so they are staged in shared memory and leave as coalesced rows:
//                       fp32(eps) for the learner, fp32(fp64(theta) + sigma * eps) for the worker
//                       (worker/worker.py:28 + policies/policy.py:40-42), plus the word count at every row end
//                       (= the key of the next row).
// All integer work and IEEE double arithmetic; HBM traffic is the rows written once (4 bytes per normal) plus 18 bytes
// of chunk records per 32 words.
#include <type_traits>
#include "common.cuh"
#include "rng_normal_core.h"

namespace {

constexpr int RNG_BLOCK = 128;          // chunks (threads) per CTA of the table / resolve / emit kernels
constexpr int RNG_JUMP_BITS = 48;       // chunk index bits served by the jump table (2^48 chunks = 2^53 words)

__device__ const uint64_t g_zig_ki[256] = ZIG_KI_INIT;
__device__ const uint64_t g_zig_wi[256] = ZIG_WI_BITS_INIT;
__device__ const uint64_t g_zig_fi[256] = ZIG_FI_BITS_INIT;
__device__ const uint64_t g_exp_tab[256] = EXP_TAB_INIT;   // wedge tests only (1.2 % of the attempts): read through L1

// state after n steps = A^n * s + G_n * inc (mod 2^128) with G_n = 1 + A + ... + A^(n-1): both universal, so the
// jump to chunk c is one (multiply, multiply, add) per set bit of c.  Entries j: n = RNGN_CHUNK * 2^j.
struct JumpTable {
    rngn_u128 mult[RNG_JUMP_BITS];
    rngn_u128 geo[RNG_JUMP_BITS];
};
__constant__ JumpTable c_jump;

struct RngArgs {
    const uint64_t* streams;   // n_streams x {state_lo, state_hi, inc_lo, inc_hi}
    int n_streams;
    int64_t n_chunks;          // per stream
    int64_t n_blocks;          // per stream: ceil(n_chunks / RNG_BLOCK)
    int64_t n_draws;           // per stream
    int64_t n_params;
    int64_t rows_per_stream;
    rngn_rec* rec;             // n_streams x n_chunks
    uint8_t* entry;            // n_streams x n_chunks (0..254; 255 = "at least a whole chunk", kept in entry_big)
    uint8_t* nout;             // n_streams x n_chunks
    int32_t* entry_big;        // n_streams x n_chunks, written only where entry == 255
    int* fail;                 // n_streams
    int32_t* blocksum;         // n_streams x n_blocks
    int64_t* blockoff;         // n_streams x n_blocks
    unsigned* status;
    int libm_fused;
    int force_serial;
};

__device__ __forceinline__ void load_tables(uint64_t* sm, rngn_tables& t, int fused) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        sm[i] = g_zig_ki[i];
        sm[256 + i] = g_zig_wi[i];
        sm[512 + i] = g_zig_fi[i];
    }
    __syncthreads();
    t.ki = sm;
    t.wi = reinterpret_cast<const double*>(sm + 256);
    t.fi = reinterpret_cast<const double*>(sm + 512);
    t.exp_tab = g_exp_tab;
    t.libm_fused = fused;
}

__device__ __forceinline__ void stream_state(const uint64_t* streams, int s, rngn_u128& st, rngn_u128& inc) {
    st.lo = streams[4 * s + 0];
    st.hi = streams[4 * s + 1];
    inc.lo = streams[4 * s + 2];
    inc.hi = streams[4 * s + 3];
}

__device__ __forceinline__ rngn_u128 jump_to_chunk(rngn_u128 s, rngn_u128 inc, int64_t c) {
    uint64_t bits = (uint64_t)c;
    for (int j = 0; bits != 0; ++j, bits >>= 1)
        if (bits & 1) s = rngn_add(rngn_mul(c_jump.mult[j], s), rngn_mul(c_jump.geo[j], inc));
    return s;
}

// state before the first word of this thread's chunk (chunk blockIdx.x * RNG_BLOCK + threadIdx.x of the stream): one
// thread jumps to the CTA's first chunk (one multiply-multiply-add per set bit of the block number: up to ~20 of them,
// a fifth of a thread's work if every thread did it), the others add their own seven bits.  Contains a __syncthreads.
__device__ __forceinline__ rngn_u128 chunk_state(rngn_u128 st, rngn_u128 inc) {
    __shared__ rngn_u128 base_s;
    if (threadIdx.x == 0) base_s = jump_to_chunk(st, inc, (int64_t)blockIdx.x * RNG_BLOCK);
    __syncthreads();
    return jump_to_chunk(base_s, inc, (int64_t)threadIdx.x);
}

__global__ void __launch_bounds__(RNG_BLOCK) rng_table_kernel(RngArgs a) {
    __shared__ uint64_t sm[768];
    rngn_tables t;
    load_tables(sm, t, a.libm_fused);
    const int s = blockIdx.y;
    const int64_t c = (int64_t)blockIdx.x * RNG_BLOCK + threadIdx.x;
    rngn_u128 st, inc;
    stream_state(a.streams, s, st, inc);
    const rngn_u128 sc = chunk_state(st, inc);
    if (c >= a.n_chunks) return;
    unsigned status = 0;
    a.rec[(int64_t)s * a.n_chunks + c] = rngn_table_chunk(sc, inc, t, &status);
    if (status) atomicOr(a.status, status);
}

__device__ __forceinline__ void store_entry(const RngArgs& a, int64_t at, int e, int k) {
    a.entry[at] = (uint8_t)(e < 255 ? e : 255);
    if (e >= 255) a.entry_big[at] = e;
    a.nout[at] = (uint8_t)k;
}

__global__ void __launch_bounds__(RNG_BLOCK) rng_resolve_kernel(RngArgs a) {
    __shared__ uint64_t sm[768];
    rngn_tables t;
    load_tables(sm, t, a.libm_fused);
    const int s = blockIdx.y;
    const int64_t c = (int64_t)blockIdx.x * RNG_BLOCK + threadIdx.x;
    if (c >= a.n_chunks) return;
    rngn_u128 st, inc;
    stream_state(a.streams, s, st, inc);
    unsigned status = 0;
    int e, k, fail = 0;
    rngn_resolve_chunk(a.rec + (int64_t)s * a.n_chunks, c, st, inc, t, &e, &k, &status, &fail);
    store_entry(a, (int64_t)s * a.n_chunks + c, e, k);
    if (fail) a.fail[s] = 1;
    if (status) atomicOr(a.status, status);
}

// one thread per stream; does nothing unless the stream failed to verify (or the caller forces it)
__global__ void __launch_bounds__(32) rng_serial_kernel(RngArgs a) {
    __shared__ uint64_t sm[768];
    rngn_tables t;
    load_tables(sm, t, a.libm_fused);
    const int s = blockIdx.x * 32 + threadIdx.x;
    if (s >= a.n_streams || !(a.fail[s] || a.force_serial)) return;
    rngn_u128 st, inc;
    stream_state(a.streams, s, st, inc);
    unsigned status = RNGN_ST_SERIAL;
    const rngn_rec* rec = a.rec + (int64_t)s * a.n_chunks;
    int64_t e = 0;
    for (int64_t c = 0; c < a.n_chunks; ++c) {          // rngn_resolve_serial with this file's entry encoding
        rngn_rec r = rec[c];
        int k;
        int64_t e_in = e;
        if (e >= RNGN_CHUNK) {
            k = 0;
            e -= RNGN_CHUNK;
        } else if ((r.start_mask >> e) & 1u) {
            k = rngn_popc(r.out_mask >> e);
            e = r.exit0;
        } else {
            uint32_t smk, omk;
            e = rngn_chunk_path(rngn_advance(st, inc, (uint64_t)c * RNGN_CHUNK), inc, t, (int)e, &smk, &omk, &status);
            k = rngn_popc(omk);
        }
        store_entry(a, (int64_t)s * a.n_chunks + c, (int)e_in, k);
    }
    atomicOr(a.status, status);
}

__device__ __forceinline__ int block_exclusive_scan(int v, int* total) {
    __shared__ int warp_sum[RNG_BLOCK / 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, x, d);
        if (lane >= d) x += y;
    }
    if (lane == 31) warp_sum[w] = x;
    __syncthreads();
    int base = 0, tot = 0;
#pragma unroll
    for (int i = 0; i < RNG_BLOCK / 32; ++i) {
        if (i < w) base += warp_sum[i];
        tot += warp_sum[i];
    }
    *total = tot;
    return base + x - v;
}

__global__ void __launch_bounds__(RNG_BLOCK) rng_blocksum_kernel(RngArgs a) {
    const int s = blockIdx.y;
    const int64_t c = (int64_t)blockIdx.x * RNG_BLOCK + threadIdx.x;
    int v = c < a.n_chunks ? a.nout[(int64_t)s * a.n_chunks + c] : 0;
    int tot;
    block_exclusive_scan(v, &tot);
    if (threadIdx.x == 0) a.blocksum[(int64_t)s * a.n_blocks + blockIdx.x] = tot;
}

// one CTA per stream: exclusive prefix of the block sums (sequential over tiles of 1024 with a running carry)
__global__ void __launch_bounds__(1024) rng_scan_kernel(RngArgs a) {
    __shared__ long long wsum[32];
    __shared__ long long carry_s;
    const int s = blockIdx.x;
    const int32_t* in = a.blocksum + (int64_t)s * a.n_blocks;
    int64_t* out = a.blockoff + (int64_t)s * a.n_blocks;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int64_t base = 0; base < a.n_blocks; base += 1024) {
        const int64_t i = base + threadIdx.x;
        long long v = i < a.n_blocks ? in[i] : 0, x = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            long long y = __shfl_up_sync(0xffffffffu, x, d);
            if (lane >= d) x += y;
        }
        if (lane == 31) wsum[w] = x;
        __syncthreads();
        long long pre = carry_s;
        for (int k = 0; k < w; ++k) pre += wsum[k];
        if (i < a.n_blocks) out[i] = pre + x - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = pre + x;
        __syncthreads();
    }
    if (threadIdx.x == 0 && carry_s < a.n_draws) atomicOr(a.status, RNGN_ST_SHORT);
}

struct EmitArgs {
    const float* theta;        // nullable
    double sigma;
    const int32_t* dest_row;   // nullable: generated row r of the launch -> row of rows_out
    float* rows_out;           // nullable
    double* rows_out_f64;      // nullable
    int64_t row_stride;
    int64_t* row_words;        // n_streams x (rows_per_stream + 1)
    uint64_t* row_state;       // nullable: n_streams x (rows_per_stream + 1) x {lo, hi}
};

// F64 = the caller also wants the raw fp64 normals: staged as doubles, transformed when they leave.  Otherwise the sink
// finishes a normal on the spot - fp32(eps), or fp32(fp64(theta[col]) + sigma * eps) with theta read through L1 (a lane
// walks consecutive columns) - and stages 4 bytes: 22 KB of shared memory per CTA instead of 39, twice the resident warps
// for a loop that is one long dependent chain per thread.
template <bool F64>
struct StageSink {
    typename std::conditional<F64, double, float>::type* stage;
    int64_t g0;                // first normal of this CTA
    int64_t next_row_end;      // index of the next normal that ends a row
    int64_t n_params;
    int64_t col;               // column of the next normal
    const float* theta;
    double sigma;
    int64_t* row_words;
    uint64_t* row_state;
    __device__ __forceinline__ void operator()(int64_t g, double v, int64_t words_after, rngn_u128 st) {
        if (F64) stage[g - g0] = v;
        else stage[g - g0] = __double2float_rn(theta ? __dadd_rn((double)__ldg(theta + col), __dmul_rn(sigma, v)) : v);
        ++col;
        if (g == next_row_end) {
            const int64_t r = (g + 1) / n_params;
            row_words[r] = words_after;
            if (row_state) {
                row_state[2 * r] = st.lo;
                row_state[2 * r + 1] = st.hi;
            }
            next_row_end += n_params;
            col = 0;
        }
    }
};

template <bool F64>
__global__ void __launch_bounds__(RNG_BLOCK) rng_emit_kernel(RngArgs a, EmitArgs o) {
    typedef typename std::conditional<F64, double, float>::type stage_t;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* sm = reinterpret_cast<uint64_t*>(smem_raw);
    stage_t* stage = reinterpret_cast<stage_t*>(sm + 768);           // RNG_BLOCK * RNGN_CHUNK values
    rngn_tables t;
    load_tables(sm, t, a.libm_fused);
    const int s = blockIdx.y;
    const int64_t c = (int64_t)blockIdx.x * RNG_BLOCK + threadIdx.x;
    const int64_t at = (int64_t)s * a.n_chunks + c;
    const int64_t g0 = a.blockoff[(int64_t)s * a.n_blocks + blockIdx.x];
    int k = c < a.n_chunks ? a.nout[at] : 0;
    int tot;
    const int pre = block_exclusive_scan(k, &tot);
    if (g0 >= a.n_draws) return;                                     // uniform per CTA
    rngn_u128 st, inc;
    stream_state(a.streams, s, st, inc);
    const rngn_u128 sc = chunk_state(st, inc);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        o.row_words[(int64_t)s * (a.rows_per_stream + 1)] = 0;
        if (o.row_state) {
            o.row_state[2 * (int64_t)s * (a.rows_per_stream + 1)] = a.streams[4 * s];
            o.row_state[2 * (int64_t)s * (a.rows_per_stream + 1) + 1] = a.streams[4 * s + 1];
        }
    }
    if (c < a.n_chunks && k > 0) {
        int e = a.entry[at];
        if (e == 255) e = a.entry_big[at];
        const int64_t first_g = g0 + pre;
        StageSink<F64> sink;
        sink.stage = stage;
        sink.g0 = g0;
        sink.n_params = a.n_params;
        const int64_t row0 = first_g / a.n_params;
        sink.next_row_end = (row0 + 1) * a.n_params - 1;
        sink.col = first_g - row0 * a.n_params;
        sink.theta = o.theta;
        sink.sigma = o.sigma;
        sink.row_words = o.row_words + (int64_t)s * (a.rows_per_stream + 1);
        sink.row_state = o.row_state ? o.row_state + 2 * (int64_t)s * (a.rows_per_stream + 1) : nullptr;
        unsigned status = 0;
        rngn_emit_chunk(sc, inc, t, c, e, first_g, a.n_draws, &status, sink);
        // status bits were already reported by the table / resolve kernels for the same attempts
    }
    __syncthreads();
    // the CTA's normals [g0, g0 + tot) leave as rows
    int64_t n_here = a.n_draws - g0 < tot ? a.n_draws - g0 : tot;
    if (threadIdx.x >= n_here) return;
    int64_t g = g0 + threadIdx.x;
    int64_t row = g / a.n_params, col = g - row * a.n_params;
    for (int64_t i = threadIdx.x; i < n_here; i += RNG_BLOCK) {
        const int64_t r_launch = (int64_t)s * a.rows_per_stream + row;
        const int64_t r_out = o.dest_row ? o.dest_row[r_launch] : r_launch;
        if (F64) {
            const double eps = (double)stage[i];
            if (o.rows_out_f64) o.rows_out_f64[r_out * o.row_stride + col] = eps;
            if (o.rows_out) {
                // worker.py:28 with fp64 noise: flat (fp32 -> fp64) + sigma * eps, two roundings, then the fp32 cast of
                // set_trainable_flat; the learner's decode is the plain cast
                const double v = o.theta ? __dadd_rn((double)o.theta[col], __dmul_rn(o.sigma, eps)) : eps;
                o.rows_out[r_out * o.row_stride + col] = __double2float_rn(v);
            }
        } else {
            o.rows_out[r_out * o.row_stride + col] = (float)stage[i];
        }
        col += RNG_BLOCK;
        while (col >= a.n_params) {
            col -= a.n_params;
            ++row;
        }
    }
}

struct Layout {
    int64_t n_chunks, n_blocks;
    size_t off_rec, off_entry, off_nout, off_big, off_fail, off_bsum, off_boff, total;
};

Layout plan(int n_streams, int64_t n_draws, double margin) {
    Layout L;
    // expected words per normal 1.022 (98.8 % one word, wedges two, tails three and more)
    int64_t words = (int64_t)((double)n_draws * margin) + 1024;
    L.n_chunks = (words + RNGN_CHUNK - 1) / RNGN_CHUNK;
    L.n_blocks = (L.n_chunks + RNG_BLOCK - 1) / RNG_BLOCK;
    const size_t nc = (size_t)n_streams * (size_t)L.n_chunks, nb = (size_t)n_streams * (size_t)L.n_blocks;
    size_t o = 0;
    L.off_rec = o;   o = dfd_align_up(o + nc * sizeof(rngn_rec), 256);
    L.off_entry = o; o = dfd_align_up(o + nc, 256);
    L.off_nout = o;  o = dfd_align_up(o + nc, 256);
    L.off_big = o;   o = dfd_align_up(o + nc * sizeof(int32_t), 256);
    L.off_fail = o;  o = dfd_align_up(o + (size_t)n_streams * sizeof(int), 256);
    L.off_bsum = o;  o = dfd_align_up(o + nb * sizeof(int32_t), 256);
    L.off_boff = o;  o = dfd_align_up(o + nb * sizeof(int64_t), 256);
    L.total = o;
    return L;
}

int upload_jump_table() {
    static bool done[64] = {};
    int dev = 0;
    DFD_CUDA(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && done[dev]) return 0;
    JumpTable jt;
    // (mult, geo) for n steps compose as: n1 then n2 -> mult = m2*m1, geo = m2*g1 + g2
    rngn_u128 m = {RNGN_MULT_LO, RNGN_MULT_HI}, g = {1, 0};
    for (int n = 1; n < RNGN_CHUNK; n <<= 1) {              // RNGN_CHUNK is a power of two: double up to it
        g = rngn_add(rngn_mul(m, g), g);
        m = rngn_mul(m, m);
    }
    for (int j = 0; j < RNG_JUMP_BITS; ++j) {
        jt.mult[j] = m;
        jt.geo[j] = g;
        g = rngn_add(rngn_mul(m, g), g);
        m = rngn_mul(m, m);
    }
    DFD_CUDA(cudaMemcpyToSymbol(c_jump, &jt, sizeof(jt)));
    if (dev >= 0 && dev < 64) done[dev] = true;
    return 0;
}

}  // namespace

static_assert((RNGN_CHUNK & (RNGN_CHUNK - 1)) == 0 && RNGN_CHUNK <= 32, "chunk must be a power of two <= 32");

extern "C" size_t dfd_rng_scratch_bytes(int n_streams, int64_t rows_per_stream, int64_t n_params, double margin) {
    if (n_streams <= 0 || rows_per_stream <= 0 || n_params <= 0) return 0;
    return plan(n_streams, rows_per_stream * n_params, margin < 1.0 ? 1.04 : margin).total + 256;
}

extern "C" int dfd_rng_normal_rows(dfd_ctx* ctx, const uint64_t* streams, int n_streams, int64_t rows_per_stream,
                                   int64_t n_params, const float* theta, double sigma, const int32_t* dest_row,
                                   float* rows_out, double* rows_out_f64, int64_t row_stride, int64_t* row_words,
                                   uint64_t* row_state, uint32_t* status, int libm_fused, int force_serial, double margin, void* scratch,
                                   size_t scratch_bytes, dfd_stream stream) {
    DFD_CHECK_ARG(ctx && streams && row_words && status && scratch, "dfd_rng_normal_rows: null argument");
    DFD_CHECK_ARG(n_streams > 0 && n_streams <= 65535, "dfd_rng_normal_rows: n_streams %d not in 1..65535", n_streams);
    DFD_CHECK_ARG(rows_per_stream > 0 && n_params > 0 && row_stride >= n_params, "dfd_rng_normal_rows: bad row geometry");
    DFD_CHECK_ARG(rows_out || rows_out_f64, "dfd_rng_normal_rows: no output");
    DFD_CHECK_ARG((reinterpret_cast<uintptr_t>(scratch) & 255) == 0, "dfd_rng_normal_rows: scratch not 256-byte aligned");
    if (margin < 1.0) margin = 1.04;
    const int64_t n_draws = rows_per_stream * n_params;
    Layout L = plan(n_streams, n_draws, margin);
    DFD_CHECK_ARG(scratch_bytes >= L.total, "dfd_rng_normal_rows: scratch %zu < %zu bytes", scratch_bytes, L.total);
    DFD_CHECK_ARG(L.n_blocks < (int64_t)1 << 31, "dfd_rng_normal_rows: stream too long");
    if (int rc = upload_jump_table()) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    unsigned char* base = static_cast<unsigned char*>(scratch);
    RngArgs a;
    a.streams = streams;
    a.n_streams = n_streams;
    a.n_chunks = L.n_chunks;
    a.n_blocks = L.n_blocks;
    a.n_draws = n_draws;
    a.n_params = n_params;
    a.rows_per_stream = rows_per_stream;
    a.rec = reinterpret_cast<rngn_rec*>(base + L.off_rec);
    a.entry = base + L.off_entry;
    a.nout = base + L.off_nout;
    a.entry_big = reinterpret_cast<int32_t*>(base + L.off_big);
    a.fail = reinterpret_cast<int*>(base + L.off_fail);
    a.blocksum = reinterpret_cast<int32_t*>(base + L.off_bsum);
    a.blockoff = reinterpret_cast<int64_t*>(base + L.off_boff);
    a.status = status;
    a.libm_fused = libm_fused;
    a.force_serial = force_serial;
    EmitArgs o;
    o.theta = theta;
    o.sigma = sigma;
    o.dest_row = dest_row;
    o.rows_out = rows_out;
    o.rows_out_f64 = rows_out_f64;
    o.row_stride = row_stride;
    o.row_words = row_words;
    o.row_state = row_state;
    DFD_CUDA(cudaMemsetAsync(status, 0, sizeof(uint32_t), st));
    DFD_CUDA(cudaMemsetAsync(a.fail, 0, (size_t)n_streams * sizeof(int), st));
    const dim3 grid((unsigned)L.n_blocks, (unsigned)n_streams);
    rng_table_kernel<<<grid, RNG_BLOCK, 0, st>>>(a);
    DFD_LAUNCHED(ctx);
    rng_resolve_kernel<<<grid, RNG_BLOCK, 0, st>>>(a);
    DFD_LAUNCHED(ctx);
    rng_serial_kernel<<<(n_streams + 31) / 32, 32, 0, st>>>(a);
    DFD_LAUNCHED(ctx);
    rng_blocksum_kernel<<<grid, RNG_BLOCK, 0, st>>>(a);
    DFD_LAUNCHED(ctx);
    rng_scan_kernel<<<n_streams, 1024, 0, st>>>(a);
    DFD_LAUNCHED(ctx);
    if (rows_out_f64) {
        const size_t smem = 768 * sizeof(uint64_t) + (size_t)RNG_BLOCK * RNGN_CHUNK * sizeof(double);
        static bool attr_set = false;
        if (!attr_set) {
            DFD_CUDA(cudaFuncSetAttribute(rng_emit_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr_set = true;
        }
        rng_emit_kernel<true><<<grid, RNG_BLOCK, smem, st>>>(a, o);
    } else {
        const size_t smem = 768 * sizeof(uint64_t) + (size_t)RNG_BLOCK * RNGN_CHUNK * sizeof(float);
        rng_emit_kernel<false><<<grid, RNG_BLOCK, smem, st>>>(a, o);
    }
    DFD_LAUNCHED(ctx);
    return 0;
}
