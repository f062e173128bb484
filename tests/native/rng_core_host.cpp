// Host build of csrc/rng_normal_core.h for tests/test_rng_core_cpu.py: the SAME per-chunk functions the kernels of
// csrc/rng_normal.cu run, driven by plain loops, so the parallel form is pinned against numpy on the CPU.
// Test infrastructure only (g++ -O2 -ffp-contract=off -shared -fPIC); nothing in the product loads it.
#include <vector>
#include <cstring>
#include "rng_normal_core.h"

static const uint64_t KI[256] = ZIG_KI_INIT;
static const uint64_t WI_BITS[256] = ZIG_WI_BITS_INIT;
static const uint64_t FI_BITS[256] = ZIG_FI_BITS_INIT;
static const uint64_t EXP_TAB[256] = EXP_TAB_INIT;

static int g_fused = 1;

static rngn_tables tables() {
    rngn_tables t;
    t.libm_fused = g_fused;
    t.exp_tab = EXP_TAB;
    t.ki = KI;
    t.wi = reinterpret_cast<const double*>(WI_BITS);
    t.fi = reinterpret_cast<const double*>(FI_BITS);
    return t;
}

extern "C" {

// the stream as numpy walks it: one attempt after the other
int rngn_host_sequential(uint64_t s_lo, uint64_t s_hi, uint64_t i_lo, uint64_t i_hi, int64_t n, double* out, int64_t* words) {
    rngn_u128 s = {s_lo, s_hi}, inc = {i_lo, i_hi};
    unsigned status = 0;
    int64_t w = 0, g = 0;
    rngn_tables t = tables();
    while (g < n) {
        rngn_attempt a = rngn_attempt_at(&s, inc, t, &status);
        w += a.len;
        if (a.out) out[g++] = a.val;
    }
    *words = w;
    return (int)status;
}

struct Sink {
    double* out;
    int64_t* row_words;
    int64_t P;
    uint64_t* row_state;
    void operator()(int64_t g, double v, int64_t words_after, rngn_u128 st) {
        out[g] = v;
        if ((g + 1) % P == 0) {
            row_words[(g + 1) / P] = words_after;
            row_state[2 * ((g + 1) / P)] = st.lo;
            row_state[2 * ((g + 1) / P) + 1] = st.hi;
        }
    }
};

// the parallel form, chunk by chunk: tables -> speculative resolve (+ serial fallback) -> scan -> emit
// force_serial = 1 runs the serial resolver regardless (to test it).  Returns status bits; *failed = a chunk did not verify.
int rngn_host_chunked(uint64_t s_lo, uint64_t s_hi, uint64_t i_lo, uint64_t i_hi, int64_t n, int64_t P, int64_t n_chunks,
                      int force_serial, double* out, int64_t* row_words, uint64_t* row_state, int* failed) {
    rngn_u128 s0 = {s_lo, s_hi}, inc = {i_lo, i_hi};
    rngn_tables t = tables();
    unsigned status = 0;
    std::vector<rngn_rec> rec(n_chunks);
    std::vector<int32_t> entry(n_chunks), nout(n_chunks);
    std::vector<int64_t> prefix(n_chunks + 1);
    for (int64_t c = 0; c < n_chunks; ++c) rec[c] = rngn_table_chunk(rngn_advance(s0, inc, (uint64_t)c * RNGN_CHUNK), inc, t, &status);
    int fail = 0;
    for (int64_t c = 0; c < n_chunks; ++c) {
        int e, k;
        rngn_resolve_chunk(rec.data(), c, s0, inc, t, &e, &k, &status, &fail);
        entry[c] = e;
        nout[c] = k;
    }
    *failed = fail;
    if (fail || force_serial) {
        rngn_resolve_serial(rec.data(), n_chunks, s0, inc, t, entry.data(), nout.data(), &status);
        status |= RNGN_ST_SERIAL;
    }
    prefix[0] = 0;
    for (int64_t c = 0; c < n_chunks; ++c) prefix[c + 1] = prefix[c] + nout[c];
    if (prefix[n_chunks] < n) status |= RNGN_ST_SHORT;
    Sink sink = {out, row_words, P, row_state};
    row_words[0] = 0;
    row_state[0] = s_lo;
    row_state[1] = s_hi;
    for (int64_t c = 0; c < n_chunks; ++c) rngn_emit_chunk(rngn_advance(s0, inc, (uint64_t)c * RNGN_CHUNK), inc, t, c, entry[c], prefix[c], n, &status, sink);
    return (int)status;
}

void rngn_host_set_fused(int f) { g_fused = f; }

int rngn_host_chunk_words(void) { return RNGN_CHUNK; }

double rngn_host_log1p_neg(double x, int fused) { return rngn_log1p_neg(x, fused); }

double rngn_host_exp_neg(double x, int fused) { return rngn_exp_neg(x, fused, EXP_TAB); }

void rngn_host_advance(uint64_t s_lo, uint64_t s_hi, uint64_t i_lo, uint64_t i_hi, uint64_t delta, uint64_t* out) {
    rngn_u128 s = {s_lo, s_hi}, inc = {i_lo, i_hi};
    rngn_u128 r = rngn_advance(s, inc, delta);
    out[0] = r.lo;
    out[1] = r.hi;
}
}
