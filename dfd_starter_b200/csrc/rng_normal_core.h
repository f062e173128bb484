/* rng_normal_core.h - numpy's Generator(PCG64).standard_normal restated so that a GPU can draw it in parallel.
 *
 * Replaces the per-return host loop of the reference's DEFAULT noise source (utils/noise_sources.py:4-20,
 * RNGNoiseSource: key = "state,inc" of a PCG64, noise = rng.standard_normal(P)).  The algorithm lives in numpy
 * (2.3.5 here; random/src/pcg64/pcg64.h, random/src/distributions/distributions.c:random_standard_normal), which is a
 * dependency and not part of /root/reference, so its PUBLISHED algorithm is restated:
 *   - PCG64 = PCG XSL-RR 128/64: state <- state * M + inc (mod 2^128), output rotr64(hi ^ lo, state >> 122), the
 *     output being taken from the NEW state; next_double = (u64 >> 11) * 2^-53;
 *   - standard normal = Marsaglia-Tsang ziggurat with 256 layers: one 64-bit word gives the layer (8 bits), the sign
 *     (1 bit) and a 52-bit abscissa; 98.8 % of the attempts return on the first comparison; the wedge test draws one
 *     more word, the tail (layer 0) draws pairs of words until accepted; a rejected wedge starts over with a new word.
 * How many words a normal consumes is data dependent, so the stream is sequential as written.  What makes it parallel:
 * an ATTEMPT that starts at word i is a pure function of words i, i+1, ... (LCG jump-ahead gives any word in
 * O(log i)), so every 32-word chunk can be simulated on its own from its first word ("path 0"), and the true path
 * joins path 0 at the first word both visit (rng_normal.cu resolves the chunk entries and verifies the join).
 *
 * The same source compiles for the device (nvcc) and for the host (g++, tests/native/rng_core_host.cpp: the CPU test
 * that pins this file against numpy itself).  Every floating-point operation is a single IEEE double operation in the
 * order of numpy's C source: on the device the explicit _rn intrinsics keep nvcc from contracting a*b+c into an FMA.
 * The two libm calls of the algorithm - log1p in the tail, exp in the wedge test - are glibc's own operation sequences
 * restated (rngn_log1p_neg, rngn_exp_neg): chains of IEEE operations and table look-ups, hence bit-identical on both
 * sides, so no draw is ever "close enough": every comparison is decided exactly as numpy decides it. */
#ifndef DFD_RNG_NORMAL_CORE_H
#define DFD_RNG_NORMAL_CORE_H

#include <stdint.h>
#include <math.h>
#include "ziggurat_tables.h"

#if defined(__CUDACC__)
#define RNGN_HD __host__ __device__ __forceinline__
#else
#define RNGN_HD static inline
#endif

#if defined(__CUDA_ARCH__)
#define RNGN_MUL(a, b) __dmul_rn((a), (b))
#define RNGN_ADD(a, b) __dadd_rn((a), (b))
#define RNGN_SUB(a, b) __dsub_rn((a), (b))
#define RNGN_DIV(a, b) __ddiv_rn((a), (b))
#define RNGN_FMA(a, b, c) __fma_rn((a), (b), (c))
#define RNGN_UMULHI(a, b) __umul64hi((a), (b))
#else
#define RNGN_FMA(a, b, c) __builtin_fma((a), (b), (c))
#define RNGN_MUL(a, b) ((a) * (b))
#define RNGN_ADD(a, b) ((a) + (b))
#define RNGN_SUB(a, b) ((a) - (b))
#define RNGN_DIV(a, b) ((a) / (b))
#define RNGN_UMULHI(a, b) ((uint64_t)(((unsigned __int128)(a) * (unsigned __int128)(b)) >> 64))
#endif

#ifndef RNGN_CHUNK
#define RNGN_CHUNK 32            /* words per chunk (one thread of the table / emit kernels); <= 32 (bit masks).  The
                                    CPU test also builds this file with tiny chunks to stress the entry resolution */
#endif
#define RNGN_TAIL_CAP 60         /* a tail loop longer than this (probability < 1e-60) is reported, not followed */

/* status bits (device word, OR-ed) */
#define RNGN_ST_TAILCAP 2u       /* tail loop cap hit */
#define RNGN_ST_SHORT 4u         /* the word budget ended before n_draws normals were produced */
#define RNGN_ST_SERIAL 8u        /* (informational) the chunk entries were resolved by the serial fallback */

typedef struct rngn_u128 {
    uint64_t lo, hi;
} rngn_u128;

/* PCG_DEFAULT_MULTIPLIER_128 = 47026247687942121848144207491837523525 */
#define RNGN_MULT_HI 0x2360ED051FC65DA4ull
#define RNGN_MULT_LO 0x4385DF649FCCF645ull

RNGN_HD rngn_u128 rngn_mul(rngn_u128 a, rngn_u128 b) {
    rngn_u128 r;
    r.lo = a.lo * b.lo;
    r.hi = RNGN_UMULHI(a.lo, b.lo) + a.lo * b.hi + a.hi * b.lo;
    return r;
}

RNGN_HD rngn_u128 rngn_add(rngn_u128 a, rngn_u128 b) {
    rngn_u128 r;
    r.lo = a.lo + b.lo;
    r.hi = a.hi + b.hi + (r.lo < a.lo ? 1ull : 0ull);
    return r;
}

/* one LCG step: pcg_setseq_128_step_r */
RNGN_HD rngn_u128 rngn_step(rngn_u128 s, rngn_u128 inc) {
    rngn_u128 m;
    m.lo = RNGN_MULT_LO;
    m.hi = RNGN_MULT_HI;
    return rngn_add(rngn_mul(s, m), inc);
}

/* pcg_output_xsl_rr_128_64 */
RNGN_HD uint64_t rngn_output(rngn_u128 s) {
    uint64_t x = s.hi ^ s.lo;
    unsigned rot = (unsigned)(s.hi >> 58);
    return (x >> rot) | (x << ((64u - rot) & 63u));
}

/* state after `delta` steps (pcg_advance_lcg_128: Brown's O(log delta) jump-ahead) */
RNGN_HD rngn_u128 rngn_advance(rngn_u128 s, rngn_u128 inc, uint64_t delta) {
    rngn_u128 acc_mult, acc_plus, cur_mult, cur_plus, one;
    acc_mult.lo = 1; acc_mult.hi = 0;
    acc_plus.lo = 0; acc_plus.hi = 0;
    cur_mult.lo = RNGN_MULT_LO; cur_mult.hi = RNGN_MULT_HI;
    cur_plus = inc;
    one.lo = 1; one.hi = 0;
    while (delta > 0) {
        if (delta & 1) {
            acc_mult = rngn_mul(acc_mult, cur_mult);
            acc_plus = rngn_add(rngn_mul(acc_plus, cur_mult), cur_plus);
        }
        cur_plus = rngn_mul(rngn_add(cur_mult, one), cur_plus);
        cur_mult = rngn_mul(cur_mult, cur_mult);
        delta >>= 1;
    }
    return rngn_add(rngn_mul(acc_mult, s), acc_plus);
}

RNGN_HD double rngn_bits_to_double(uint64_t b) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)b);
#else
    union { uint64_t u; double d; } c;
    c.u = b;
    return c.d;
#endif
}

RNGN_HD uint64_t rngn_double_to_bits(double d) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(d);
#else
    union { uint64_t u; double d; } c;
    c.d = d;
    return c.u;
#endif
}

RNGN_HD int32_t rngn_hi32(double d) { return (int32_t)(rngn_double_to_bits(d) >> 32); }

RNGN_HD double rngn_with_hi32(double d, int32_t hi) {
    return rngn_bits_to_double((rngn_double_to_bits(d) & 0xffffffffull) | ((uint64_t)(uint32_t)hi << 32));
}

/* log1p(x) for x in (-1, 0]: the operation sequence of glibc 2.39's dbl-64 log1p (sysdeps/ieee754/dbl-64/s_log1p.c:
 * fdlibm's algorithm with the degree-7 polynomial in the split form R1 + z2*R2 + z4*R3 + z6*R4).  glibc ships TWO
 * builds of that one source and picks by CPU (sysdeps/x86_64/fpu/multiarch/s_log1p.c, ifunc): the plain one and one
 * compiled with -mfma -mavx2 in which gcc contracted thirteen multiply-adds into FMAs.  Which roundings were fused is a
 * property of that binary (read off libm.so.6's `__log1p_fma`; pinned by tests/test_rng_core_cpu.py against
 * math.log1p on this machine, and against the plain build through GLIBC_TUNABLES=glibc.cpu.hwcaps=-FMA): `fused`
 * selects it, and the host probes which build its libm uses (noise_sources.libm_fused).
 * Only the branches reachable from x = -u, u in [0, 1) a multiple of 2^-53, are kept. */
RNGN_HD double rngn_log1p_neg(double x, int fused) {
    const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10;
    const double Lp1 = 6.666666666666735130e-01, Lp2 = 3.999999999940941908e-01, Lp3 = 2.857142874366239149e-01,
                 Lp4 = 2.222219843214978396e-01, Lp5 = 1.818357216161805012e-01, Lp6 = 1.531383769920937332e-01,
                 Lp7 = 1.479819860511658591e-01;
    double hfsq, f = 0.0, c = 0.0, s, z, R, u, z2, z4, z6, R2, R3, R4, kd, c2, v, w;
    int32_t k, hx, hu = 0, ax;
    hx = rngn_hi32(x);
    ax = hx & 0x7fffffff;
    k = 1;
    if (ax < 0x3e200000) {                       /* |x| < 2^-29 */
        if (ax < 0x3c900000) return x;           /* |x| < 2^-54: only x = -0 (u = 0) gets here */
        if (fused) return RNGN_FMA(-RNGN_MUL(x, x), 0.5, x);
        return RNGN_SUB(x, RNGN_MUL(RNGN_MUL(x, x), 0.5));
    }
    if (hx > 0 || hx <= (int32_t)0xbfd2bec3) {   /* -0.2929 < x < 0.41422 */
        k = 0;
        f = x;
        hu = 1;
    }
    if (k != 0) {
        u = RNGN_ADD(1.0, x);
        hu = rngn_hi32(u);
        k = (hu >> 20) - 1023;
        c = (k > 0) ? RNGN_SUB(1.0, RNGN_SUB(u, x)) : RNGN_SUB(x, RNGN_SUB(u, 1.0));
        c = RNGN_DIV(c, u);
        hu &= 0x000fffff;
        if (hu < 0x6a09e) {
            u = rngn_with_hi32(u, hu | 0x3ff00000);
        } else {
            k += 1;
            u = rngn_with_hi32(u, hu | 0x3fe00000);
            hu = (0x00100000 - hu) >> 2;
        }
        f = RNGN_SUB(u, 1.0);
    }
    kd = (double)k;
    hfsq = RNGN_MUL(RNGN_MUL(0.5, f), f);
    if (hu == 0) {                               /* |f| < 2^-20 */
        if (f == 0.0) {
            if (k == 0) return 0.0;
            if (fused) return RNGN_FMA(kd, ln2_hi, RNGN_FMA(kd, ln2_lo, c));
            c = RNGN_ADD(c, RNGN_MUL(kd, ln2_lo));
            return RNGN_ADD(RNGN_MUL(kd, ln2_hi), c);
        }
        if (fused) {
            R = RNGN_MUL(RNGN_FMA(-0.66666666666666666, f, 1.0), hfsq);
            if (k == 0) return RNGN_SUB(f, R);
            return RNGN_FMA(kd, ln2_hi, -RNGN_SUB(RNGN_SUB(R, RNGN_FMA(kd, ln2_lo, c)), f));
        }
        R = RNGN_MUL(hfsq, RNGN_SUB(1.0, RNGN_MUL(0.66666666666666666, f)));
        if (k == 0) return RNGN_SUB(f, R);
        return RNGN_SUB(RNGN_MUL(kd, ln2_hi), RNGN_SUB(RNGN_SUB(R, RNGN_ADD(RNGN_MUL(kd, ln2_lo), c)), f));
    }
    s = RNGN_DIV(f, RNGN_ADD(2.0, f));
    z = RNGN_MUL(s, s);
    z2 = RNGN_MUL(z, z);
    z4 = RNGN_MUL(z2, z2);
    z6 = RNGN_MUL(z4, z2);
    if (fused) {
        R2 = RNGN_FMA(z, Lp3, Lp2);
        R3 = RNGN_FMA(z, Lp5, Lp4);
        R4 = RNGN_FMA(z, Lp7, Lp6);
        R = RNGN_FMA(z6, R4, RNGN_FMA(z4, R3, RNGN_FMA(z, Lp1, RNGN_MUL(z2, R2))));
        v = RNGN_MUL(RNGN_ADD(R, hfsq), s);
        if (k == 0) return RNGN_SUB(f, RNGN_SUB(hfsq, v));
        c2 = RNGN_FMA(kd, ln2_lo, c);
        w = RNGN_SUB(RNGN_SUB(hfsq, RNGN_ADD(c2, v)), f);
        return RNGN_FMA(kd, ln2_hi, -w);
    }
    R2 = RNGN_ADD(Lp2, RNGN_MUL(z, Lp3));
    R3 = RNGN_ADD(Lp4, RNGN_MUL(z, Lp5));
    R4 = RNGN_ADD(Lp6, RNGN_MUL(z, Lp7));
    R = RNGN_ADD(RNGN_ADD(RNGN_ADD(RNGN_MUL(z, Lp1), RNGN_MUL(z2, R2)), RNGN_MUL(z4, R3)), RNGN_MUL(z6, R4));
    v = RNGN_MUL(s, RNGN_ADD(hfsq, R));
    if (k == 0) return RNGN_SUB(f, RNGN_SUB(hfsq, v));
    return RNGN_SUB(RNGN_MUL(kd, ln2_hi), RNGN_SUB(RNGN_SUB(hfsq, RNGN_ADD(v, RNGN_ADD(RNGN_MUL(kd, ln2_lo), c))), f));
}

/* exp(x) for x in [-8, 0]: glibc 2.39's dbl-64 exp (sysdeps/ieee754/dbl-64/e_exp.c - Szabolcs Nagy's table-driven
 * algorithm from ARM's optimized routines: k = round(x * 128/ln2) through the 1.5 * 2^52 shift, r = x - k * ln2/128 in two
 * pieces, 2^(k/128) from a 128-entry table (tab), degree-5 polynomial), in the operation order of the two builds glibc
 * selects from by CPU: `fused` = the -mfma build (read off libm.so.6's `__exp_fma`: seven FMAs), else the plain build.
 * Only what x = -0.5 * v * v, |v| < 3.66, reaches: the main path and the |x| < 2^-54 shortcut. */
RNGN_HD double rngn_exp_neg(double x, int fused, const uint64_t* tab) {
    const double InvLn2N = 0x1.71547652b82fep7, Shift = 0x1.8p52, NegLn2hiN = -0x1.62e42fefa0000p-8,
                 NegLn2loN = -0x1.cf79abc9e3b3ap-47, C2 = 0x1.ffffffffffdbdp-2, C3 = 0x1.555555555543cp-3,
                 C4 = 0x1.55555cf172b91p-5, C5 = 0x1.1111167a4d017p-7;
    uint32_t abstop = (uint32_t)(rngn_double_to_bits(x) >> 52) & 0x7ff;
    if (abstop < 0x3c9) return RNGN_ADD(1.0, x);             /* |x| < 2^-54 */
    double kd, r, r2, tmp, scale;
    uint64_t ki, idx, sbits;
    kd = fused ? RNGN_FMA(x, InvLn2N, Shift) : RNGN_ADD(RNGN_MUL(InvLn2N, x), Shift);
    ki = rngn_double_to_bits(kd);
    kd = RNGN_SUB(kd, Shift);
    if (fused) r = RNGN_FMA(kd, NegLn2loN, RNGN_FMA(kd, NegLn2hiN, x));
    else r = RNGN_ADD(RNGN_ADD(x, RNGN_MUL(kd, NegLn2hiN)), RNGN_MUL(kd, NegLn2loN));
    idx = 2 * (ki & 127);
    sbits = tab[idx + 1] + (ki << 45);
    r2 = RNGN_MUL(r, r);
    if (fused) {
        tmp = RNGN_FMA(RNGN_MUL(r2, r2), RNGN_FMA(r, C5, C4),
                       RNGN_FMA(RNGN_FMA(r, C3, C2), r2, RNGN_ADD(r, rngn_bits_to_double(tab[idx]))));
        scale = rngn_bits_to_double(sbits);
        return RNGN_FMA(scale, tmp, scale);
    }
    tmp = RNGN_ADD(RNGN_ADD(RNGN_ADD(rngn_bits_to_double(tab[idx]), r), RNGN_MUL(r2, RNGN_ADD(C2, RNGN_MUL(r, C3)))),
                   RNGN_MUL(RNGN_MUL(r2, r2), RNGN_ADD(C4, RNGN_MUL(r, C5))));
    scale = rngn_bits_to_double(sbits);
    return RNGN_ADD(scale, RNGN_MUL(scale, tmp));
}

/* the ziggurat tables as seen by the simulation: pointers so that the device can keep them in shared memory */
typedef struct rngn_tables {
    const uint64_t* ki;
    const double* wi;
    const double* fi;
    const uint64_t* exp_tab;   /* glibc's __exp_data.tab */
    int libm_fused;    /* which of glibc's two builds of log1p / exp the host's numpy calls (1 = -mfma) */
} rngn_tables;

typedef struct rngn_attempt {
    int len;        /* words consumed */
    int out;        /* 1 = a normal was returned, 0 = rejected wedge (the next word starts a fresh attempt) */
    double val;
} rngn_attempt;

#define RNGN_NOR_R 3.6541528853610087963519472518
#define RNGN_NOR_INV_R 0.27366123732975827203338247596

RNGN_HD double rngn_next_double(rngn_u128* s, rngn_u128 inc) {
    *s = rngn_step(*s, inc);
    return RNGN_MUL((double)(rngn_output(*s) >> 11), 1.0 / 9007199254740992.0);
}

/* One attempt of random_standard_normal's outer loop starting with the NEXT word of the stream (*s is the state before
 * that word; on return it is the state after the attempt's last word). */
RNGN_HD rngn_attempt rngn_attempt_at(rngn_u128* s, rngn_u128 inc, const rngn_tables t, unsigned* status) {
    rngn_attempt a;
    *s = rngn_step(*s, inc);
    uint64_t r = rngn_output(*s);
    int idx = (int)(r & 0xff);
    r >>= 8;
    int sign = (int)(r & 0x1);
    uint64_t rabs = (r >> 1) & 0x000fffffffffffffull;
    double x = RNGN_MUL((double)rabs, t.wi[idx]);
    if (sign) x = -x;
    a.len = 1;
    a.out = 1;
    a.val = x;
    if (rabs < t.ki[idx]) return a;
    if (idx == 0) {
        for (int it = 0;; ++it) {
            double xx = RNGN_MUL(-RNGN_NOR_INV_R, rngn_log1p_neg(-rngn_next_double(s, inc), t.libm_fused));
            double yy = -rngn_log1p_neg(-rngn_next_double(s, inc), t.libm_fused);
            a.len += 2;
            if (RNGN_ADD(yy, yy) > RNGN_MUL(xx, xx)) {
                a.val = ((rabs >> 8) & 0x1) ? -RNGN_ADD(RNGN_NOR_R, xx) : RNGN_ADD(RNGN_NOR_R, xx);
                return a;
            }
            if (it >= RNGN_TAIL_CAP) {
                *status |= RNGN_ST_TAILCAP;
                a.val = 0.0;
                return a;
            }
        }
    }
    double u = rngn_next_double(s, inc);
    a.len = 2;
    double lhs = RNGN_ADD(RNGN_MUL(RNGN_SUB(t.fi[idx - 1], t.fi[idx]), u), t.fi[idx]);
    double rhs = rngn_exp_neg(RNGN_MUL(RNGN_MUL(-0.5, x), x), t.libm_fused, t.exp_tab);
    a.out = lhs < rhs ? 1 : 0;
    return a;
}

/* Path of one chunk entered at word offset `e` (e < RNGN_CHUNK): attempts start at e, e + len, ... while < RNGN_CHUNK.
 * s_chunk = state before the chunk's first word.  start_mask / out_mask: bit p set when an attempt starts at word p /
 * starts there and returns a normal.  Returns the exit offset (words of the next chunks already consumed). */
RNGN_HD int rngn_chunk_path(rngn_u128 s_chunk, rngn_u128 inc, const rngn_tables t, int e, uint32_t* start_mask,
                            uint32_t* out_mask, unsigned* status) {
    rngn_u128 s = s_chunk;
    for (int i = 0; i < e; ++i) s = rngn_step(s, inc);
    int pos = e;
    uint32_t sm = 0, om = 0;
    while (pos < RNGN_CHUNK) {
        rngn_attempt a = rngn_attempt_at(&s, inc, t, status);
        sm |= 1u << pos;
        if (a.out) om |= 1u << pos;
        pos += a.len;
    }
    *start_mask = sm;
    *out_mask = om;
    return pos - RNGN_CHUNK;
}

/* ---- the three per-chunk steps of the parallel form (kernels in rng_normal.cu; the host test runs the same functions) ---- */

typedef struct rngn_rec {           /* path 0 of a chunk: 16 bytes */
    uint32_t start_mask, out_mask;
    int32_t exit0;
    int32_t pad;
} rngn_rec;

RNGN_HD int rngn_popc(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return __popc(v);
#else
    return __builtin_popcount(v);
#endif
}

/* step 1: path 0 of a chunk; s_chunk = state before the chunk's first word */
RNGN_HD rngn_rec rngn_table_chunk(rngn_u128 s_chunk, rngn_u128 inc, const rngn_tables t, unsigned* status) {
    rngn_rec r;
    r.exit0 = rngn_chunk_path(s_chunk, inc, t, 0, &r.start_mask, &r.out_mask, status);
    r.pad = 0;
    return r;
}

/* step 2: entry offset and number of normals of chunk c, SPECULATING that the previous chunk was left through its
 * path 0 (entry = exit0 of chunk c-1), and verifying for this chunk that its own exit under that entry is its path-0
 * exit.  If every chunk verifies, induction from chunk 0 (entry 0) makes every speculated entry the true one.
 * *fail is set when this chunk does not verify (the serial resolver then redoes the stream). */
RNGN_HD void rngn_resolve_chunk(const rngn_rec* rec, int64_t c, rngn_u128 s0, rngn_u128 inc, const rngn_tables t,
                                int* entry, int* nout, unsigned* status, int* fail) {
    int e = c == 0 ? 0 : rec[c - 1].exit0;
    rngn_rec r = rec[c];
    *entry = e;
    if (e < RNGN_CHUNK && ((r.start_mask >> e) & 1u)) {
        *nout = rngn_popc(r.out_mask >> e);
        return;
    }
    if (e >= RNGN_CHUNK) {
        *nout = 0;
        if (e - RNGN_CHUNK != r.exit0) *fail = 1;
        return;
    }
    uint32_t sm, om;
    rngn_u128 s = rngn_advance(s0, inc, (uint64_t)c * RNGN_CHUNK);
    int ex = rngn_chunk_path(s, inc, t, e, &sm, &om, status);
    *nout = rngn_popc(om);
    if (ex != r.exit0) *fail = 1;
}

/* serial resolver of one stream (only when a chunk failed to verify): follows the true path chunk by chunk */
RNGN_HD void rngn_resolve_serial(const rngn_rec* rec, int64_t n_chunks, rngn_u128 s0, rngn_u128 inc, const rngn_tables t,
                                 int32_t* entry, int32_t* nout, unsigned* status) {
    int64_t e = 0;
    for (int64_t c = 0; c < n_chunks; ++c) {
        rngn_rec r = rec[c];
        entry[c] = (int32_t)e;
        if (e >= RNGN_CHUNK) {
            nout[c] = 0;
            e -= RNGN_CHUNK;
        } else if ((r.start_mask >> e) & 1u) {
            nout[c] = rngn_popc(r.out_mask >> e);
            e = r.exit0;
        } else {
            uint32_t sm, om;
            rngn_u128 s = rngn_advance(s0, inc, (uint64_t)c * RNGN_CHUNK);
            e = rngn_chunk_path(s, inc, t, (int)e, &sm, &om, status);
            nout[c] = rngn_popc(om);
        }
    }
}

/* step 3: replay chunk c from its entry and hand every normal to sink(g, value, words_after, state_after) - g = index
 * of the normal in its stream, words_after / state_after = stream words consumed / generator state once this normal has
 * been returned (the next normal's key) */
template <typename Sink>
RNGN_HD void rngn_emit_chunk(rngn_u128 s_chunk, rngn_u128 inc, const rngn_tables t, int64_t c, int entry, int64_t first_g,
                             int64_t n_draws, unsigned* status, Sink& sink) {
    if (entry >= RNGN_CHUNK || first_g >= n_draws) return;
    rngn_u128 s = s_chunk;
    for (int i = 0; i < entry; ++i) s = rngn_step(s, inc);
    int pos = entry;
    int64_t g = first_g;
    while (pos < RNGN_CHUNK && g < n_draws) {
        rngn_attempt a = rngn_attempt_at(&s, inc, t, status);
        pos += a.len;
        if (a.out) {
            sink(g, a.val, c * RNGN_CHUNK + pos, s);
            ++g;
        }
    }
}

#endif
