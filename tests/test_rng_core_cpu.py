"""SURVEY.md §8(f) row N4 (first half): the device restatement of numpy's Generator(PCG64).standard_normal
(csrc/rng_normal_core.h - what csrc/rng_normal.cu runs per chunk) built for the HOST and pinned against numpy itself,
the reference's own generator (utils/noise_sources.py:4-20): sequential form, the parallel chunk form (speculative
entry resolution, serial resolver, row keys), and the two glibc log1p builds.  No GPU; the oracle here is numpy."""
import ctypes as C
import math
import os
import struct
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
u64, i64 = C.c_uint64, C.c_int64


def _build(tmp, chunk=None):
    out = os.path.join(tmp, "librngcore%s.so" % (chunk or ""))
    cmd = ["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-I", os.path.join(ROOT, "dfd_starter_b200", "csrc"),
           os.path.join(ROOT, "tests", "native", "rng_core_host.cpp"), "-o", out]
    if chunk:
        cmd.insert(1, "-DRNGN_CHUNK=%d" % chunk)
    subprocess.check_call(cmd)
    L = C.CDLL(out)
    L.rngn_host_log1p_neg.restype = C.c_double
    L.rngn_host_log1p_neg.argtypes = [C.c_double, C.c_int]
    return L


@pytest.fixture(scope="module")
def core(tmp_path_factory):
    from dfd_starter_b200.noise_sources import libm_fused
    L = _build(str(tmp_path_factory.mktemp("rngcore")))
    L.rngn_host_set_fused(libm_fused())
    return L


def _split(v):
    return u64(v & (2 ** 64 - 1)), u64(v >> 64)


def _state(rng):
    st = rng.bit_generator.state["state"]
    return st["state"], st["inc"]


def _chunked(L, s, inc, n, P, force_serial=0):
    cw = L.rngn_host_chunk_words()
    nc = (int(n * 1.04) + 1024 + cw - 1) // cw
    out, rows, failed = np.zeros(n), np.zeros(n // P + 1, dtype=np.int64), C.c_int(0)
    states = np.zeros((n // P + 1, 2), dtype=np.uint64)
    st = L.rngn_host_chunked(*_split(s), *_split(inc), i64(n), i64(P), i64(nc), force_serial, out.ctypes.data_as(C.c_void_p),
                             rows.ctypes.data_as(C.c_void_p), states.ctypes.data_as(C.c_void_p), C.byref(failed))
    # the state handed out at a row boundary must be the stream advanced by the word count handed out with it
    for w, (lo, hi) in zip(rows, states):
        adv = (u64 * 2)()
        L.rngn_host_advance(*_split(s), *_split(inc), u64(int(w)), adv)
        assert (int(adv[0]), int(adv[1])) == (int(lo), int(hi))
    return st, out, rows, failed.value


def _reference_rows(seed, n_rows, P):
    """RNGNoiseSource.sample() n_rows times: keys and noise, by numpy."""
    rng = np.random.default_rng(np.random.SeedSequence(seed))
    s, inc = _state(rng)
    keys, ref = [], []
    for _ in range(n_rows):
        keys.append(_state(rng)[0])
        ref.append(rng.standard_normal(P))
    keys.append(_state(rng)[0])
    return s, inc, keys, np.concatenate(ref)


def test_sequential_stream_is_numpys(core):
    """one attempt after the other = random_standard_normal: 4 M normals (about 1 000 tail draws through log1p, 50 000
    wedge tests through exp) bit-identical, and the generator ends in numpy's state."""
    rng = np.random.default_rng(np.random.SeedSequence(123))
    s, inc = _state(rng)
    n = 4_000_000
    ref = rng.standard_normal(n)
    out, words = np.empty(n), i64(0)
    st = core.rngn_host_sequential(*_split(s), *_split(inc), i64(n), out.ctypes.data_as(C.c_void_p), C.byref(words))
    assert st == 0
    assert np.array_equal(out.view(np.uint64), ref.view(np.uint64))
    assert 1.02 < words.value / n < 1.024
    adv = (u64 * 2)()
    core.rngn_host_advance(*_split(s), *_split(inc), u64(words.value), adv)
    assert (adv[0] | (adv[1] << 64)) == _state(rng)[0]


@pytest.mark.parametrize("seed,n_rows,P", [(123, 150, 6406), (5, 300, 1000), (9, 8, 8), (11, 1, 33), (12, 1, 1)])
def test_chunk_form_is_numpys(core, seed, n_rows, P):
    """tables -> speculative entries -> scan -> emit (the kernels' per-chunk functions in plain loops): normals
    bit-identical, every row key = numpy's state at that sample() call; same through the serial resolver."""
    s, inc, keys, ref = _reference_rows(seed, n_rows, P)
    for force_serial in (0, 1):
        st, out, rows, failed = _chunked(core, s, inc, n_rows * P, P, force_serial)
        assert (st & 7) == 0 and failed == 0
        assert np.array_equal(out.view(np.uint64), ref.view(np.uint64))
        got = []
        for w in rows:
            adv = (u64 * 2)()
            core.rngn_host_advance(*_split(s), *_split(inc), u64(int(w)), adv)
            got.append(adv[0] | (adv[1] << 64))
        assert got == keys


def test_entry_resolution_under_stress(tmp_path):
    """4-word chunks make multi-word attempts straddle chunk boundaries all the time: speculated entries fail to verify
    for some streams and the serial resolver takes over - the output must not change."""
    from dfd_starter_b200.noise_sources import libm_fused
    L = _build(str(tmp_path), chunk=4)
    L.rngn_host_set_fused(libm_fused())
    seeds = np.random.default_rng(42).integers(0, 2 ** 62, 1500)
    n_failed = 0
    for sd in seeds:
        g = np.random.default_rng(int(sd))
        s, inc = _state(g)
        ref = g.standard_normal(700)
        st, out, rows, failed = _chunked(L, s, inc, 700, 700)
        n_failed += failed
        assert (st & 7) == 0 and np.array_equal(out.view(np.uint64), ref.view(np.uint64))
        assert bool(st & 8) == bool(failed)
    assert n_failed > 0


def test_short_word_budget_is_reported(core):
    g = np.random.default_rng(1)
    s, inc = _state(g)
    out, rows, failed = np.zeros(700), np.zeros(2, dtype=np.int64), C.c_int(0)
    states = np.zeros((2, 2), dtype=np.uint64)
    st = core.rngn_host_chunked(*_split(s), *_split(inc), i64(700), i64(700), i64(22), 0, out.ctypes.data_as(C.c_void_p),
                                rows.ctypes.data_as(C.c_void_p), states.ctypes.data_as(C.c_void_p), C.byref(failed))
    assert st & 4


_LOG1P_CHECK = r"""
import ctypes as C, math, struct, sys
import numpy as np
L = C.CDLL(sys.argv[1]); fused = int(sys.argv[2])
L.rngn_host_log1p_neg.restype = C.c_double; L.rngn_host_log1p_neg.argtypes = [C.c_double, C.c_int]
L.rngn_host_exp_neg.restype = C.c_double; L.rngn_host_exp_neg.argtypes = [C.c_double, C.c_int]
sys.path.insert(0, sys.argv[3])
from dfd_starter_b200.noise_sources import libm_fused
assert libm_fused() == fused, "probe disagrees"
r = np.random.default_rng(1)
us = [r.integers(0, 2 ** 53, 400000, dtype=np.uint64) * 2.0 ** -53,
      r.integers(1, 2 ** 30, 5000, dtype=np.uint64) * 2.0 ** -53, r.integers(1, 2 ** 24, 5000, dtype=np.uint64) * 2.0 ** -53,
      0.5 + r.integers(-2 ** 12, 2 ** 12, 5000).astype(np.float64) * 2.0 ** -33,
      0.5 + r.integers(-2 ** 12, 2 ** 12, 5000).astype(np.float64) * 2.0 ** -53,
      0.75 + r.integers(-2 ** 12, 2 ** 12, 5000).astype(np.float64) * 2.0 ** -53,
      1.0 - r.integers(1, 2 ** 12, 5000).astype(np.float64) * 2.0 ** -53, np.array([0.0, 0.5, 0.25, 0.75])]
for hi in (0x3fd2bec2, 0x3fd2bec3, 0x3fd2bec4, 0x3fd2bec5):       # the k = 0 / k != 0 branch boundary (-0.2929)
    us.append(np.array([round(struct.unpack("<d", struct.pack("<Q", (hi << 32) | lo))[0] * 2 ** 53) / 2 ** 53
                        for lo in (0, 0x80000000, 0xffffffff, 0x12345678)]))
bad = sum(struct.pack("<d", L.rngn_host_log1p_neg(-float(u), fused)) != struct.pack("<d", math.log1p(-float(u)))
          for u in np.concatenate(us))
v = r.random(400000) * 3.66
xs = np.concatenate([-0.5 * v * v, -r.random(20000) * 1e-12, -r.random(1000) * 2.0 ** -53, -r.random(1000) * 2.0 ** -60, [-0.0, -6.7]])
bad_exp = sum(L.rngn_host_exp_neg(float(x), fused) != math.exp(float(x)) for x in xs)
print("log1p mismatches", bad, "exp mismatches", bad_exp)
sys.exit(1 if bad or bad_exp else 0)
"""


@pytest.mark.parametrize("fused", [1, 0])
def test_log1p_and_exp_are_glibcs_in_both_builds(tmp_path, fused):
    """rngn_log1p_neg / rngn_exp_neg against this machine's libm: 435 000 arguments of the ziggurat tail's domain
    (random, tiny, |f| < 2^-20, the branch boundary) and 420 000 of the wedge test's (-v^2/2, tiny): bit-identical -
    the -mfma builds as the process normally resolves them, the plain builds with the FMA capability masked
    (GLIBC_TUNABLES), each detected by the probe the product uses."""
    from dfd_starter_b200.noise_sources import libm_fused
    if not fused and libm_fused() == 0:
        pytest.skip("this CPU already runs the plain build (covered by fused=... of the other case)")
    _build(str(tmp_path))
    env = dict(os.environ)
    if not fused:
        env["GLIBC_TUNABLES"] = "glibc.cpu.hwcaps=-FMA,-AVX2"
    elif libm_fused() == 0:
        pytest.skip("no FMA on this CPU")
    p = subprocess.run([sys.executable, "-c", _LOG1P_CHECK, os.path.join(str(tmp_path), "librngcore.so"), str(fused), ROOT],
                       env=env, capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr


def test_pcg64_advance_helper():
    from dfd_starter_b200.noise_sources import pcg64_advance
    rng = np.random.default_rng(np.random.SeedSequence(77))
    s, inc = _state(rng)
    rng.bit_generator.random_raw(1000)
    assert pcg64_advance(s, inc, 1000) == _state(rng)[0]
