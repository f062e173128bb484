"""SURVEY.md §8(f) row N4 (first half), -m gpu: RNGNoiseSource drawn on the device (csrc/rng_normal.cu through
dfd_rng_normal_rows) against numpy's own Generator(PCG64).standard_normal - the reference's noise source
(utils/noise_sources.py:4-20) IS numpy.  Bit-exact bar: every normal (fp64), every key, the worker's members
fp32(fp64(flat) + sigma * eps), the generator's state after the batch."""
import numpy as np
import pytest
import torch

from dfd_starter_b200.device import get_context

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def D():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import __graft_entry__ as G
    G.build()
    import dfd_starter_b200 as D
    return D


def _state(rng):
    st = rng.bit_generator.state["state"]
    return int(st["state"]), int(st["inc"])


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint64 if a.dtype == np.float64 else np.uint32)


@pytest.mark.parametrize("n_rows,P,force_serial", [(24, 6406, False), (3, 171042, False), (40, 1000, True), (1, 1, False),
                                                    (7, 33, False)])
def test_one_stream_rows_are_numpys(D, n_rows, P, force_serial):
    """the worker's form: ONE stream, consecutive sample() calls = consecutive rows.  fp64 normals, fp32 casts and the
    key of every row (stream state at the row boundary) identical to numpy's."""
    from dfd_starter_b200.noise_sources import device_normal_rows, pcg64_advance
    ctx = get_context(0)
    rng = np.random.default_rng(np.random.SeedSequence(321 + n_rows))
    s, inc = _state(rng)
    keys, ref = [], []
    for _ in range(n_rows):
        keys.append(_state(rng)[0])
        ref.append(rng.standard_normal(P))
    keys.append(_state(rng)[0])
    ref = np.stack(ref)
    Ps = (P + 3) // 4 * 4
    out, out64, words, status = device_normal_rows(ctx, [(s, inc)], n_rows, P, row_stride=Ps, want_f64=True,
                                                   force_serial=force_serial)
    assert (status & 7) == 0 and bool(status & 8) == force_serial
    got64 = out64.cpu().numpy().reshape(n_rows, Ps)[:, :P]
    got32 = out.cpu().numpy().reshape(n_rows, Ps)[:, :P]
    assert np.array_equal(_bits(got64), _bits(ref))
    assert np.array_equal(_bits(got32), _bits(ref.astype(np.float32)))
    assert [pcg64_advance(s, inc, w) for w in words[0]] == keys
    assert [words.state(0, r) for r in range(n_rows + 1)] == keys


def test_many_streams_decode_form(D):
    """the learner's form: one stream per return key, one row each (decode, finite_differences.py:87), scattered to
    chosen rows."""
    from dfd_starter_b200.noise_sources import device_normal_rows, pcg64_advance
    ctx = get_context(0)
    n, P = 300, 2049
    seeds = np.random.default_rng(8).integers(0, 2 ** 62, n)
    streams, ref, ends = [], [], []
    for sd in seeds:
        g = np.random.default_rng(int(sd))
        streams.append(_state(g))
        ref.append(g.standard_normal(P))
        ends.append(_state(g)[0])
    dest = np.random.RandomState(0).permutation(n + 5)[:n].astype(np.int32)
    out, out64, words, status = device_normal_rows(ctx, streams, 1, P, row_stride=P + 3, dest_row=dest, want_f64=True)
    assert (status & 7) == 0
    got = out64.cpu().numpy().reshape(-1, P + 3)
    for j in range(n):
        assert np.array_equal(_bits(got[dest[j], :P]), _bits(ref[j])), j
        assert pcg64_advance(streams[j][0], streams[j][1], words[j][1]) == ends[j]
        assert words[j][0] == 0 and words.state(j, 0) == streams[j][0] and words.state(j, 1) == ends[j]


def test_worker_members_built_in_kernel(D):
    """worker/worker.py:28 + policies/policy.py:40-42 with fp64 noise: fp32(fp64(flat) + sigma * eps), product and sum
    rounded separately - bit-identical to numpy's expression; RNGNoiseSource.sample_rows returns sample()'s keys and
    leaves the generator where n sample() calls leave it."""
    ctx = get_context(0)
    P, n, sigma = 6092, 17, 0.02
    flat = (np.random.RandomState(3).randn(P) * 0.3).astype(np.float32)
    src, ref = D.RNGNoiseSource(P, 55), D.RNGNoiseSource(P, 55, device=False)
    out = torch.zeros(n * P, dtype=torch.float32, device=ctx.device)
    keys = src.sample_rows(ctx, n, out, P, theta=torch.from_numpy(flat).to(ctx.device), sigma=sigma)
    got = out.cpu().numpy().reshape(n, P)
    for j in range(n):
        key, eps = ref.sample()
        assert keys[j] == key
        assert np.array_equal(_bits(got[j]), _bits((flat + sigma * eps).astype(np.float32))), j
    assert _state(src.rng) == _state(ref.rng)
    # decode_rows: the learner's redraw from the keys, and the generator's state after it
    out2 = torch.zeros(n * P, dtype=torch.float32, device=ctx.device)
    src.decode_rows(ctx, keys[::-1], out2, P)
    got2 = out2.cpu().numpy().reshape(n, P)
    for j, k in enumerate(keys[::-1]):
        assert np.array_equal(_bits(got2[j]), _bits(np.asarray(ref.decode(k), dtype=np.float32))), j
    assert _state(src.rng) == _state(ref.rng)


def test_long_stream_and_throughput(D):
    """a Humanoid-sized slice: 64 rows x 171 042 normals from one stream (11 M normals, 170 tail-loop rounds, several
    CTA-boundary crossings per row) bit-identical; prints the device rate next to numpy's."""
    import time
    from dfd_starter_b200.noise_sources import device_normal_rows
    ctx = get_context(0)
    n_rows, P = 64, 171042
    rng = np.random.default_rng(np.random.SeedSequence(99))
    s, inc = _state(rng)
    t0 = time.perf_counter()
    ref = rng.standard_normal(n_rows * P)
    t_np = time.perf_counter() - t0
    out = torch.zeros(n_rows * P, dtype=torch.float32, device=ctx.device)
    device_normal_rows(ctx, [(s, inc)], n_rows, P, out=out, row_stride=P)             # warm-up (jump table, attributes)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out, _, words, status = device_normal_rows(ctx, [(s, inc)], n_rows, P, out=out, row_stride=P)
    e1.record()
    torch.cuda.synchronize()
    assert (status & 7) == 0
    assert np.array_equal(_bits(out.cpu().numpy()), _bits(ref.astype(np.float32)))
    print("\nrng rows: %d normals, device %.3f ms (incl. host round trips), numpy %.1f ms" % (n_rows * P, e0.elapsed_time(e1), t_np * 1e3))


def test_more_streams_than_one_launch_serves(D):
    """70 000 return keys in one decode (grid.y of a launch is 65 535): split into launches by the host wrapper, rows and
    end states still numpy's."""
    from dfd_starter_b200.noise_sources import device_normal_rows
    ctx = get_context(0)
    n, P = 70_000, 5
    base = np.random.default_rng(4)
    s0, inc = _state(base)
    streams = [(s0 + 977 * j, inc) for j in range(n)]                 # arbitrary distinct states of one sequence
    out = torch.zeros(n * 8, dtype=torch.float32, device=ctx.device)
    out, _, marks, status = device_normal_rows(ctx, streams, 1, P, out=out, row_stride=8)
    assert (status & 7) == 0
    got = out.cpu().numpy().reshape(n, 8)
    g = np.random.default_rng(0)
    for j in (0, 1, 65534, 65535, 65536, n - 1):
        g.bit_generator.state = {"bit_generator": "PCG64", "state": {"state": streams[j][0] % (1 << 128), "inc": inc}, "has_uint32": 0, "uinteger": 0}
        ref = g.standard_normal(P)
        assert np.array_equal(_bits(got[j, :P]), _bits(ref.astype(np.float32))), j
        assert marks.state(j, 1) == _state(g)[0]


def test_learner_takes_keyed_batches_as_arrays(D):
    """FiniteDifferences.step on the batched Worker's untouched ReturnBatch (keys + arrays, no per-return objects) = the
    per-record form of the reference loop: same gradient bits, same theta, same generator state afterwards."""
    import contextlib
    import io
    P = D.MujocoPolicy(17, 6, seed=124, device=0).num_params

    class Omega(object):
        omega, min_omega, max_omega = 0.3, 0.0, 1.0

    def run(as_arrays):
        torch.manual_seed(124)
        pol = D.MujocoPolicy(17, 6, seed=124, device=0)
        src = D.RNGNoiseSource(P, 77)
        opt = D.DSGD([torch.nn.Parameter(torch.zeros(1))], lr=0.01)
        opt.coef = np.sqrt(P)
        fd = D.FiniteDifferences(pol, opt, Omega(), src, noise_std=0.02, batch_size=40, max_delayed_return=3)
        w = D.Worker(pol, D.SyntheticAgent(pol, 4, seed=1), src, None, sigma=0.02, eval_prob=0.2, random_seed=5)
        grads = []
        for _ in range(3):
            w.epoch = fd.epoch
            rets = w.collect_returns(40)
            batch = rets.non_eval() if as_arrays else [r for r in rets if not r.is_eval]
            with contextlib.redirect_stdout(io.StringIO()):
                fd.step(batch, 0.0, 0.0, 0.0)
            grads.append(np.array(fd.gradient_memory, copy=True))
        return grads, pol.get_trainable_flat(), _state(src.rng)

    ga, ta, sa = run(True)
    gr, tr, sr = run(False)
    for a, b in zip(ga, gr):
        assert np.array_equal(_bits(a.astype(np.float32)), _bits(b.astype(np.float32)))
    assert np.array_equal(_bits(ta), _bits(tr)) and sa == sr
