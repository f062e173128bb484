"""CPU-only host logic: noise-source interface and streams, worker draws, policy init,
state_dict layouts, DSGD lr map — all against the golden fixtures captured from the reference."""
import json
import os

import numpy as np
import pytest
import torch

from dfd_starter_b200.noise_sources import SharedNoiseTable, RNGNoiseSource, SimpleNoiseSource, parse_key
from dfd_starter_b200 import policies as P
from dfd_starter_b200.dsgd import DSGD, affine_transform, is_dsgd
from dfd_starter_b200.worker import Worker
from dfd_starter_b200.fd_return import FDReturn


@pytest.fixture(scope="module")
def noise_json(golden_dir):
    with open(os.path.join(golden_dir, "noise.json")) as f:
        return json.load(f)


def test_noise_source_interface_matches_reference(noise_json):
    g = noise_json["tables"]["1000000_6092_123"]
    t = SharedNoiseTable(g["size"], g["n_params"], g["seed"], upload=False)
    keys = []
    for _ in range(8):
        k, v = t.sample()
        assert isinstance(k, str) and v.dtype == np.float32 and v.shape == (6092,)
        assert np.shares_memory(v, t._table)          # a view, like the reference
        assert np.array_equal(t.decode(k), v)
        keys.append(k)
    assert keys == g["keys"]
    assert np.array_equal(t.decode("-%s" % keys[0]), -t.decode(keys[0]))
    assert parse_key("+17") == (17, 1) and parse_key("-17") == (17, -1) and parse_key("17") == (17, 1)
    with pytest.raises(AssertionError):
        SharedNoiseTable(10, 10, upload=False)


def test_sample_indices_is_the_same_stream(noise_json):
    g = noise_json["tables"]["200000_5197_124"]
    t = SharedNoiseTable(g["size"], g["n_params"], g["seed"], upload=False)
    assert [str(i) for i in t.sample_indices(8)] == g["keys"]


def test_rng_and_simple_noise_sources_follow_the_reference(golden_dir):
    """utils/noise_sources.py:4-33: RNGNoiseSource keys are the PCG64 `state,inc` before the draw
    (App. C words; keys recorded from the reference class itself), SimpleNoiseSource's key is the vector."""
    g = np.load(os.path.join(golden_dir, "fd_steps_hostnoise.npz"))
    src = RNGNoiseSource(8, 123)
    key, noise = src.sample()
    assert key == "%s,%s" % (str(g["pcg_seed123_state"]), str(g["pcg_seed123_inc"]))
    assert noise.dtype == np.float64 and np.array_equal(noise, g["pcg_seed123_normals"])
    assert np.array_equal(src.decode(key), noise)
    src = RNGNoiseSource(6092, int(g["seed"]))
    keys = [src.sample()[0] for _ in range(int(g["N"]))]
    assert keys == [str(k) for k in g["rng_s0_keys"]]
    s = SimpleNoiseSource(5, 77)
    k, v = s.sample()
    assert k is v and s.decode(k) is k and np.array_equal(v, np.random.RandomState(77).randn(5))


class _StubPolicy(object):
    num_params = 6092
    input_shape = 17


def test_worker_draws_match_reference_worker(noise_json):
    g = noise_json["worker"]
    t = SharedNoiseTable(g["size"], 6092, g["seed"], upload=False)
    w = Worker(_StubPolicy(), None, t, None, sigma=0.02, eval_prob=g["eval_prob"], random_seed=g["seed"])
    flags, idx = w.draw_batch(g["batch"])
    assert flags.tolist() == g["flags"]
    assert [str(i) for i in idx[~flags]] == g["keys"]


def test_policy_init_bit_identical(golden_dir):
    g = np.load(os.path.join(golden_dir, "mujoco_c2.npz"))
    torch.manual_seed(124)
    th, _ = P.initial_parameters("mujoco", 17, 6, 124)
    assert np.array_equal(th, g["theta"])
    g = np.load(os.path.join(golden_dir, "discrete_c1.npz"))
    torch.manual_seed(124)
    th, buf = P.initial_parameters("discrete", 2, 9, 124)
    assert np.array_equal(th, g["theta"])
    assert buf.shape == (263,)


def test_layouts_and_state_dict_roundtrip(golden_dir):
    assert P.build_layout("mujoco", 17, 6).num_params == 6092
    assert P.build_layout("mujoco", 376, 17, 256, 256).num_params == 171042
    assert P.build_layout("atari", 0, 6).num_params == 678294
    L = P.build_layout("impala", 0, 15)
    assert (L.num_params, L.num_buffers) == (1158709, 5367)
    off = {e["name"]: e["off"] for e in L.entries if e["param"]}
    # SURVEY.md App. B offsets
    assert off["model.0.fc.1.weight"] == 102438 and off["model.0.core.weight_ih_l0"] == 626982
    assert off["model.0.policy.1.bias"] == 1158694 and off["model.0.resnet2.0.0.weight"] == 56390
    g = np.load(os.path.join(golden_dir, "discrete_c1.npz"))
    Ld = P.build_layout("discrete", 2, 9)
    theta, buf = Ld.split_state(g["serialized"])
    assert np.array_equal(theta, g["theta"])
    assert np.array_equal(Ld.join_state(theta, buf), g["serialized"])
    with pytest.raises(ValueError):
        Ld.split_state(g["serialized"][:-1])


def test_dsgd_lr_scale_map():
    class Om(object):
        omega, min_omega, max_omega = 0.3, 0.0, 1.0
    opt = DSGD([torch.nn.Parameter(torch.zeros(6092))], lr=0.01)
    assert is_dsgd(opt) and abs(opt.coef - np.sqrt(6092)) < 1e-12
    opt.adjust_lr(Om())
    assert abs(opt.lr_scale - (0.23 + 0.3 * 0.77)) < 1e-12
    assert affine_transform(1.0, 0.0, 0.0, 0.23, 1.0) == 0.23
    assert not is_dsgd(torch.optim.SGD([torch.nn.Parameter(torch.zeros(3))], lr=0.1))


def test_fd_return_record_roundtrip():
    r = FDReturn()
    r.epoch, r.encoded_noise, r.reward = 3, "123", 1.5
    q = FDReturn()
    q.deserialize(r.serialize())
    assert (q.epoch, q.encoded_noise, q.reward, q.is_eval) == (3, "123", 1.5, False)


def test_return_batch_is_a_sequence_of_fdreturns_and_keeps_the_arrays():
    """ReturnBatch: lazily built FDReturn records (reference fields, learner/fd_return.py:5-23) over the
    SoA arrays; the SoA view disappears once a record object has been handed out."""
    from dfd_starter_b200.fd_return import ReturnBatch, FDReturn
    idx = np.array([0, 11, 22, 11, 22], dtype=np.int64)
    sign = np.array([0, 1, 1, -1, -1], dtype=np.int8)
    rb = ReturnBatch(7, idx, sign, np.arange(5) * 0.5, np.zeros(5), np.full(5, 3), sign == 0)
    assert len(rb) == 5 and rb.soa is not None
    e, i, s, r = rb.soa
    assert e.tolist() == [7] * 5 and i.tolist() == idx.tolist() and s.tolist() == sign.tolist()
    assert [rb.key(j) for j in range(5)] == ["0", "+11", "+22", "-11", "-22"]
    recs = list(rb)
    assert rb.soa is None                       # records may now be edited by the caller
    assert all(isinstance(x, FDReturn) for x in recs)
    assert recs[0].is_eval and recs[0].encoded_noise == "0" and recs[3].encoded_noise == "-11"
    assert recs[2].reward == 1.0 and recs[2].epoch == 7 and recs[2].timesteps == 3
    one_sided = ReturnBatch(0, idx[1:3], np.ones(2, np.int8), np.zeros(2), np.zeros(2), np.ones(2), np.zeros(2, bool))
    assert [one_sided.key(j) for j in range(2)] == ["11", "22"]


def test_welford_running_stat_follows_the_reference(golden_dir):
    """utils/math_helpers.py:7-105 run by the reference: per-member sequential updates, learner-wide merge, mean / std
    (constant feature -> std 1), (de)serialisation.  Bit-exact (same numpy fp32 operations)."""
    from dfd_starter_b200.obs_stats import WelfordRunningStat
    g = np.load(os.path.join(golden_dir, "obs_stats.npz"))
    obs, select = g["obs"], g["select"]
    M, E, K = obs.shape
    glob = WelfordRunningStat(K)
    for m in range(M):
        st = WelfordRunningStat(K)
        for e in range(E):
            if select[m, e]:
                st.increment(obs[m, e], 1)
        assert np.array_equal(np.asarray(st.serialize(), dtype=np.float64), g["rows"][m]), m
        glob.increment_from_obs_stats_update(st.serialize())
        assert np.array_equal(np.asarray(glob.serialize(), dtype=np.float64), g["merged"][m]), m
    assert np.array_equal(glob.mean, g["mean"]) and np.array_equal(glob.std, g["std"])
    assert np.array_equal(np.clip(np.subtract(obs, glob.mean) / glob.std, -10, 10), g["normed"])
    back = WelfordRunningStat(K)
    back.deserialize(glob.serialize())
    assert np.array_equal(np.asarray(back.mean, dtype=np.float64), g["back_mean"])
    assert np.array_equal(np.asarray(back.std, dtype=np.float64), g["back_std"])
    fresh = WelfordRunningStat(K)
    assert np.array_equal(fresh.mean, np.zeros(K)) and np.array_equal(fresh.std, np.ones(K))     # count < 2


# ---------------------------------------------------------------- round 2: compute_vbn, default init order, flat-gradient surface
class _CpuCtx(object):
    device = torch.device("cpu")


def _policy_shell(cls, kind, n_act, theta, buf, in_shape):
    """A policy object without a device context: the HOST logic of compute_vbn / set_grad_from_flat is plain torch over
    views of theta / the buffer vector and can be checked here; everything that launches a kernel still needs the GPU."""
    p = object.__new__(cls)
    p.layout = P.build_layout(kind, 0 if kind in ("atari", "impala") else in_shape, n_act)
    p.num_params = p.layout.num_params
    p.theta, p.buffers = torch.from_numpy(np.array(theta, np.float32)), torch.from_numpy(np.array(buf, np.float32))
    p._entry = {e["name"]: e for e in p.layout.entries}
    p._flat_param, p.ctx, p.input_shape = None, _CpuCtx(), in_shape
    return p


def _rel(a, b):
    return float(np.max(np.abs(a - b) / (np.abs(b) + 1e-3)))


def test_compute_vbn_atari_and_impala_follow_the_reference(golden_dir):
    """policy.py:31-34, impala.py:13-17: buffers after the reference's own compute_vbn (tests/golden/make_golden.py
    gen_vbn) vs the train-mode pass of policies.py on the same synthetic theta / buffers / inputs."""
    from oracle import dfd_oracle as O
    g = np.load(os.path.join(golden_dir, "vbn.npz"))
    L = O.atari_layout(6)
    a = _policy_shell(P.AtariPolicy, "atari", 6, O.synthetic_theta(L, 21), O.synthetic_buffers(L, 22), (4, 84, 84))
    x = torch.rand(6, 4, 84, 84, generator=torch.Generator().manual_seed(int(g["atari_obs_seed"])))
    a.compute_vbn(x)
    a.compute_vbn(x[:3].numpy())                       # arrays are accepted as well; num_batches_tracked counts calls
    assert _rel(a.buffers.numpy(), g["atari_buffers_after"]) < 1e-4
    nbt = [e for e in a.layout.entries if e["name"].endswith("num_batches_tracked")]
    assert all(a.buffers[e["off"]] == g["atari_buffers_after"][e["off"]] for e in nbt)

    L = O.impala_layout(15)
    p = _policy_shell(P.ImpalaPolicy, "impala", 15, O.synthetic_theta(L, 31), O.synthetic_buffers(L, 32), (3, 64, 64))
    p.reset()
    frames = torch.randint(0, 256, (4, 1, 1, 3, 64, 64), generator=torch.Generator().manual_seed(int(g["impala_frame_seed"]))).float()
    buf = [{"frame": frames[i], "reward": torch.tensor(g["impala_reward"][i]).view(1, 1),
            "done": torch.tensor(g["impala_done"][i]).view(1, 1)} for i in range(4)]
    p.compute_vbn(buf)
    assert _rel(p.buffers.numpy(), g["impala_buffers_after"]) < 1e-4
    # the pass also advances the carried LSTM state by the 4 stacked entries (impala.py:166-173,184)
    np.testing.assert_allclose(p.state[0].reshape(-1).numpy(), g["impala_state_h"], atol=2e-6)
    np.testing.assert_allclose(p.state[1].reshape(-1).numpy(), g["impala_state_c"], atol=2e-6)


def test_default_init_follows_construction_order(golden_dir):
    """ImpalaCNN constructs its modules stage by stage but registers them list by list; the initial theta
    must equal the reference constructor's under the same torch seed (sha256 of the whole vector)."""
    import hashlib
    g = np.load(os.path.join(golden_dir, "default_init.npz"))
    for kind, n_act in (("impala", 15), ("atari", 6)):
        torch.manual_seed(124)
        th, _ = P.initial_parameters(kind, 0, n_act, 124)
        assert np.array_equal(th[g[kind + "_probe_idx"]], g[kind + "_probe"]), kind
        assert hashlib.sha256(th.tobytes()).hexdigest() == str(g[kind + "_sha256"]), kind


def test_flat_gradient_surface_drives_a_torch_optimizer():
    """policy.py:63-84: set_grad_from_flat accumulates like p.backward(grad); parameters() is one flat nn.Parameter
    that shares theta's storage, so a stock optimizer's step IS the theta update."""
    th = np.linspace(-1, 1, 6092).astype(np.float32)
    p = _policy_shell(P.MujocoPolicy, "mujoco", 6, th, np.zeros(0), 17)
    assert not p.get_grad_as_flat().any()
    opt = torch.optim.SGD(p.parameters(), lr=0.5)
    g = np.random.RandomState(0).randn(6092)
    p.set_grad_from_flat(g)
    p.set_grad_from_flat(g)                          # accumulates
    np.testing.assert_allclose(p.get_grad_as_flat(), 2 * g.astype(np.float32), rtol=1e-6)
    opt.step()
    np.testing.assert_allclose(p.theta.numpy(), th - g.astype(np.float32), atol=1e-6)
    opt.zero_grad()
    assert not p.get_grad_as_flat().any()
    with pytest.raises(ValueError):
        p.set_grad_from_flat(g[:10])


def test_pair_order_validates_the_paired_layout():
    """paired=True must not trust the [plus | minus] layout blindly."""
    from dfd_starter_b200.finite_differences import FiniteDifferences
    from dfd_starter_b200._lib import DfdError
    po = FiniteDifferences._pair_order
    idx = np.array([7, 3, 9, 7, 3, 9])
    assert po(idx, np.array([1, 1, 1, -1, -1, -1], np.int8)) is None                 # canonical: untouched
    # interleaved RPC chunks + an eval member (sign 0): regrouped, plus members in arrival order
    sel = po(np.array([0, 7, 7, 3, 3]), np.array([0, 1, -1, -1, 1], np.int8))
    assert sel.tolist() == [1, 4, 2, 3]
    # a row drawn twice pairs up with itself in any consistent way
    sel = po(np.array([5, 5, 5, 5]), np.array([1, -1, -1, 1], np.int8))
    assert sorted(sel[:2].tolist()) == [0, 3] and sorted(sel[2:].tolist()) == [1, 2]
    for bad_idx, bad_sign in (([5, 9, 5, 11], [1, 1, -1, -1]), ([5, 5, 5], [1, -1, 1]), ([5, 6], [1, 1])):
        with pytest.raises(DfdError):
            po(np.array(bad_idx), np.array(bad_sign, np.int8))


def test_exchange_mode_is_validated_without_a_gpu():
    """The learner's exchange mode is fixed per learner (every rank must take the same exchange path every step): only
    'fd_return' and 'general' exist, and the check runs before any device work (the constructor needs a GPU afterwards, so
    the accepted values are read from the source)."""
    import inspect
    import dfd_starter_b200.finite_differences as F
    src = inspect.getsource(F.FiniteDifferences.__init__)
    assert 'exchange_mode not in ("fd_return", "general")' in src
    assert src.index("exchange_mode not in") < src.index("get_context(")


def test_lazy_keys_behave_like_the_key_strings():
    """Batches drawn on the device carry their RNGNoiseSource keys as numbers (noise_sources.LazyKeys): the strings the
    reference puts in FDReturn.encoded_noise (utils/noise_sources.py:11: "state,inc"; worker.py:34: "0" for eval members)
    appear on demand, through ReturnBatch.key / iteration / select / concat alike."""
    import numpy as np
    from dfd_starter_b200.noise_sources import LazyKeys
    from dfd_starter_b200.fd_return import ReturnBatch
    rng = np.random.default_rng(3)
    states = [int.from_bytes(rng.bytes(16), "little") for _ in range(5)]
    inc = int.from_bytes(rng.bytes(16), "little") | 1
    m64 = (1 << 64) - 1
    streams = np.array([[s & m64, s >> 64, inc & m64, inc >> 64] for s in states], dtype=np.uint64)
    is_eval = np.array([False, True, False, False, True])
    keys = LazyKeys(streams, ~is_eval)
    want = ["0" if e else "{},{}".format(s, inc) for s, e in zip(states, is_eval)]
    assert list(keys) == want and len(keys) == 5 and keys[2] == want[2] and keys[1:3] == want[1:3]
    rb = ReturnBatch(np.full(5, 7), np.arange(5), np.ones(5, np.int8), np.arange(5.0), np.zeros(5), np.ones(5), is_eval, keys=keys)
    assert [rb.key(j) for j in range(5)] == want
    ne = rb.non_eval()
    assert isinstance(ne.keys, LazyKeys) and [r.encoded_noise for r in ne] == [w for w, e in zip(want, is_eval) if not e]
    both = ReturnBatch.concat([ne, rb])
    assert [both.key(j) for j in range(len(both))] == [w for w, e in zip(want, is_eval) if not e] + want
    assert rb.soa is None                      # keyed batches never take the table-index path
