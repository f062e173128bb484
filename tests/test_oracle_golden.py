"""Pins the CPU oracle (oracle/dfd_oracle.py) to outputs of the REFERENCE itself,
captured by tests/golden/make_golden.py (which imports /root/reference).  The
reference's own tests pin nothing on this path, so these fixtures
are the pin.  CPU only."""
import json
import os

import numpy as np
import pytest

from oracle import dfd_oracle as O


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


@pytest.fixture(scope="module")
def noise_json(golden_dir):
    with open(os.path.join(golden_dir, "noise.json")) as f:
        return json.load(f)


@pytest.mark.parametrize("key", ["1000000_6092_123", "200000_5197_124", "25000000_6092_124", "25000000_678294_124"])
def test_noise_table_bit_exact(noise_json, key):
    g = noise_json["tables"][key]
    t = O.NoiseTableOracle(g["size"], g["n_params"], g["seed"])
    assert t.sha256() == g["sha256"]
    assert [float(x) for x in t.table[:4]] == g["head"]
    assert [float(x) for x in t.table[-2:]] == g["tail"]
    keys = [t.sample()[0] for _ in range(8)]
    assert keys == g["keys"]
    i0 = int(keys[0])
    row = t.decode(keys[0]).astype(np.float64)
    assert np.array_equal(t.decode(keys[0]), t.table[i0:i0 + g["n_params"]])
    assert np.dot(row, row) == g["norm2_first"]


def test_survey_appendix_c_vectors(noise_json):
    """The SURVEY.md App. C numbers, captured independently during the survey."""
    g = noise_json["tables"]["25000000_6092_124"]
    assert g["sha256"] == "45e38c48c6df925f7b5c1f0564ceea9a02b35b0eb3f4edb366106ccfdac25994"
    assert [int(k) for k in g["keys"]] == [10700291, 7636593, 9022969, 13125983, 6182621, 10229889, 19192692, 18538305]
    assert abs(g["norm2_first"] - 6126.280724) < 1e-5
    g2 = noise_json["tables"]["1000000_6092_123"]
    assert g2["sha256"].startswith("55f7116e1d0eeebd")
    assert [int(k) for k in g2["keys"][:6]] == [850700, 451572, 337896, 813500, 804784, 863692]


def test_worker_flag_and_index_streams(noise_json):
    g = noise_json["worker"]
    noise = O.NoiseTableOracle(g["size"], 6092, g["seed"])
    flags, idx = O.draw_flags_and_indices(np.random.RandomState(g["seed"]), noise, g["eval_prob"], g["batch"])
    assert flags == g["flags"]
    assert [str(i) for i in idx] == g["keys"]


def test_perturbation_bit_exact_and_mujoco_forward(golden_dir):
    g = _load(golden_dir, "mujoco_c2.npz")
    t = O.NoiseTableOracle(int(g["table_size"]), 6092, int(g["table_seed"]))
    L = O.mujoco_layout(17, 6)
    assert L.num_params == 6092
    for m, (i, s) in enumerate(zip(g["idx"], g["sign"])):
        th = O.perturb(g["theta"], float(g["sigma"]), t.decode(str(i)), int(s))
        assert np.array_equal(th, g["theta_members"][m])          # bit-exact
        mean, std = O.mujoco_forward(L, th, g["obs"][m])
        np.testing.assert_allclose(np.concatenate([mean, std], -1), g["out"][m], rtol=0, atol=1e-6)


def test_humanoid_width_forward(golden_dir):
    g = _load(golden_dir, "mujoco_c3.npz")
    L = O.mujoco_layout(376, 17, 256, 256)
    assert L.num_params == 171042
    theta = O.synthetic_theta(L, int(g["theta_seed"]))
    t = O.NoiseTableOracle(int(g["table_size"]), L.num_params, int(g["table_seed"]))
    for m, (i, s) in enumerate(zip(g["idx"], g["sign"])):
        th = O.perturb(theta, float(g["sigma"]), t.decode(str(i)), int(s))
        mean, std = O.mujoco_forward(L, th, g["obs"][m])
        np.testing.assert_allclose(np.concatenate([mean, std], -1), g["out"][m], rtol=0, atol=2e-6)


def test_discrete_forward_and_state_dict_layout(golden_dir):
    g = _load(golden_dir, "discrete_c1.npz")
    L = O.discrete_layout(2, 9)
    assert (L.num_params, L.num_state) == (5197, 5460)
    theta, buffers = L.split_state(g["serialized"])
    assert np.array_equal(theta, g["theta"])
    assert np.array_equal(L.join_state(theta, buffers), g["serialized"])
    t = O.NoiseTableOracle(int(g["table_size"]), L.num_params, int(g["table_seed"]))
    for m, (i, s) in enumerate(zip(g["idx"], g["sign"])):
        th = O.perturb(theta, float(g["sigma"]), t.decode(str(i)), int(s))
        assert np.array_equal(th, g["theta_members"][m])
        np.testing.assert_allclose(O.discrete_forward(L, th, buffers, g["obs"][m]), g["out"][m], rtol=0, atol=1e-6)


def test_atari_forward(golden_dir):
    import torch
    g = _load(golden_dir, "atari_c4.npz")
    L = O.atari_layout(6)
    assert L.num_params == 678294
    theta = O.synthetic_theta(L, int(g["theta_seed"]))
    buffers = O.synthetic_buffers(L, int(g["buffer_seed"]))
    obs = torch.rand(3, 2, 4, 84, 84, generator=torch.Generator().manual_seed(int(g["obs_seed"]))).numpy()
    t = O.NoiseTableOracle(int(g["table_size"]), L.num_params, int(g["table_seed"]))
    for m, (i, s) in enumerate(zip(g["idx"], g["sign"])):
        th = O.perturb(theta, float(g["sigma"]), t.decode(str(i)), int(s))
        np.testing.assert_allclose(O.atari_forward(L, th, buffers, obs[m]), g["out"][m], rtol=0, atol=1e-6)


def test_impala_forward(golden_dir):
    import torch
    g = _load(golden_dir, "impala_c5.npz")
    L = O.impala_layout(15)
    assert L.num_params == 1158709
    theta = O.synthetic_theta(L, int(g["theta_seed"]))
    buffers = O.synthetic_buffers(L, int(g["buffer_seed"]))
    gen = torch.Generator().manual_seed(int(g["frame_seed"]))
    frames = torch.randint(0, 256, (2, 2, 3, 64, 64), generator=gen).float().numpy()
    h0 = (0.3 * torch.randn(2, 2, 256, generator=gen)).numpy()
    c0 = (0.3 * torch.randn(2, 2, 256, generator=gen)).numpy()
    t = O.NoiseTableOracle(int(g["table_size"]), L.num_params, int(g["table_seed"]))
    for m, (i, s) in enumerate(zip(g["idx"], g["sign"])):
        th = O.perturb(theta, float(g["sigma"]), t.decode(str(i)), int(s))
        p, h1, c1 = O.impala_forward(L, th, buffers, frames[m], g["reward"][m], g["done"][m], h0[m], c0[m])
        np.testing.assert_allclose(p, g["probs"][m], rtol=0, atol=1e-6)
        np.testing.assert_allclose(h1, g["h1"][m], rtol=0, atol=2e-6)
        np.testing.assert_allclose(c1, g["c1"][m], rtol=0, atol=2e-6)


def test_estimator_steps_match_reference(golden_dir):
    """FiniteDifferences.step + DSGD, 14 steps: fd_return mode, fd_state mode with
    delayed / too-old returns, std==0, baseline None, antithetic keys."""
    g = _load(golden_dir, "fd_steps.npz")
    noise = O.NoiseTableOracle(int(g["table_size"]), 6092, int(g["table_seed"]))
    fd = O.FiniteDifferencesOracle(g["theta0"], noise, float(g["sigma"]), float(g["lr"]),
                                   max_delayed_return=int(g["H"]), omega=float(g["omega"]))
    for s in range(int(g["n_steps"])):
        batch = [O.Ret(int(e), str(k), float(r)) for e, k, r in
                 zip(g["s%d_epochs" % s], g["s%d_keys" % s], g["s%d_rewards" % s])]
        b = float(g["s%d_baseline" % s])
        upd = fd.step(batch, None if np.isnan(b) else b)
        ref_g = g["s%d_grad" % s]
        assert np.max(np.abs(fd.gradient_memory - ref_g)) <= 1e-12 * np.max(np.abs(ref_g)), s
        assert np.array_equal(fd.theta, g["s%d_theta" % s]), s
        assert abs(upd - float(g["s%d_update" % s])) <= 1e-7 * abs(upd)
        assert fd.discarded_returns == int(g["s%d_discarded" % s])
        assert fd.epoch == int(g["s%d_epoch_after" % s])
    assert fd.step([], 0.0) == 0 and fd.epoch == int(g["n_steps"])


def test_rng_noise_source_stream_pins(golden_dir):
    """utils/noise_sources.py:4-20: key = PCG64 `state,inc` before the draw; SURVEY.md App. C words and normals."""
    g = _load(golden_dir, "fd_steps_hostnoise.npz")
    src = O.RNGNoiseSourceOracle(8, 123)
    key, noise = src.sample()
    assert key == "%s,%s" % (str(g["pcg_seed123_state"]), str(g["pcg_seed123_inc"]))
    assert key == "160078363690744033601390112987726904141,17686443629577124697969402389330893883"
    assert np.array_equal(noise, g["pcg_seed123_normals"])
    np.testing.assert_allclose(noise[:5], [-0.98912135, -0.36778665, 1.28792526, 0.19397442, 0.9202309], atol=5e-9)
    k2, n2 = src.sample()
    assert np.array_equal(src.decode(key), noise) and np.array_equal(src.decode(k2), n2)
    # decode rewinds the SAME generator (as the reference does): the next sample continues after the decoded vector
    assert src.sample()[0] != k2


@pytest.mark.parametrize("name", ["simple", "rng"])
def test_estimator_steps_with_host_noise_sources_match_reference(golden_dir, name):
    """Unmodified reference FiniteDifferences.step driven by SimpleNoiseSource / RNGNoiseSource (fp64 noise,
    delayed and too-old returns): the restatement follows it to fp64 rounding."""
    g = _load(golden_dir, "fd_steps_hostnoise.npz")
    P = g[name + "_theta0"].shape[0]
    src = (O.SimpleNoiseSourceOracle if name == "simple" else O.RNGNoiseSourceOracle)(P, int(g["seed"]))
    fd = O.FiniteDifferencesOracle(g[name + "_theta0"], src, float(g["sigma"]), float(g["lr"]),
                                   max_delayed_return=int(g["H"]), omega=float(g["omega"]))
    for s in range(int(g["n_steps"])):
        keys = [src.sample()[0] for _ in range(int(g["N"]))]
        if name == "rng":
            assert keys == [str(k) for k in g["rng_s%d_keys" % s]], s
        batch = [O.Ret(int(e), k, float(r)) for e, k, r in
                 zip(g["%s_s%d_epochs" % (name, s)], keys, g["%s_s%d_rewards" % (name, s)])]
        upd = fd.step(batch, 0.05 * s)
        ref_g = g["%s_s%d_grad" % (name, s)]
        assert np.max(np.abs(fd.gradient_memory - ref_g)) <= 1e-12 * np.max(np.abs(ref_g)), s
        assert np.array_equal(fd.theta, g["%s_s%d_theta" % (name, s)]), s
        assert abs(upd - float(g["%s_s%d_update" % (name, s)])) <= 1e-7 * abs(upd)
        assert fd.discarded_returns == int(g["%s_s%d_discarded" % (name, s)])


def test_closed_form_matches_stepwise(golden_dir):
    g = _load(golden_dir, "fd_steps.npz")
    noise = O.NoiseTableOracle(int(g["table_size"]), 6092, int(g["table_seed"]))
    keys = g["s0_keys"]
    idx = [int(k) for k in keys]
    cf = O.fd_gradient_closed_form(noise.table, idx, [1] * len(idx), g["s0_rewards"], float(g["sigma"]), 6092,
                                   baseline=float(g["s0_baseline"]))
    ref_g = g["s0_grad"]
    assert np.max(np.abs(cf - ref_g)) <= 2e-6 * np.max(np.abs(ref_g))


def test_dsgd_sanity():
    """P=6092, lr=0.01, omega=0 -> update 0.01*sqrt(6092)*0.23."""
    rng = np.random.RandomState(0)
    th = rng.randn(6092).astype(np.float32)
    g = rng.randn(6092).astype(np.float32)
    new = O.dsgd_step(th, g, 0.01, O.dsgd_lr_scale(0.0, 0.0, 1.0))
    assert abs(np.linalg.norm(th - new) - 0.1795179) < 2e-5


# ---------------------------------------------------------------- N3 strategy distances / novelty / history
STRATEGY_DISTANCES = ["l2_dist", "categorical_tvd", "gaussian_wasserstein_dist_from_strategies",
                      "categorical_bhattacharrya_dist", "gaussian_bhattacharrya_dist"]


@pytest.mark.parametrize("name", STRATEGY_DISTANCES)
def test_strategy_distance_functions_match_reference(golden_dir, name):
    g = _load(golden_dir, "strategy.npz")
    a, b = (g["cat_a"], g["cat_b"]) if name.startswith("categorical") else (g["gauss_a"], g["gauss_b"])
    if name == "gaussian_bhattacharrya_dist":        # the reference sums with einsum there: last-ulp fp32 differences
        np.testing.assert_allclose(O.strategy_distance(name, a[None], b), g["d_" + name], rtol=1e-6, atol=0)
    else:
        np.testing.assert_array_equal(O.strategy_distance(name, a, b), g["d_" + name])
    if name == "categorical_tvd":
        assert O.strategy_novelty(name, a, b) == float(g["novelty_tvd"])


@pytest.mark.parametrize("name", ["mujoco", "discrete"])
def test_strategy_history_follows_reference(golden_dir, name):
    """The reference StrategyHandler through 12 submissions with a 4-entry history: same replacement decisions, same
    next point to replace, same strategy tensor, same novelties."""
    g = _load(golden_dir, "strategy.npz")
    if name == "mujoco":
        lay = O.mujoco_layout(17, 6)
        ev = lambda flat, zeta: np.concatenate(O.mujoco_forward(lay, flat, zeta), -1)          # noqa: E731
        dist = "gaussian_wasserstein_dist_from_strategies"
    else:
        lay = O.discrete_layout(2, 9)
        _, buffers = lay.split_state(g["discrete_serialized"])
        ev = lambda flat, zeta: O.discrete_forward(lay, flat, buffers, zeta)                  # noqa: E731
        dist = "categorical_tvd"
    h = O.StrategyHistoryOracle(ev, dist, max_history_size=4)
    zeta = g[name + "_zeta"]
    for t in range(int(g[name + "_n_events"])):
        res = h.add_policy(g["%s_e%d_flat" % (name, t)])
        assert (-2 if res is None else res) == int(g["%s_e%d_submit" % (name, t)]), t
        if t in (1, 3, 6, 9):
            h.set_zeta(zeta)
        want = g["%s_e%d_tensor" % (name, t)]
        assert np.asarray(h.strategy_tensor).shape == want.shape, t
        if want.size:
            np.testing.assert_allclose(h.strategy_tensor, want, rtol=0, atol=2e-6)
        assert h.worst_point_idx == int(g["%s_e%d_worst" % (name, t)]), t
        nov = h.compute_novelty(g["%s_e%d_probe" % (name, t)])
        assert abs(nov - float(g["%s_e%d_novelty" % (name, t)])) <= 1e-5 * max(1.0, abs(nov)), t
