"""GPU parity tests (-m gpu) of the "direct-from-table" tensor paths: a layer that is linear in its weights is evaluated as
x.theta^T + s*(x.(sigma*eps)^T) with both weight operands arriving by TMA (fp16 copy of theta; sigma-scaled fp16 mirror of
the noise table) - csrc/mlp_forward_direct.cu, csrc/cnn_forward_tc.cu.  The checker is the CPU oracle / the exact fp32 path;
the stated tolerance of these paths is 2e-3 on outputs that are probabilities or tanh-bounded (fp16 operands carry tf32's
10-bit mantissa; fp32 accumulate)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import dfd_oracle as O  # noqa: E402  (checker only)


@pytest.fixture(scope="module")
def D():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import __graft_entry__ as G
    G.build()
    import dfd_starter_b200 as D
    return D


def _atari(D, precision, table, theta, buf):
    pol = D.AtariPolicy((84, 84), 6, seed=124, device=0, precision=precision).bind_table(table)
    pol.set_trainable_flat(theta)
    pol.set_buffers(buf)
    return pol


@pytest.mark.parametrize("M,E,layout", [(4, 1, "pairs"), (4, 2, "pairs"), (6, 8, "pairs"), (4, 2, "unrelated"), (3, 1, "odd"),
                                        (3, 5, "odd"), (2, 16, "odd16"), (40, 1, "pairs")])
def test_atari_tensor_path_vs_oracle(D, M, E, layout):
    """policies/atari.py:35-51 through atari_forward_tc_kernel: pair CTAs with a shared table row (all eight table
    alignments occur), pairs of unrelated members (eval member with sign 0 among them), one CTA per member for odd member
    counts, up to 16 (member, observation) columns; against the CPU oracle per member and the exact fp32 kernel."""
    L = O.atari_layout(6)
    P = L.num_params
    table = D.SharedNoiseTable(2_000_000, P, 123, device=0)
    theta, buf = O.synthetic_theta(L, 51), O.synthetic_buffers(L, 52)
    tc, exact = _atari(D, 1, table, theta, buf), _atari(D, 0, table, theta, buf)
    rng = np.random.RandomState(100 * M + E)
    if layout == "pairs":
        half = (rng.randint(0, (2_000_000 - P) // 8, size=M // 2) * 8 + np.arange(M // 2) % 8).astype(np.int64)
        idx = np.concatenate([half, half])
        sign = np.concatenate([np.ones(M // 2), -np.ones(M // 2)]).astype(np.int8)
    elif layout == "unrelated":
        idx = rng.randint(0, 2_000_000 - P, size=M).astype(np.int64)
        sign = np.array([1, 0, -1, 1], dtype=np.int8)
    else:
        idx = rng.randint(0, 2_000_000 - P, size=M).astype(np.int64)
        sign = np.array([1, -1, 0], dtype=np.int8)[:M]
    obs = rng.rand(M, E, 4, 84, 84).astype(np.float32)
    args = (torch.from_numpy(idx).cuda(), torch.from_numpy(sign).cuda(), torch.from_numpy(obs).cuda(), 0.02)
    out = tc.forward_members(*args).cpu().numpy()
    ref = exact.forward_members(*args).cpu().numpy()
    assert np.isfinite(out).all()
    np.testing.assert_allclose(out.sum(-1), 1.0, atol=1e-5)
    err = np.abs(out - ref)
    assert err.max() <= 2e-3, (err.max(), np.unravel_index(np.argmax(err), err.shape))
    assert err.max() > 0.0                         # the tensor path did run (the exact kernel would match itself bit for bit)
    for m in sorted(set([0, M // 2, M - 1])):
        th = theta if sign[m] == 0 else O.perturb(theta, 0.02, table._table[idx[m]:idx[m] + P], int(sign[m]))
        np.testing.assert_allclose(out[m], O.atari_forward(L, th, buf, obs[m]), rtol=0, atol=2e-3)
    out2 = tc.forward_members(*args).cpu().numpy()
    assert np.array_equal(out, out2)               # deterministic


def test_atari_tensor_path_golden(D, golden_dir):
    """The reference's own outputs (tests/golden/atari_c4.npz, produced by the unmodified AtariPolicy) at the tensor path's
    stated tolerance."""
    import os
    g = np.load(os.path.join(golden_dir, "atari_c4.npz"))
    L = O.atari_layout(6)
    table = D.SharedNoiseTable(int(g["table_size"]), L.num_params, int(g["table_seed"]), device=0)
    pol = _atari(D, 1, table, O.synthetic_theta(L, int(g["theta_seed"])), O.synthetic_buffers(L, int(g["buffer_seed"])))
    obs = torch.rand(3, 2, 4, 84, 84, generator=torch.Generator().manual_seed(int(g["obs_seed"])))
    out = pol.forward_members(torch.from_numpy(g["idx"].astype(np.int64)).cuda(),
                              torch.from_numpy(g["sign"].astype(np.int8)).cuda(), obs.cuda(), float(g["sigma"])).cpu().numpy()
    np.testing.assert_allclose(out, g["out"], rtol=0, atol=2e-3)


def test_direct_and_streaming_wide_mlp_agree(D, monkeypatch):
    """C3 shape: the direct-from-table kernel and the streaming (weights built in shared memory) kernel are two tensor
    paths of the same forward; both within 2e-3 of the exact fp32 path, on pairs and on unrelated / unperturbed members."""
    n_in, h, n_act, M, E = 376, 256, 17, 302, 130
    L = O.mujoco_layout(n_in, n_act, h, h)
    P = L.num_params
    table = D.SharedNoiseTable(3_000_000, P, 123, device=0)
    theta = O.synthetic_theta(L, 7)
    pols = {}
    for name, prec in (("tensor", 1), ("exact", 0)):
        pols[name] = D.MujocoPolicy(n_in, n_act, seed=5, h1=h, h2=h, device=0, precision=prec).bind_table(table)
        pols[name].set_trainable_flat(theta)
    rng = np.random.RandomState(3)
    idx = rng.randint(0, 3_000_000 - P, size=M).astype(np.int64)
    idx[M // 2:] = idx[:M // 2]
    sign = np.concatenate([np.ones(M // 2), -np.ones(M // 2)]).astype(np.int8)
    sign[5] = 0
    idx[M - 1] = 17                                # an unrelated last member
    obs = rng.randn(M, E, n_in).astype(np.float32)
    args = (torch.from_numpy(idx).cuda(), torch.from_numpy(sign).cuda(), torch.from_numpy(obs).cuda(), 0.02)
    direct = pols["tensor"].forward_members(*args).cpu().numpy()
    monkeypatch.setenv("DFD_TC_NO_DIRECT", "1")
    stream = pols["tensor"].forward_members(*args).cpu().numpy()
    monkeypatch.delenv("DFD_TC_NO_DIRECT")
    ref = pols["exact"].forward_members(*args).cpu().numpy()
    assert np.abs(direct - ref).max() <= 2e-3 and np.abs(stream - ref).max() <= 2e-3
    assert not np.array_equal(direct, stream)      # two different kernels ran
    for m in (0, 5, M // 2, M - 1):
        th = theta if sign[m] == 0 else O.perturb(theta, 0.02, table._table[idx[m]:idx[m] + P], int(sign[m]))
        mean, std = O.mujoco_forward(L, th, obs[m])
        np.testing.assert_allclose(direct[m], np.concatenate([mean, std], -1), rtol=0, atol=2e-3)


@pytest.mark.parametrize("level", [2, 3])
@pytest.mark.parametrize("shared,M,E", [(True, 4, 2), (False, 4, 2), (False, 3, 1), (True, 40, 1)])
def test_impala_tcgen05_path_vs_oracle(D, shared, M, E, level):
    """policies/impala.py:136-186 at precision level 2 (mma.sync trunk + TMA-fed tcgen05 swap-AB dense tail with class-mapped
    LSTM rows, csrc/impala_tail.cuh inside impala_forward_kernel) and level 3 (impala_direct_kernel: tcgen05 trunk with the
    residual stream in TMEM + the same tail).  Pair CTAs with a shared table row, pairs of unrelated
    members (an unperturbed eval member among them), odd member counts (one CTA per member); non-zero incoming state, one
    finished environment, rewards outside [-1, 1].  Stated tolerance of the tensor paths: 2e-3 on the action
    probabilities, 1e-2 on the carried LSTM state (15 convolutions deep), against the CPU oracle for every member."""
    L = O.impala_layout(15)
    P = L.num_params
    table = D.SharedNoiseTable(2_500_000, P, 123, device=0)
    pol = D.ImpalaPolicy((3, 64, 64), 15, seed=124, device=0, precision=level).bind_table(table)
    theta, buf = O.synthetic_theta(L, 43), O.synthetic_buffers(L, 44)
    pol.set_trainable_flat(theta)
    pol.set_buffers(buf)
    rng = np.random.RandomState(7 * M + shared)
    half = (rng.randint(0, (2_500_000 - P) // 8, size=M // 2) * 8 + np.arange(M // 2) % 8).astype(np.int64)
    idx = np.concatenate([half, half]) if shared else rng.randint(0, 2_500_000 - P, size=M).astype(np.int64)
    sign = (np.concatenate([np.ones(M // 2), -np.ones(M // 2)]) if shared else np.array([1, 0, -1, 1][:M])).astype(np.int8)
    frames = rng.randint(0, 256, size=(M, E, 3, 64, 64)).astype(np.float32)
    reward = rng.uniform(-2, 2, size=(M, E)).astype(np.float32)
    done = np.zeros((M, E), bool)
    done[1, 0] = True
    h0 = (0.3 * rng.randn(M, E, 256)).astype(np.float32)
    c0 = (0.3 * rng.randn(M, E, 256)).astype(np.float32)
    args = (torch.from_numpy(idx).cuda(), torch.from_numpy(sign).cuda(), torch.from_numpy(frames).cuda(),
            torch.from_numpy(reward).cuda(), torch.from_numpy(done).cuda(), torch.from_numpy(h0).cuda(),
            torch.from_numpy(c0).cuda(), 0.02)
    probs, h1, c1 = [t.cpu().numpy() for t in pol.forward_members_impala(*args)]
    assert np.isfinite(probs).all() and np.isfinite(h1).all() and np.isfinite(c1).all()
    err = [0.0, 0.0, 0.0]
    check = range(M) if M <= 8 else (0, 1, 7, M // 2, M // 2 + 7, M - 1)
    for m in check:
        th = theta if sign[m] == 0 else O.perturb(theta, 0.02, table._table[idx[m]:idx[m] + P], int(sign[m]))
        rp, rh, rc = O.impala_forward(L, th, buf, frames[m], reward[m], done[m], h0[m], c0[m])
        err = [max(err[0], np.abs(probs[m] - rp).max()), max(err[1], np.abs(h1[m] - rh).max()), max(err[2], np.abs(c1[m] - rc).max())]
    print("impala level %d max-abs errors: probs %.2e h %.2e c %.2e" % ((level,) + tuple(err)))
    assert err[0] <= 2e-3 and err[1] <= 1e-2 and err[2] <= 1e-2, err
    again = [t.cpu().numpy() for t in pol.forward_members_impala(*args)]
    assert np.array_equal(probs, again[0]) and np.array_equal(h1, again[1])      # deterministic
    # the mma.sync path (level 1) is a different kernel with the same contract
    pol1 = D.ImpalaPolicy((3, 64, 64), 15, seed=124, device=0, precision=1).bind_table(table)
    pol1.set_trainable_flat(theta)
    pol1.set_buffers(buf)
    p1 = pol1.forward_members_impala(*args)[0].cpu().numpy()
    assert np.abs(p1 - probs).max() <= 4e-3 and not np.array_equal(p1, probs)


class _HostPolicy(object):
    def __init__(self, theta):
        self.theta_h = np.array(theta, dtype=np.float32)
        self.num_params = len(theta)

    def get_trainable_flat(self):
        return self.theta_h.copy()

    def set_trainable_flat(self, flat):
        self.theta_h = np.array(flat, dtype=np.float32)


class _Omega(object):
    omega, min_omega, max_omega = 0.3, 0.0, 1.0


@pytest.mark.parametrize("paired", [False, True])
def test_fd_state_dots_from_the_fp16_mirror(D, paired):
    """With the sigma-scaled fp16 mirror of the table registered, the eps . d pass of fd_state batches
    (learner/finite_differences.py:87-89,107) reads the mirror - half the bytes.  The dot enters ||lambda||^2 as a small
    correction, so the estimator stays within its 1e-5 gradient tolerance of the oracle: 10 steps, H = 5; one-sided members with
    delayed, too-old and current returns, and antithetic pairs from delayed epochs."""
    import contextlib
    import io
    rng = np.random.RandomState(9)
    P, N, H, sigma = 30498, 64, 5, 0.05
    table = D.SharedNoiseTable(2_000_000, P, 123, device=0)
    table.device_table.ensure_scaled16(sigma, P)
    theta = (rng.randn(P) * 0.1).astype(np.float32)
    opt = D.DSGD([torch.nn.Parameter(torch.zeros(P))], lr=0.05)
    fd = D.FiniteDifferences(_HostPolicy(theta), opt, _Omega(), table, noise_std=sigma, batch_size=N, max_delayed_return=H,
                             paired=paired)
    ofd = O.FiniteDifferencesOracle(theta, O.NoiseTableOracle(2_000_000, P, 123), sigma, 0.05, max_delayed_return=H, omega=0.3)
    for s in range(10):
        if paired:
            i = rng.randint(0, 2_000_000 - P, size=N // 2).astype(np.int64)
            idx, sign = np.concatenate([i, i]), np.concatenate([np.ones(N // 2), -np.ones(N // 2)]).astype(np.int8)
            keys = ["+%d" % v for v in i] + ["-%d" % v for v in i]
        else:
            idx, sign = rng.randint(0, 2_000_000 - P, size=N).astype(np.int64), np.ones(N, dtype=np.int8)
            keys = [str(int(v)) for v in idx]
        rewards = rng.randn(N) * 2.0
        if paired:          # the two members of a pair were evaluated against one FDState: same epoch, inside the window
            eh = fd.epoch - rng.randint(0, min(s, H) + 1, size=N // 2)
            epochs = np.concatenate([eh, eh])
        else:               # one-sided members: any epoch, some one epoch too old (discarded)
            epochs = fd.epoch - rng.randint(0, H + 2, size=N)
        with contextlib.redirect_stdout(io.StringIO()):
            upd = fd.step_arrays(epochs, idx, sign, rewards, 0.5)
            oupd = ofd.step([O.Ret(int(e), k, float(r)) for e, k, r in zip(epochs, keys, rewards)], 0.5)
        rel = float(np.max(np.abs(fd.gradient_memory - ofd.gradient_memory)) / np.max(np.abs(ofd.gradient_memory)))
        assert rel <= 1e-5, (s, rel)
        assert abs(upd - oupd) <= 1e-5 * oupd


def test_resident_theta_direct_kernel_for_64x64_nets():
    """mlp_forward_ws16_kernel (opt-in: DFD_WS16=1): C2 shape 17-64-64-6 with the 17-wide first-layer rows fetched as eight
    class boxes (rows 8q + c) straight from the table mirror, theta image resident, four items in flight.  Runs in a
    subprocess because the switch is read once per process; checked against the CPU oracle per member (pairs, unrelated and
    unperturbed members, ragged observation counts)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r'''
import sys, numpy as np, torch
sys.path.insert(0, %r)
import dfd_starter_b200 as D
from oracle import dfd_oracle as O
L = O.mujoco_layout(17, 6, 64, 64); P = L.num_params
table = D.SharedNoiseTable(1_000_000, P, 123, device=0)
theta = O.synthetic_theta(L, 3)
rng = np.random.RandomState(1)
for M, E, prec, tol in ((600, 128, 1, 2e-3), (301, 200, 2, 4e-3)):
    pol = D.MujocoPolicy(17, 6, seed=3, device=0, precision=prec).bind_table(table)
    pol.set_trainable_flat(theta)
    idx = rng.randint(0, 1_000_000 - P, size=M).astype(np.int64)
    if M %% 2 == 0:
        idx[M // 2:] = idx[:M // 2]
    sign = np.where(np.arange(M) < M // 2, 1, -1).astype(np.int8)
    sign[3] = 0
    obs = rng.randn(M, E, 17).astype(np.float32)
    out = pol.forward_members(torch.from_numpy(idx).cuda(), torch.from_numpy(sign).cuda(), torch.from_numpy(obs).cuda(), 0.02).cpu().numpy()
    assert pol.ctx._scaled16 is not None
    worst = 0.0
    for m in (0, 3, M // 2, M - 1, 17):
        th = theta if sign[m] == 0 else O.perturb(theta, 0.02, table._table[idx[m]:idx[m] + P], int(sign[m]))
        mean, std = O.mujoco_forward(L, th, obs[m])
        worst = max(worst, float(np.abs(out[m] - np.concatenate([mean, std], -1)).max()))
    assert worst <= tol, worst
    assert np.isfinite(out).all()
print("ws16 ok")
''' % root
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=280, cwd=root,
                         env=dict(os.environ, DFD_WS16="1"))
    assert out.returncode == 0 and "ws16 ok" in out.stdout, (out.stdout[-1500:], out.stderr[-1500:])
