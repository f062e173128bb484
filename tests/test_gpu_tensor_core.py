"""tcgen05 / TMEM path of the per-member MLP forward (precision=1: tf32 operands, fp32 accumulate)
against the exact fp32 path and the CPU oracle.  Stated tolerance (outputs are post-tanh, |y| <= 1):
max-abs <= 2e-3 and mean-abs <= 3e-4 with unit-scale synthetic weights (tf32 keeps a 10-bit
mantissa: 2^-11 relative rounding of every operand through three layers; measured on B200:
max 1.0e-3, mean 1.1e-4), and the mean error must stay below 2 % of the member-to-member
signal the estimator differences."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import dfd_oracle as O  # noqa: E402


@pytest.fixture(scope="module")
def D():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import __graft_entry__ as G
    G.build()
    import dfd_starter_b200 as D
    return D


def _check(D, n_in, h, n_act, E, M, seed, atol, precision=1):
    L = O.mujoco_layout(n_in, n_act, h, h)
    table = D.SharedNoiseTable(1_000_000, L.num_params, 123, device=0)
    pol = D.MujocoPolicy(n_in, n_act, seed=seed, h1=h, h2=h, device=0, precision=precision).bind_table(table)
    exact = D.MujocoPolicy(n_in, n_act, seed=seed, h1=h, h2=h, device=0, precision=0).bind_table(table)
    theta = O.synthetic_theta(L, seed)
    pol.set_trainable_flat(theta)
    exact.set_trainable_flat(theta)
    rng = np.random.RandomState(seed + E)
    idx = rng.randint(0, 1_000_000 - L.num_params, size=M).astype(np.int64)
    sign = rng.choice([-1, 0, 1], size=M).astype(np.int8)
    obs = rng.randn(M, E, n_in).astype(np.float32)
    args = (torch.from_numpy(idx).cuda(), torch.from_numpy(sign).cuda(), torch.from_numpy(obs).cuda(), 0.02)
    out = pol.forward_members(*args).cpu().numpy()
    ref = exact.forward_members(*args).cpu().numpy()
    for m in range(min(M, 3)):      # and the CPU oracle for a few members
        th = theta if sign[m] == 0 else O.perturb(theta, 0.02, table._table[idx[m]:idx[m] + L.num_params], int(sign[m]))
        mean, std = O.mujoco_forward(L, th, obs[m])
        np.testing.assert_allclose(ref[m], np.concatenate([mean, std], -1), rtol=0, atol=2e-5)
    err = np.abs(out - ref)
    assert err.max() <= atol and err.mean() <= 3e-4, (err.max(), err.mean())
    # the finite-difference signal (difference between members) must survive the rounding
    sig = np.abs(ref - ref.mean(0, keepdims=True)).mean()
    assert err.mean() < 0.02 * sig + 1e-6, (err.mean(), sig)


@pytest.mark.parametrize("E", [128, 1, 37, 200])
def test_halfcheetah_shape_tf32(D, E):
    _check(D, 17, 64, 6, E, M=24, seed=3, atol=2e-3)


@pytest.mark.parametrize("E", [128, 1, 37, 200, 300])
def test_halfcheetah_shape_tf32_tanh_approx(D, E):
    """precision=2: tf32 operands + tanh.approx.f32 (2^-11 relative per activation, activations truncated - not
    rounded - to tf32 by the tensor core); stated tolerance max-abs 4e-3, mean-abs 3e-4 (checked inside _check)."""
    _check(D, 17, 64, 6, E, M=300, seed=4, atol=4e-3, precision=2)


def test_halfcheetah_shape_tf32_many_members(D):
    """more members than persistent CTAs x pipeline stages: every ring slot is reused several times"""
    _check(D, 17, 64, 6, 128, M=1500, seed=6, atol=2e-3, precision=1)


@pytest.mark.parametrize("shape", [(8, 3), (31, 16), (12, 5)])
def test_small_net_other_shapes_tf32(D, shape):
    _check(D, shape[0], 64, shape[1], 128, M=200, seed=8, atol=2e-3, precision=1)


@pytest.mark.parametrize("E", [128, 16])
def test_humanoid_shape_tf32(D, E):
    _check(D, 376, 256, 17, E, M=6, seed=5, atol=2e-3)
