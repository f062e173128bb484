"""smoke(): one small invocation of the hot path on cuda:0, checked against the CPU oracle."""
import os
import sys

import numpy as np
import torch


def smoke():
    root = os.path.dirname(os.path.abspath(__file__))
    if root not in sys.path:
        sys.path.insert(0, root)
    from oracle import dfd_oracle as O          # the checker (test infrastructure)
    import dfd_starter_b200 as D

    torch.manual_seed(124)
    table = D.SharedNoiseTable(1_000_000, 6092, 123, device=0)
    policy = D.MujocoPolicy(17, 6, seed=124, device=0).bind_table(table)
    oracle_table = O.NoiseTableOracle(1_000_000, 6092, 123)
    L = O.mujoco_layout(17, 6)
    theta0 = policy.get_trainable_flat().copy()
    sigma = 0.02

    # perturbed forward of 8 members x 4 observations
    idx = table.sample_indices(8)
    sign = np.array([1, -1, 1, 1, -1, 0, 1, -1], dtype=np.int8)
    obs = torch.randn(8, 4, 17, generator=torch.Generator().manual_seed(0))
    out = policy.forward_members(torch.from_numpy(idx).cuda(), torch.from_numpy(sign).cuda(), obs.cuda(), sigma).cpu().numpy()
    for m in range(8):
        th = O.perturb(theta0, sigma, oracle_table.table[idx[m]:idx[m] + 6092], 1) if sign[m] >= 0 else \
            O.perturb(theta0, sigma, oracle_table.table[idx[m]:idx[m] + 6092], -1)
        if sign[m] == 0:
            th = theta0
        mean, std = O.mujoco_forward(L, th, obs[m].numpy())
        ref = np.concatenate([mean, std], -1)
        assert np.max(np.abs(out[m] - ref)) < 1e-5, ("forward mismatch", m, np.max(np.abs(out[m] - ref)))

    # one learner step (fd_return mode) against the oracle estimator
    class Omega(object):
        omega, min_omega, max_omega = 0.3, 0.0, 1.0
    opt = D.DSGD([torch.nn.Parameter(torch.zeros(6092))], lr=0.01)
    learner = D.FiniteDifferences(policy, opt, Omega(), table, noise_std=sigma, batch_size=32, max_delayed_return=4)
    ofd = O.FiniteDifferencesOracle(theta0, oracle_table, sigma, 0.01, max_delayed_return=4, omega=0.3)
    rng = np.random.RandomState(0)
    batch = []
    for i in table.sample_indices(32):
        r = D.FDReturn()
        r.epoch, r.encoded_noise, r.reward = 0, str(int(i)), float(rng.randn() * 3 + 1)
        batch.append(r)
    upd = learner.step(batch, 0.1, 0.0, 0.0)
    oupd = ofd.step([O.Ret(b.epoch, b.encoded_noise, b.reward) for b in batch], 0.1)
    g, og = learner.gradient_memory, ofd.gradient_memory
    rel = np.max(np.abs(g - og)) / np.max(np.abs(og))
    assert rel < 1e-5, ("gradient mismatch", rel)
    assert abs(upd - oupd) < 1e-5 * oupd, ("update size", upd, oupd)
    assert np.max(np.abs(policy.get_trainable_flat() - ofd.theta)) < 1e-6
    print("smoke ok: forward max-abs < 1e-5, gradient rel-max %.2e, update %.6f (oracle %.6f), launches %d"
          % (rel, upd, oupd, policy.ctx.launch_count()))
