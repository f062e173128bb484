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
    _smoke_tensor_paths(O, D)


def _smoke_tensor_paths(O, D):
    """Small launches of the tcgen05 / TMA kernels the bench spends its time in, each checked against the oracle, so the
    driver's kernel list of smoke() names them: resident-weight MLP forward (mlp_forward_ws_kernel), streaming MLP
    forward with TMA-fed weight tiles straight from the table mirror (mlp_forward_direct_kernel) and with weights built in
    shared memory (mlp_forward_stream_kernel), TMA-fed wide-row reduction (fd_reduce_tma_kernel, through one learner step
    of a Humanoid-sized policy), one-kernel learner step (fd_tail_kernel), Atari forward on tcgen05 (atari_forward_tc_kernel),
    IMPALA forward with tensor-core convolutions and the TMA-fed tcgen05 dense tail (impala_forward_kernel, level 2)."""
    sigma = 0.02
    rng = np.random.RandomState(11)

    def mlp(n_in, h, n_act, M, E, precision, atol, table):
        L = O.mujoco_layout(n_in, n_act, h, h)
        pol = D.MujocoPolicy(n_in, n_act, seed=3, h1=h, h2=h, device=0, precision=precision).bind_table(table)
        theta = O.synthetic_theta(L, 3)
        pol.set_trainable_flat(theta)
        half = rng.randint(0, table.size - L.num_params, size=M // 2).astype(np.int64)
        idx, sign = np.concatenate([half, half]), np.concatenate([np.ones(M // 2), -np.ones(M // 2)]).astype(np.int8)
        obs = rng.randn(M, E, n_in).astype(np.float32)
        out = pol.forward_members(torch.from_numpy(idx).cuda(), torch.from_numpy(sign).cuda(), torch.from_numpy(obs).cuda(),
                                  sigma).cpu().numpy()
        worst = 0.0
        for m in (0, M // 2, M - 1):
            th = O.perturb(theta, sigma, table._table[idx[m]:idx[m] + L.num_params], int(sign[m]))
            mean, std = O.mujoco_forward(L, th, obs[m])
            worst = max(worst, float(np.max(np.abs(out[m] - np.concatenate([mean, std], -1)))))
        assert worst <= atol, ("tensor-path forward mismatch", n_in, h, worst)
        return pol, theta, worst

    t_small = D.SharedNoiseTable(1_000_000, 6092, 123, device=0)
    _, _, e_ws = mlp(17, 64, 6, 16, 128, 2, 4e-3, t_small)                       # mlp_forward_ws_kernel
    t_wide = D.SharedNoiseTable(2_000_000, 171042, 123, device=0)
    pol, theta, e_dr = mlp(376, 256, 17, 300, 128, 1, 2e-3, t_wide)              # mlp_forward_direct_kernel, 2 items per CTA
    os.environ["DFD_TC_NO_DIRECT"] = "1"
    _, _, e_st = mlp(376, 256, 17, 300, 128, 1, 2e-3, t_wide)                    # mlp_forward_stream_kernel
    del os.environ["DFD_TC_NO_DIRECT"]

    # one learner step at Humanoid width: prepare + fd_reduce_tma_kernel + DSGD, against the fp64 closed form
    class Omega(object):
        omega, min_omega, max_omega = 0.0, 0.0, 1.0
    P, R = 171042, 48
    opt = D.DSGD([torch.nn.Parameter(torch.zeros(P))], lr=0.01)
    fd = D.FiniteDifferences(pol, opt, Omega(), t_wide, noise_std=sigma, batch_size=2 * R, max_delayed_return=2, paired=True)
    idx = t_wide.sample_indices(R)
    rew = rng.randn(2 * R)
    sign = np.concatenate([np.ones(R), -np.ones(R)]).astype(np.int8)
    fd.step_arrays(np.zeros(2 * R, np.int64), np.concatenate([idx, idx]), sign, rew, 0.0)
    ref = O.fd_gradient_closed_form(t_wide._table, np.concatenate([idx, idx]), sign, rew, sigma, P)
    e_red = float(np.max(np.abs(fd.gradient_memory - ref)) / np.max(np.abs(ref)))
    assert e_red <= 1e-5, ("wide reduction mismatch", e_red)

    # Atari on tcgen05: implicit-GEMM convolutions + TMA-fed swap-AB first Linear, one antithetic pair
    La = O.atari_layout(6)
    t_at = D.SharedNoiseTable(2_000_000, La.num_params, 123, device=0)
    apol = D.AtariPolicy((84, 84), 6, seed=124, device=0, precision=1).bind_table(t_at)
    tha, bufa = O.synthetic_theta(La, 51), O.synthetic_buffers(La, 52)
    apol.set_trainable_flat(tha)
    apol.set_buffers(bufa)
    ia = int(rng.randint(0, 2_000_000 - La.num_params))
    aobs = rng.rand(2, 1, 4, 84, 84).astype(np.float32)
    aout = apol.forward_members(torch.tensor([ia, ia]).cuda(), torch.tensor([1, -1], dtype=torch.int8).cuda(),
                                torch.from_numpy(aobs).cuda(), sigma).cpu().numpy()
    e_at = 0.0
    for m, sg in enumerate((1, -1)):
        thm = O.perturb(tha, sigma, t_at._table[ia:ia + La.num_params], sg)
        e_at = max(e_at, float(np.abs(aout[m] - O.atari_forward(La, thm, bufa, aobs[m])).max()))
    assert e_at <= 2e-3, ("atari tensor-path forward mismatch", e_at)

    # IMPALA, tcgen05 trunk (shifted-descriptor implicit GEMMs) + TMA-fed tcgen05 dense tail (level 3), one antithetic pair
    L = O.impala_layout(15)
    t_imp = D.SharedNoiseTable(2_000_000, L.num_params, 123, device=0)
    ipol = D.ImpalaPolicy((3, 64, 64), 15, seed=124, device=0, precision=3).bind_table(t_imp)
    th, buf = O.synthetic_theta(L, 43), O.synthetic_buffers(L, 44)
    ipol.set_trainable_flat(th)
    ipol.set_buffers(buf)
    i0 = int(rng.randint(0, 2_000_000 - L.num_params))
    idx, sign = np.array([i0, i0], np.int64), np.array([1, -1], np.int8)
    frames = rng.randint(0, 256, size=(2, 1, 3, 64, 64)).astype(np.float32)
    zero = np.zeros((2, 1, 256), np.float32)
    probs, h1, c1 = ipol.forward_members_impala(torch.from_numpy(idx).cuda(), torch.from_numpy(sign).cuda(),
                                                torch.from_numpy(frames).cuda(), torch.zeros(2, 1).cuda(),
                                                torch.zeros(2, 1, dtype=torch.bool).cuda(), torch.from_numpy(zero).cuda(),
                                                torch.from_numpy(zero).cuda(), sigma)
    e_imp = 0.0
    for m in range(2):
        thm = O.perturb(th, sigma, t_imp._table[i0:i0 + L.num_params], int(sign[m]))
        rp, _, _ = O.impala_forward(L, thm, buf, frames[m], np.zeros(1, np.float32), np.zeros(1, bool), zero[m], zero[m])
        e_imp = max(e_imp, float(np.abs(probs[m].cpu().numpy() - rp).max()))
    assert e_imp <= 2e-3, ("impala tensor-core forward mismatch", e_imp)
    print("smoke ok (tensor / TMA paths): ws forward %.1e, direct forward %.1e, stream forward %.1e, wide reduction rel %.1e, "
          "atari probs %.1e, impala probs %.1e, launches %d" % (e_ws, e_dr, e_st, e_red, e_at, e_imp, pol.ctx.launch_count()))
    _smoke_rng_rows(D)


def _smoke_rng_rows(D):
    """RNGNoiseSource rows drawn on the device (csrc/rng_normal.cu) against numpy's own generator: bit-identical."""
    import numpy as np
    import torch
    from dfd_starter_b200.device import get_context
    ctx = get_context(0)
    P, n = 6092, 6
    src, ref = D.RNGNoiseSource(P, 7), D.RNGNoiseSource(P, 7, device=False)
    out = torch.zeros(n * P, dtype=torch.float32, device=ctx.device)
    keys = src.sample_rows(ctx, n, out, P)
    got = out.cpu().numpy().reshape(n, P)
    for j in range(n):
        key, eps = ref.sample()
        assert keys[j] == key and np.array_equal(got[j].view(np.uint32), eps.astype(np.float32).view(np.uint32)), ("rng row", j)
    print("smoke ok (device RNGNoiseSource): %d rows x %d normals bit-identical to numpy, keys equal" % (n, P))
