#!/usr/bin/env python
"""bench.py — the hot path of BASELINE.json's north_star on synthetic data.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C2|C3|C1] [--obs-per-member E]
    python bench.py --impl reference ...        # the CPU restatement of the reference on the host cores

One STEP = one learner epoch over one population batch on each GPU:
    perturbed forward of all members (theta +/- sigma*eps generated in-kernel from table offsets,
    E synthetic observations per member) -> synthetic return per member -> dfd_fd_prepare ->
    dfd_fd_reduce (the eps-weighted gradient reduction) -> [NCCL allreduce of P floats, N > 1] ->
    dfd_dsgd_step (theta update + theta-history / distance rows).
metric  = perturbed-policy env-steps/s = members * E * n_gpus / step time (whole job);
          FD-gradient estimates/s (= steps/s) is reported beside it.
value   : inputs resident in HBM, fresh noise indices and a different observation buffer every step
          (table replicas 400 MB and the rotating observation buffers exceed L2; stated in config).
e2e     : the same step through the reference-facing objects (Worker.evaluate -> FDReturn list ->
          FiniteDifferences.step) with HOST observations / indices / returns copied in and results
          copied out every step.
Weak scaling: every rank evaluates the workload's full per-GPU population (no data-path
collective in the forward; one parameter-sized allreduce in the estimator).
Beside the headline (default run, one GPU) the same JSON line carries: `workloads` (C3 / C4 / C5 at the same N),
`operating_points` (C2 at E = 1, E = 16, exact fp32) and `noise_sources.rng` - the reference drivers' DEFAULT noise
source (RNGNoiseSource: numpy PCG64 + ziggurat) drawn on the device bit-identically to numpy: rows/s at the C2 / C3 row
shapes next to numpy's rate on this host, and the reference-style worker -> learner step with the rows drawn on the
device vs on the host.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: kind, n_in, h1, h2, n_act, antithetic pairs per GPU, default E, description
    "C2": dict(kind="mujoco", n_in=17, h1=64, h2=64, n_act=6, pairs=1024, E=128,
               desc="C2 HalfCheetah-shaped MLP 17-64-64-6, 1024 antithetic pairs per GPU, fd_return"),
    "C3": dict(kind="mujoco", n_in=376, h1=256, h2=256, n_act=17, pairs=1024, E=128,
               desc="C3 Humanoid-shaped MLP 376-256-256-17, 8192 antithetic pairs over 8 GPUs (1024 per GPU), fd_return"),
    "C4": dict(kind="atari", n_in=4 * 84 * 84, h1=0, h2=0, n_act=6, pairs=512, E=1,
               desc="C4 Atari 2-conv CNN policy (84x84x4 frames), 512 antithetic pairs, fd_return"),
    "C5": dict(kind="impala", n_in=3 * 64 * 64, h1=0, h2=0, n_act=15, pairs=256, E=1,
               desc="C5 IMPALA-CNN + LSTM policy (64x64x3 frames), 2048 antithetic pairs over 8 GPUs (256 per GPU), "
                    "fd_state estimator (returns from the last 10 epochs)"),
    "C1": dict(kind="discrete", n_in=2, h1=64, h2=64, n_act=9, pairs=20, E=128,
               desc="C1 simple_trap-shaped discrete MLP 2-64-64-9, 20 antithetic pairs, fd_return"),
}
# which committed `ncu --set full` capture (profiles/rNN/<name>_raw_metrics.csv, newest round first) holds the dominant
# kernel of a (workload, kernel) pair; roofline.traffic is READ from it at run time (dram__bytes_read.sum + dram__bytes_write.sum)
NCU_CAPTURE = {("C2", "policy_forward"): "prof_fwd_c2", ("C3", "policy_forward"): "prof_fwd_c3",
               ("C4", "policy_forward"): "prof_fwd_c4", ("C5", "policy_forward"): "prof_fwd_c5",
               ("C2", "fd_reduce"): "prof_reduce_c2", ("C3", "fd_reduce"): "prof_reduce_c3",
               ("C5", "fd_reduce"): "prof_reduce_c5", ("cold", "fd_reduce"): "prof_reduce_cold"}
_UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def ncu_traffic(workload, kernel):
    """(bytes per launch, capture path) from the newest committed capture of this kernel, or (None, None)."""
    name = NCU_CAPTURE.get((workload, kernel))
    prof = os.path.join(ROOT, "profiles")
    if name is None or not os.path.isdir(prof):
        return None, None
    for rnd in sorted((d for d in os.listdir(prof) if os.path.isdir(os.path.join(prof, d))), reverse=True):
        path = os.path.join(prof, rnd, name + "_raw_metrics.csv")
        if not os.path.exists(path):
            continue
        tot, seen = 0.0, 0
        with open(path) as f:
            for line in f:
                c = line.rstrip("\n").split(",")
                if len(c) >= 3 and c[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum") and c[1] in _UNIT:
                    tot += float(c[2]) * _UNIT[c[1]]
                    seen += 1
        if seen == 2:
            return tot, "profiles/%s/%s_raw_metrics.csv" % (rnd, name)
    return None, None


TABLE_SIZE = 25_000_000
TABLE_SEED = 124
SIGMA = 0.02
LR = 0.01
H = 10


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS))
    ap.add_argument("--obs-per-member", type=int, default=0)
    ap.add_argument("--pairs", type=int, default=0, help="override antithetic pairs per GPU")
    ap.add_argument("--table-size", type=int, default=TABLE_SIZE)
    ap.add_argument("--graph", default="auto", choices=["auto", "on", "off"])
    ap.add_argument("--precision", default="auto", choices=["auto", "fp32", "tf32", "tf32a"],
                    help="forward arithmetic: fp32 CUDA cores (exact path, atol 1e-5), tf32 = tcgen05 tensor cores with "
                         "tf32 operands / fp32 accumulate / ~1e-6 tanh (max-abs 2e-3 vs fp32), tf32a = the same with the "
                         "single-instruction tanh.approx.f32 (2^-11 relative; max-abs 4e-3 vs fp32); "
                         "auto = tf32a for MuJoCo MLPs with >= 32 observations per member, tensor-core convolutions for IMPALA "
                         "(fp16 operands / fp32 accumulate, max-abs 2e-3 on the action probabilities vs fp32), fp32 otherwise")
    ap.add_argument("--profile-mode", action="store_true",
                    help="for runs under ncu: timed steps only (no clock-load loop, per-kernel timing, e2e or CPU baseline)")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N > 1: gradient exchange as one NVLink peer-memory kernel (default) or NCCL collectives")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-workloads", action="store_true",
                    help="default C2 run only: skip the C3 / C4 / C5 workloads and the E = 1 / 16 / fp32 operating points")
    ap.add_argument("--cpu-sample-members", type=int, default=0)
    ap.add_argument("--reference-port", action="store_true",
                    help="--impl reference: time the oracle port even when the reference itself is importable")
    return ap.parse_args()


def workload(args):
    w = dict(WORKLOADS[args.workload])
    if args.obs_per_member:
        w["E"] = args.obs_per_member
    if args.pairs:
        w["pairs"] = args.pairs
    w["name"] = args.workload
    w["members"] = 2 * w["pairs"]
    w["out_width"] = 2 * w["n_act"] if w["kind"] == "mujoco" else w["n_act"]
    return w


def forward_precision(kind, precision, E):
    """--precision -> (tensor path on?, dfd_policy_desc.precision level).  MuJoCo MLPs: tcgen05 tf32 (level 1: accurate
    tanh, level 2: tanh.approx) when asked for, or by default from 32 observations per member; IMPALA: tensor-core
    convolutions + dense tail on tcgen05 (level 3) unless fp32 is asked for; Atari: tcgen05 convolutions and first Linear with TMA-fed weight tiles
    (level 1) unless fp32 is asked for; Discrete: the exact fp32 kernels only."""
    if kind == "mujoco":
        on = precision in ("tf32", "tf32a") or (precision == "auto" and E >= 32)
        return on, (1 if precision == "tf32" else 2)
    if kind == "impala":
        return precision != "fp32", 3       # level 3: tcgen05 trunk (a filter tap = a shifted UMMA descriptor) + TMA-fed tcgen05 dense tail
    if kind == "atari":
        return precision != "fp32", 1
    return False, 0


def layer_flops_per_obs(w):
    if w["kind"] == "atari":
        return 5934080
    if w["kind"] == "impala":
        return 62268928
    return 2 * (w["n_in"] * w["h1"] + w["h1"] * w["h2"] + w["h2"] * w["out_width"])


# ----------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port timed on the host cores
# ----------------------------------------------------------------------------------------------
_G = {}


def _ref_member(m):
    """One member the way the reference worker evaluates it (worker/worker.py:26-32 +
    policy.forward), batched over the member's E observations (favourable to the CPU: the
    reference makes E separate batch-1 calls)."""
    O, L, theta, table, idx, sign, obs, kind, bufs = (_G[k] for k in ("O", "L", "theta", "table", "idx", "sign", "obs", "kind", "bufs"))
    th = O.perturb(theta, SIGMA, table[idx[m]:idx[m] + theta.shape[0]], int(sign[m]))
    if kind == "mujoco":
        mean, std = O.mujoco_forward(L, th, obs)
        out = np.concatenate([mean, std], -1)
    elif kind == "discrete":
        out = O.discrete_forward(L, th, bufs, obs)
    elif kind == "atari":
        out = O.atari_forward(L, th, bufs, obs.reshape(-1, 4, 84, 84))
    else:
        n = obs.shape[0]
        out, _, _ = O.impala_forward(L, th, bufs, obs.reshape(-1, 3, 64, 64) * 255.0, np.zeros(n, np.float32),
                                     np.zeros(n, bool), np.zeros((n, 256), np.float32), np.zeros((n, 256), np.float32))
    return float(-np.mean((out - _G["target"]) ** 2))


def run_reference(args, w, as_baseline=False):
    import torch
    from oracle import dfd_oracle as O
    torch.set_num_threads(1)          # the reference clients run single-threaded (run_client.py:15)
    cores = len(os.sched_getaffinity(0))
    P_layout = {"mujoco": lambda: O.mujoco_layout(w["n_in"], w["n_act"], w["h1"], w["h2"]),
                "discrete": lambda: O.discrete_layout(w["n_in"], w["n_act"], w["h1"], w["h2"]),
                "atari": lambda: O.atari_layout(w["n_act"]), "impala": lambda: O.impala_layout(w["n_act"])}[w["kind"]]()
    P = P_layout.num_params
    noise = O.NoiseTableOracle(args.table_size, P, TABLE_SEED)
    theta = O.synthetic_theta(P_layout, 1)
    bufs = O.synthetic_buffers(P_layout, 2) if P_layout.num_buffer else None
    M, E = w["members"], w["E"]
    sample = args.cpu_sample_members or min(M, max(cores * 32, 256))
    rng = np.random.RandomState(0)
    obs = (rng.rand(E, w["n_in"]) if w["kind"] in ("atari", "impala") else rng.randn(E, w["n_in"])).astype(np.float32)
    _G.update(O=O, L=P_layout, theta=theta, table=noise.table, kind=w["kind"], bufs=bufs, obs=obs,
              target=np.tanh(rng.randn(w["out_width"])).astype(np.float32) * 0.5)
    import multiprocessing as mp
    steps, warm = (args.steps, args.warmup) if not as_baseline else (2, 1)
    steps = max(1, min(steps, 5))     # each step is already seconds of CPU work (the count actually run is printed)
    warm = max(0, min(warm, 1))
    times = []
    for it in range(warm + steps):
        pairs_idx = np.array([int(noise.sample()[0]) for _ in range(w["pairs"])], dtype=np.int64)
        idx = np.concatenate([pairs_idx, pairs_idx])
        sign = np.concatenate([np.ones(w["pairs"]), -np.ones(w["pairs"])]).astype(np.int8)
        _G.update(idx=idx, sign=sign, theta=theta)
        members = list(range(0, M, max(1, M // sample)))[:sample]
        # N single-threaded client processes, like the reference's run_client.py fleet; forked after
        # _G holds this step's theta / indices so the children see them
        pool = mp.get_context("fork").Pool(cores) if cores > 1 else None
        t0 = time.perf_counter()
        if pool is not None:
            rewards_s = pool.map(_ref_member, members, chunksize=max(1, len(members) // (cores * 2)))
            pool.close()
        else:
            rewards_s = [_ref_member(m) for m in members]
        t_fwd = (time.perf_counter() - t0) * (M / len(members))
        rewards = rng.randn(M)
        rewards[:len(rewards_s)] = rewards_s
        # the estimator restatement on a bounded number of returns (its cost is linear in returns x P)
        n_fd = min(M, max(64, int(2.0e8 // P) // 2 * 2))
        sel = np.concatenate([np.arange(n_fd // 2), w["pairs"] + np.arange(n_fd // 2)])
        fd = O.FiniteDifferencesOracle(theta, noise, SIGMA, LR, max_delayed_return=H, omega=0.0)
        batch = [O.Ret(0, ("+%d" if s > 0 else "-%d") % i, float(r)) for i, s, r in zip(idx[sel], sign[sel], rewards[sel])]
        t1 = time.perf_counter()
        fd.step(batch, 0.0)
        t_fd = (time.perf_counter() - t1) * (M / n_fd)
        theta = fd.theta
        if it >= warm:
            times.append((t_fwd, t_fd))
    t_fwd = float(np.mean([t[0] for t in times]))
    t_fd = float(np.mean([t[1] for t in times]))
    step_s = t_fwd + t_fd
    value = M * E / step_s
    sample_txt = ("forward: %d of %d members x %d obs per step on %d processes x 1 torch thread, scaled x%.1f; "
                  "estimator: %d of %d returns (FiniteDifferences.step restatement, numpy BLAS threads), scaled linearly"
                  % (sample, M, E, cores, M / sample, n_fd, M))
    return dict(value=value, unit="env-steps/s", cores=cores, kind="port", sample=sample_txt,
                ms_per_step=step_s * 1e3, fd_estimates_per_s=1.0 / t_fd, forward_s=t_fwd, estimator_s=t_fd, steps_run=steps)


# ----------------------------------------------------------------------------------------------
# reference arm proper: the UNMODIFIED reference, imported from baseline/_ref (or $DFD_REFERENCE)
# ----------------------------------------------------------------------------------------------
def reference_root():
    """Where an importable copy of the unmodified reference lives: $DFD_REFERENCE, else baseline/_ref (installed once
    with pip --target from the read-only checkout, git-ignored, travels to the GPU box).  None -> the oracle port."""
    for cand in (os.environ.get("DFD_REFERENCE"), os.path.join(ROOT, "baseline", "_ref")):
        if cand and os.path.exists(os.path.join(cand, "learner", "finite_differences.py")):
            return cand
    return None


_R = {}


def _real_member_chunk(task):
    """worker/worker.py:26-32 for a chunk of members, in a forked single-threaded client process: new_flat = flat +
    sigma * eps -> set_trainable_flat -> policy.forward(obs) -> set_trainable_flat(flat)."""
    import torch
    members, flat = task
    pol, table, idx, sign, obs, kind, target = (_R[k] for k in ("policy", "table", "idx", "sign", "obs", "kind", "target"))
    out_r = []
    with torch.no_grad():
        for m in members:
            eps = table.decode(str(int(idx[m])))
            new_flat = flat + SIGMA * (eps if sign[m] > 0 else -eps)
            pol.set_trainable_flat(new_flat)
            if kind == "mujoco":
                mean, std = pol.forward(obs)
                out = np.concatenate([mean.numpy(), std.numpy()], -1)
            elif kind == "impala":
                pol.reset()
                outs = []
                for e in range(obs.shape[0]):           # E independent single-step environments (impala.py:136-186)
                    pol.reset()
                    outs.append(pol.forward({"frame": obs[e:e + 1].view(1, 1, 3, 64, 64), "reward": torch.zeros(1, 1),
                                             "done": torch.zeros(1, 1, dtype=torch.bool)}).reshape(-1).numpy())
                out = np.stack(outs)
            else:
                out = pol.forward(obs).numpy()
            pol.set_trainable_flat(flat)
            out_r.append(float(-np.mean((out - target) ** 2)))
    return out_r


def run_reference_real(args, w, ref_root, as_baseline=False):
    """--impl reference when the reference itself is importable: its own SharedNoiseTable, policy classes,
    FiniteDifferences.step and DSGD, unmodified, on the host cores.  Per step: (1) the worker loop of worker.py:26-32 over
    a bounded sample of the members on `cores` forked single-threaded processes (run_client.py:15), each member's E
    observations in one `policy.forward` call; (2) the unmodified learner step on a bounded number of returns; both
    scaled linearly to the full population (stated in `sample`)."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden", "_shim"))      # `gym` is absent from the image
    sys.path.insert(0, ref_root)
    from utils import SharedNoiseTable                    # noqa: E402  (reference)
    from utils import torch_helpers
    from policies import MujocoPolicy, DiscretePolicy, AtariPolicy, ImpalaPolicy
    from learner import FiniteDifferences, FDReturn
    from dsgd import DSGD
    import torch.nn as nn
    torch.set_num_threads(1)
    cores = len(os.sched_getaffinity(0))
    torch.manual_seed(TABLE_SEED)
    kind = w["kind"]
    if kind == "mujoco":
        h1, h2 = w["h1"], w["h2"]

        class _Widths(MujocoPolicy):
            """hidden widths as parameters (the reference hard-codes 64 x 64, mujoco.py:33-34; BASELINE config 3 names
            256 x 256); the 64 x 64 default builds the stock model"""
            def _build_model(self):
                self.model = nn.Sequential(nn.Linear(self.input_shape, h1), nn.Tanh(), nn.Linear(h1, h2), nn.Tanh(),
                                           nn.Linear(h2, self.output_shape * 2), torch_helpers.MapContinuousToAction())
        policy = MujocoPolicy(w["n_in"], w["n_act"], seed=TABLE_SEED) if (h1, h2) == (64, 64) else _Widths(w["n_in"], w["n_act"], seed=TABLE_SEED)
    elif kind == "discrete":
        policy = DiscretePolicy(w["n_in"], w["n_act"], seed=TABLE_SEED)
    elif kind == "atari":
        policy = AtariPolicy((84, 84), w["n_act"], seed=TABLE_SEED)
    else:
        policy = ImpalaPolicy((3, 64, 64), w["n_act"], seed=TABLE_SEED)
    P = policy.num_params

    class SignedTable(SharedNoiseTable):
        """antithetic keys '+i' / '-i' (the extension the B200 arm's batch uses) on the reference's own table"""
        def decode(self, key):
            key = str(key)
            if key[0] == "-":
                return -super().decode(key[1:])
            return super().decode(key[1:] if key[0] == "+" else key)
    table = SignedTable(args.table_size, P, TABLE_SEED)

    class Omega(object):
        omega, min_omega, max_omega = 0.0, 0.0, 1.0
    opt = DSGD(policy.parameters(), lr=LR)
    learner = FiniteDifferences(policy, opt, Omega(), table, noise_std=SIGMA, batch_size=w["members"], ent_coef=0.0,
                                max_delayed_return=H)
    M, E, R = w["members"], w["E"], w["pairs"]
    rng = np.random.RandomState(0)
    if kind == "atari":
        obs = torch.from_numpy(rng.rand(E, 4, 84, 84).astype(np.float32))
    elif kind == "impala":
        obs = torch.from_numpy(np.floor(rng.rand(E, 3, 64, 64) * 255.0).astype(np.float32))
    else:
        obs = torch.from_numpy(rng.randn(E, w["n_in"]).astype(np.float32))
    target = np.tanh(rng.randn(w["out_width"])).astype(np.float32) * 0.5
    # bounded sample: about a second of member evaluations and about a second of estimator per step
    per_member_s = {"mujoco": 2.5e-4 + 4e-9 * P, "discrete": 4e-4, "atari": 3e-3, "impala": 1.2e-2}[kind] * max(1.0, E / 16.0)
    sample = args.cpu_sample_members or int(min(M, max(cores * 4, min(cores * 32, cores * 1.0 / per_member_s))))
    n_fd = min(M, max(64, int(2.0e8 // P) // 2 * 2))
    steps, warm = (args.steps, args.warmup) if not as_baseline else (2, 1)
    warm = min(warm, 2)
    import multiprocessing as mp
    _R.update(policy=policy, table=table, obs=obs, kind=kind, target=target)
    times = []
    t_budget = time.perf_counter() + 240.0
    done_steps = 0
    import io
    import contextlib
    for it in range(warm + steps):
        pairs_idx = np.array([int(table.sample()[0]) for _ in range(R)], dtype=np.int64)
        idx = np.concatenate([pairs_idx, pairs_idx])
        sign = np.concatenate([np.ones(R), -np.ones(R)]).astype(np.int8)
        _R.update(idx=idx, sign=sign)
        flat = policy.get_trainable_flat().copy()
        members = list(range(0, M, max(1, M // sample)))[:sample]
        n_tasks = max(1, min(len(members), cores * 2))
        tasks = [(members[i::n_tasks], flat) for i in range(n_tasks)]
        pool = mp.get_context("fork").Pool(cores) if cores > 1 else None      # forked after _R holds this step's indices
        t0 = time.perf_counter()
        if pool is not None:
            chunks = pool.map(_real_member_chunk, tasks, chunksize=1)
            pool.close()
        else:
            chunks = [_real_member_chunk(t) for t in tasks]
        t_fwd = (time.perf_counter() - t0) * (M / len(members))
        rewards = rng.randn(M)
        got = [r for c in chunks for r in c]
        rewards[:len(got)] = got
        sel = np.concatenate([np.arange(n_fd // 2), R + np.arange(n_fd // 2)])
        batch = []
        for i, sgn, r in zip(idx[sel], sign[sel], rewards[sel]):
            ret = FDReturn()
            ret.epoch, ret.encoded_noise, ret.reward = learner.epoch, ("+%d" if sgn > 0 else "-%d") % i, float(r)
            batch.append(ret)
        t1 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            learner.step(batch, 0.0, 0.0, 0.0)                 # unmodified finite_differences.py:24-64 + DSGD.step
        t_fd = (time.perf_counter() - t1) * (M / n_fd)
        if it >= warm:
            times.append((t_fwd, t_fd))
            done_steps += 1
        if time.perf_counter() > t_budget and done_steps >= 1:
            break
    t_fwd = float(np.mean([t[0] for t in times]))
    t_fd = float(np.mean([t[1] for t in times]))
    step_s = t_fwd + t_fd
    sample_txt = ("unmodified reference from %s. forward: worker.py:26-32 loop over %d of %d members x %d obs per step on %d "
                  "forked processes x 1 torch thread, scaled x%.1f; estimator: FiniteDifferences.step + DSGD on %d of %d "
                  "returns (numpy BLAS threads), scaled linearly; %d timed steps"
                  % (os.path.relpath(ref_root, ROOT), len(members), M, E, cores, M / len(members), n_fd, M, done_steps))
    return dict(value=M * E / step_s, unit="env-steps/s", cores=cores, kind="reference", sample=sample_txt,
                ms_per_step=step_s * 1e3, fd_estimates_per_s=1.0 / t_fd, forward_s=t_fwd, estimator_s=t_fd,
                steps_run=done_steps)


def reference_main(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ref_root = None if args.reference_port else reference_root()
    if ref_root is not None:
        try:
            r = run_reference_real(args, w, ref_root)
        except Exception as e:          # an unimportable copy: say so and time the port instead
            sys.stderr.write("bench: reference at %s could not be run (%s: %s); timing the oracle port\n" % (ref_root, type(e).__name__, e))
            r = run_reference(args, w)
    else:
        r = run_reference(args, w)
    line = {
        "impl": "reference", "metric": "perturbed-policy env-steps/sec", "value": r["value"], "unit": "env-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "steps_run": r.get("steps_run"),
        "ms_per_step": r["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(args, w),
        "fd_estimates_per_s": r["fd_estimates_per_s"],
        "cpu_baseline": {"value": r["value"], "unit": "env-steps/s", "cores": r["cores"], "kind": r["kind"],
                         "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def bench_config(args, w):
    fd_mode = "fd_state (returns spread over the last %d epochs: theta-history distance rows + dot pass)" % H \
        if w["kind"] == "impala" else "fd_return (all returns from the current epoch)"
    return {"workload": w["desc"], "policy": "%s %d-%d-%d-%d" % (w["kind"], w["n_in"], w["h1"], w["h2"], w["out_width"]),
            "pairs_per_gpu": w["pairs"], "members_per_gpu": w["members"], "obs_per_member": w["E"],
            "table": "SharedNoiseTable(%d, P, %d)" % (args.table_size, TABLE_SEED), "sigma": SIGMA,
            "estimator": fd_mode + ", antithetic pairs merged per table row",
            "l2": "fresh noise indices and a different observation buffer each step; table replicas (4x table) and the "
                  "rotating observation buffers exceed the 126 MB L2; the reduction re-reads rows the forward of the same "
                  "step touched (L2 reuse by design)"}


# ----------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------
class ClockSampler(object):
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,utilization.gpu"

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for t, line in self.rows:
            if t < t0 or t > t1 + 0.1:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
# the B200 arm
# ----------------------------------------------------------------------------------------------
class Env(object):
    """What every workload of one bench process shares: rank layout, process group, device context, measured peaks."""
    pass


def b200_env(args):
    import torch
    import torch.distributed as dist
    env = Env()
    rank = env.rank = int(os.environ.get("RANK", "0"))
    world = env.world = int(os.environ.get("WORLD_SIZE", "1"))
    local = env.local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback (use --impl reference for the CPU port)")
    torch.cuda.set_device(local)
    env.affinity = bind_rank_to_cores(local, world)
    env.pg = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on the C-level stdout when the communicator is created: point fd 1 at stderr
        # for that moment so stdout carries the one JSON line only
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
        env.pg = dist.group.WORLD
    import __graft_entry__ as G
    if rank == 0:
        G.build()
    if world > 1:
        dist.barrier()
    from dfd_starter_b200.device import get_context
    env.ctx = get_context(local)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    env.hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    env.peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"
    return env


def bind_rank_to_cores(local, world):
    """N > 1: give every rank its own slice of the host cores (preferring the cores of the GPU's NUMA node), so eight
    Python chains + their pinned-memory pages do not migrate over each other.  Pinned buffers are allocated AFTER this,
    i.e. first-touched on the rank's own node.  Returns a short description for the JSON line."""
    try:
        avail = sorted(os.sched_getaffinity(0))
        if world <= 1 or len(avail) < 2 * world:
            return {"bound": False, "cores_available": len(avail)}
        import torch
        node_cpus = None
        try:
            bus = torch.cuda.get_device_properties(local).pci_bus_id
            dom = torch.cuda.get_device_properties(local).pci_domain_id
            dev_id = torch.cuda.get_device_properties(local).pci_device_id
            path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/local_cpulist" % (dom, bus, dev_id)
            if os.path.exists(path):
                cpus = []
                for part in open(path).read().strip().split(","):
                    lo, _, hi = part.partition("-")
                    cpus += list(range(int(lo), int(hi or lo) + 1))
                node_cpus = [c for c in cpus if c in avail]
        except Exception:
            node_cpus = None
        per = len(avail) // world
        mine = avail[local * per:(local + 1) * per]
        if node_cpus and len(node_cpus) >= per:
            # ranks whose GPUs share a node split that node's cores between them
            sharers = max(1, world // max(1, len(avail) // max(len(node_cpus), 1)))
            k = local % sharers
            per_n = max(1, len(node_cpus) // sharers)
            cand = node_cpus[k * per_n:(k + 1) * per_n]
            if cand:
                mine = cand
        os.sched_setaffinity(0, mine)
        return {"bound": True, "cores": len(mine), "first_core": mine[0], "numa_local": bool(node_cpus)}
    except Exception as e:       # never fatal
        return {"bound": False, "error": str(e)}


def run_workload(env, args, w, full=True):
    """One workload on this process group: device-resident steps (value), per-kernel timings and roofline, the untimed
    parity step, e2e through the reference-facing objects.  full=False (the other named configs riding in the same JSON
    line): fewer e2e steps, no clock sampling, no at-scale / cold reduction runs, no CPU baseline."""
    import torch
    import torch.distributed as dist
    import ctypes as C
    import dfd_starter_b200 as D
    from dfd_starter_b200 import _lib
    from dfd_starter_b200.device import ptr
    rank, world, local, pg, ctx = env.rank, env.world, env.local, env.pg, env.ctx
    lib, dev, hbm_peak, peak_src = ctx.lib, ctx.device, env.hbm_peak, env.peak_src
    M, E, R = w["members"], w["E"], w["pairs"]
    torch.manual_seed(TABLE_SEED)
    use_tc, tc_level = forward_precision(w["kind"], args.precision, E)
    if w["kind"] in ("mujoco", "discrete"):
        cls = D.MujocoPolicy if w["kind"] == "mujoco" else D.DiscretePolicy
        policy = cls(w["n_in"], w["n_act"], seed=TABLE_SEED, h1=w["h1"], h2=w["h2"], device=local,
                     precision=tc_level if use_tc else 0)
        obs_shape = (w["n_in"],)
    elif w["kind"] == "atari":
        policy = D.AtariPolicy((84, 84), w["n_act"], seed=TABLE_SEED, device=local, precision=1 if use_tc else 0)
        obs_shape = (4, 84, 84)
    else:
        policy = D.ImpalaPolicy((3, 64, 64), w["n_act"], seed=TABLE_SEED, device=local, precision=tc_level if use_tc else 0)
        obs_shape = (3, 64, 64)
    is_impala = w["kind"] == "impala"
    P = policy.num_params
    table = D.SharedNoiseTable(args.table_size, P, TABLE_SEED, device=local)
    policy.bind_table(table)
    if is_impala and use_tc:
        # the bench calls dfd_impala_forward directly: register the sigma-scaled fp16 table mirror the dense tail
        # streams its weight tiles from (ImpalaPolicy.forward_members_impala does this on first use)
        table.device_table.ensure_scaled16(SIGMA, P)

    class Omega(object):
        omega, min_omega, max_omega = 0.0, 0.0, 1.0
    opt = D.DSGD([torch.nn.Parameter(torch.zeros(1))], lr=LR)
    opt.coef = np.sqrt(P)
    # sharded population: the exchange runs over NVLink peer memory (dist.PeerExchange), no NCCL call on the step
    xchg = None
    if world > 1 and args.exchange == "peer":
        from dfd_starter_b200.dist import PeerExchange
        xchg = PeerExchange(ctx, P, pg)
    # fd_state (IMPALA, delayed returns): the per-return norms are rank-local, so the standardisation cannot be deferred:
    # rewards gathered over peer memory, then the gradient summed over peer memory ("general" mode, two peer kernels)
    learner = D.FiniteDifferences(policy, opt, Omega(), table, noise_std=SIGMA, batch_size=M, max_delayed_return=H,
                                  paired=True, process_group=pg, peer_exchange=xchg,
                                  exchange_mode="general" if (is_impala and xchg is not None) else "fd_return")

    CYC = H  # graphs / index sets / observation buffers cycle with the history ring
    g = torch.Generator().manual_seed(1234 + rank)
    # every rank draws from its own slice of the index stream (same table on every rank)
    for _ in range(rank):
        table.sample_indices(R * CYC)
    idx_sets = [table.sample_indices(R) for _ in range(CYC)]
    idx_host = torch.stack([torch.from_numpy(np.concatenate([i, i])) for i in idx_sets]).pin_memory()
    sign_host = torch.from_numpy(np.concatenate([np.ones(R), -np.ones(R)]).astype(np.int8)).pin_memory()
    n_obs_buf = (CYC if M * E * w["n_in"] * 4 * CYC < (8 << 30) else 3) if full else min(3, CYC)
    if w["kind"] in ("atari", "impala"):
        obs_host = torch.rand((n_obs_buf, M, E) + obs_shape, generator=g)
        if is_impala:
            obs_host = (obs_host * 255.0).floor()
        obs_host = obs_host.pin_memory()
    else:
        obs_host = torch.randn((n_obs_buf, M, E) + obs_shape, generator=g).pin_memory()
    idx_d = idx_host.to(dev)
    sign_d = sign_host.to(dev)
    obs_d = obs_host.to(dev)
    target = (torch.tanh(torch.randn(w["out_width"], generator=torch.Generator().manual_seed(7))) * 0.5).to(dev)
    out_d = torch.empty(M, E, w["out_width"], device=dev)
    reward_d = torch.empty(M, dtype=torch.float64, device=dev)
    stats_d = torch.empty(M * world, dtype=torch.float64, device=dev) if world > 1 else None
    hist_row_d = None
    if is_impala:
        zero_r = torch.zeros(M, E, device=dev)
        zero_done = torch.zeros(M, E, dtype=torch.uint8, device=dev)
        h_d = torch.zeros(M, E, 256, device=dev)
        c_d = torch.zeros(M, E, 256, device=dev)
        h1_d, c1_d = torch.empty_like(h_d), torch.empty_like(c_d)
        # fd_state mode: returns spread over the current epoch (-1) and the H history rows
        hist_row_d = torch.randint(-1, H, (M,), generator=g).to(torch.int32).to(dev)

    def forward_only(c):
        if is_impala:
            _lib.check(lib.dfd_impala_forward(ctx.handle, C.byref(policy.desc), table.device_table.ref(), ptr(policy.theta),
                                              ptr(policy.buffers), ptr(idx_d[c]), ptr(sign_d), M, SIGMA,
                                              ptr(obs_d[c % n_obs_buf]), ptr(zero_r), ptr(zero_done), ptr(h_d), ptr(c_d), E,
                                              ptr(out_d), ptr(h1_d), ptr(c1_d), None, 0, ctx.stream))
        else:
            policy.forward_members(idx_d[c], sign_d, obs_d[c % n_obs_buf], SIGMA, out=out_d)

    def device_step(k):
        c = k % CYC
        forward_only(c)
        _lib.check(lib.dfd_synthetic_reward(ctx.handle, ptr(out_d), M, E, w["out_width"], ptr(target), ptr(reward_d),
                                            ctx.stream))
        if world > 1 and xchg is None:
            dist.all_gather_into_tensor(stats_d, reward_d, group=pg)
        learner.step_device(idx_d[c], sign_d, reward_d, M, 0.0, hist_row_d=hist_row_d,
                            stats_d=stats_d if xchg is None else None)

    # NCCL collectives issued through torch.distributed are stream-ordered and graph-capturable, so the
    # sharded step replays as one graph too (plain launches remain the fallback if capture fails)
    use_graph = args.graph in ("on", "auto")
    # warm-up (fills the history ring so the ring position cycles with period H)
    n_warm = max(args.warmup, 3, H + 1)
    for k in range(n_warm):
        device_step(k)
    torch.cuda.synchronize()
    # ---------------- untimed parity step: this step's gradient against the fp64 closed form of the UNSHARDED batch ----
    parity = None
    if not args.profile_mode:
        try:
            parity = parity_step(env, w, table, learner, n_warm % CYC, idx_sets, hist_row_d, reward_d,
                                 lambda: device_step(n_warm))
        except Exception as e:       # reported, never fatal for the headline
            parity = {"error": "%s: %s" % (type(e).__name__, e)}
        n_warm += 1
        torch.cuda.synchronize()
    graphs = None
    launches_per_graph = LAUNCHES_PER_STEP
    if use_graph:
        try:
            graphs = []
            k0 = n_warm
            for c in range(CYC):
                gr = torch.cuda.CUDAGraph()
                l0 = ctx.launch_count()
                with torch.cuda.graph(gr):
                    device_step(k0 + c)
                launches_per_graph = ctx.launch_count() - l0      # this library's kernels captured into one step
                graphs.append(gr)
            torch.cuda.synchronize()
            n_warm += CYC          # capture advanced the learner's ring bookkeeping by CYC steps
        except Exception as e:     # plain launches are always available
            if rank == 0:
                sys.stderr.write("bench: CUDA-graph capture failed (%s); using plain launches\n" % e)
            graphs = None
            torch.cuda.synchronize()

    def run_step(k):
        if graphs is not None:
            graphs[k % CYC].replay()
        else:
            device_step(k)

    for k in range(3):
        run_step(n_warm + k)
    n_warm += 3
    torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0 and full:
        sampler.start()
        time.sleep(0.3)
    launches0 = ctx.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t_wall0 = time.perf_counter()
    ev0.record()
    for k in range(args.steps):
        run_step(n_warm + k)
    ev1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_wall1 = time.perf_counter()
    ms_total = ev0.elapsed_time(ev1)
    launches_plain = ctx.launch_count() - launches0
    if world > 1:
        t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    if args.profile_mode:
        return {"profile_mode": True, "ms_per_step": ms_step, "launches": launches_plain}
    # keep the same step running ~1.5 s so nvidia-smi (100 ms period) sees the clocks under this load
    t_load0 = time.perf_counter()
    kk = 0
    while full and time.perf_counter() - t_load0 < 1.5:
        for _ in range(50):
            run_step(n_warm + args.steps + kk)
            kk += 1
        torch.cuda.synchronize()
    t_load1 = time.perf_counter()
    clocks = sampler.stop(t_wall0, t_load1) if (rank == 0 and full) else None
    n_done = n_warm + args.steps + kk

    # ---------------- per-kernel durations (CUDA events on the launch stream, same inputs, back to back) ----
    def time_calls(fn, reps):
        fn(0)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        gr = None
        if use_graph:
            try:
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr):
                    for r in range(reps):
                        fn(r)
                torch.cuda.synchronize()
            except Exception:
                gr = None
        a.record()
        if gr is not None:
            gr.replay()
        else:
            for r in range(reps):
                fn(r)
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps * 1e3   # us per call

    reps = 4 * CYC
    rows = _lib.DfdFdRows(learner._row_ptr.data_ptr(), learner._row_coef.data_ptr(), learner._rows_cap)
    minus1 = torch.full((M,), -1, dtype=torch.int32, device=dev)
    from dfd_starter_b200.device import aligned_ptr
    # CYC prepared row lists so consecutive reduce launches stream different table rows
    rowsets = []
    for c in range(CYC):
        rp = torch.zeros(R + H, dtype=torch.int64, device=dev)
        rc = torch.zeros(R + H, dtype=torch.float32, device=dev)
        rs = _lib.DfdFdRows(rp.data_ptr(), rc.data_ptr(), R + H)
        _lib.check(lib.dfd_fd_prepare(ctx.handle, table.device_table.ref(), P, ptr(reward_d), ptr(idx_d[c]), ptr(sign_d),
                                      ptr(minus1), M, 1, 0.0, SIGMA, ptr(learner.dist), learner.Ps, 0, None, 0,
                                      C.byref(rs), aligned_ptr(learner._prep_scratch), learner._prep_scratch.numel() - 256,
                                      ctx.stream))
        rowsets.append((rp, rc, rs))
    grad_tmp = torch.empty(P, device=dev)

    def k_forward(r):
        forward_only(r % CYC)

    def k_reduce(r):
        _lib.check(lib.dfd_fd_reduce(ctx.handle, C.byref(rowsets[r % CYC][2]), R, P, ptr(grad_tmp),
                                     aligned_ptr(learner._red_scratch), learner._red_scratch.numel() - 256, ctx.stream))

    def k_prepare(r):
        _lib.check(lib.dfd_fd_prepare(ctx.handle, table.device_table.ref(), P, ptr(reward_d), ptr(idx_d[r % CYC]),
                                      ptr(sign_d), ptr(minus1), M, 1, 0.0, SIGMA, ptr(learner.dist), learner.Ps, 0, None,
                                      0, C.byref(rows), aligned_ptr(learner._prep_scratch),
                                      learner._prep_scratch.numel() - 256, ctx.stream))

    us_tail = None
    tail_scratch = learner._fused_scratch_for(M, 1) if world == 1 else None
    if tail_scratch is not None:
        def k_tail(r):      # the one-kernel learner step with lr = 0 (theta and the ring are left as they are)
            _lib.check(lib.dfd_fd_step_fused(ctx.handle, table.device_table.ref(), P, ptr(reward_d), ptr(idx_d[r % CYC]),
                                             ptr(sign_d), M, 1, 0.0, SIGMA, ptr(learner.theta), ptr(grad_tmp), 0.0, 1.0,
                                             ptr(learner.hist), ptr(learner.dist), learner.Ps, len(learner._hist_epoch), -1,
                                             ptr(learner._update_size), None, 0, 1, aligned_ptr(tail_scratch),
                                             tail_scratch.numel() - 256, ctx.stream))
        us_tail = time_calls(k_tail, reps)
    us_forward = time_calls(k_forward, reps)
    us_reduce = time_calls(k_reduce, reps)
    us_prepare = time_calls(k_prepare, reps)
    red_bytes = R * P * 4 + P * 4                       # rows*P*4 + P*4
    fwd_bytes = R * P * 4 + M * E * w["n_in"] * 4 + M * E * w["out_width"] * 4
    fwd_flops = M * E * layer_flops_per_obs(w)
    kernels = {
        "fd_reduce": {"us": us_reduce, "algorithmic_bytes": red_bytes, "achieved_GBps": red_bytes / us_reduce * 1e-3},
        "policy_forward": {"us": us_forward, "algorithmic_bytes": fwd_bytes, "flops": fwd_flops,
                           "achieved_GBps": fwd_bytes / us_forward * 1e-3, "achieved_TFLOPs": fwd_flops / us_forward * 1e-6},
        "fd_prepare": {"us": us_prepare},
    }
    if us_tail is not None:     # what the step actually launches for short parameter vectors (replaces prepare + reduce + DSGD)
        kernels["fd_step_fused"] = {"us": us_tail, "algorithmic_bytes": red_bytes, "achieved_GBps": red_bytes / us_tail * 1e-3}
    dominant = "policy_forward" if us_forward >= us_reduce else "fd_reduce"
    dk = kernels[dominant]
    roofline = {"kernel": dominant, "bound": "hbm", "achieved": dk["achieved_GBps"], "peak": hbm_peak, "unit": "GB/s",
                "frac": dk["achieved_GBps"] / hbm_peak,
                "traffic": None, "traffic_source": None, "peak_source": peak_src,
                "share_of_step": dk["us"] / (ms_step * 1e3),
                "fd_reduce": {"achieved": kernels["fd_reduce"]["achieved_GBps"], "frac": kernels["fd_reduce"]["achieved_GBps"] / hbm_peak,
                              "us": us_reduce, "algorithmic_bytes": red_bytes}}

    if E == WORKLOADS[w["name"]]["E"] and not args.pairs and use_tc == forward_precision(w["kind"], "auto", E)[0]:
        roofline["traffic"], roofline["traffic_source"] = ncu_traffic(w["name"], dominant)
        if roofline["traffic"] is not None:
            roofline["traffic_source"] += " (ncu --set full, one launch: dram__bytes_read.sum + dram__bytes_write.sum)"
    # ---------------- the reduction at a size where HBM, not launch latency, is the bound ----------------------
    # The default workload's reduction streams 25 MB (3.8 us of HBM time): it is latency bound.  The same kernel entry
    # point on the per-GPU reduction of BASELINE config 3 (1024 table rows x 171 042 parameters, 700 MB per launch, row
    # sets rotated so a launch never finds its rows in L2) shows what it does when bandwidth is the limit.
    if rank == 0 and full and args.table_size > 171042 + 1024:
        try:
            Pb, Rb, NSET = 171042, 1024, 6
            dt_ = table.device_table
            rng_b = np.random.RandomState(1)
            gb = torch.empty(Pb, device=dev)
            sb = ctx.zeros_bytes(lib.dfd_fd_reduce_scratch_bytes(ctx.handle, Pb, Rb))
            big_sets = []
            for c in range(NSET):
                ix = rng_b.randint(0, args.table_size - Pb, size=Rb).astype(np.int64)
                rp = torch.from_numpy(dt_.replicas.data_ptr() + 4 * ((ix & 3) * dt_.stride + (ix - (ix & 3)))).to(dev)
                rc = torch.from_numpy(rng_b.randn(Rb).astype(np.float32)).to(dev)
                big_sets.append((rp, rc, _lib.DfdFdRows(rp.data_ptr(), rc.data_ptr(), Rb)))

            def k_reduce_big(r):
                _lib.check(lib.dfd_fd_reduce(ctx.handle, C.byref(big_sets[r % NSET][2]), Rb, Pb, ptr(gb), aligned_ptr(sb),
                                             sb.numel() - 256, ctx.stream))
            us_big = min(time_calls(k_reduce_big, 4 * NSET) for _ in range(3))
            big_bytes = Rb * Pb * 4 + Pb * 4
            roofline["fd_reduce_at_scale"] = {
                "workload": "C3 per-GPU reduction: 1024 table rows x 171042 parameters (8192 pairs over 8 GPUs)",
                "algorithmic_bytes": big_bytes, "us": us_big, "achieved": big_bytes / us_big * 1e-3,
                "frac": big_bytes / us_big * 1e-3 / hbm_peak, "unit": "GB/s"}
            del big_sets, gb, sb
        except Exception as e:      # reported, never fatal for the headline
            roofline["fd_reduce_at_scale"] = {"error": str(e)}

    # ---------------- the reduction on rows that CANNOT be L2-resident ----------------------------------------------
    # Rows of the 25 M-entry table overlap (1024 rows x 684 KB drawn from 100 MB), so L2 serves part of every launch above.
    # Here every row is its own disjoint slice of a multi-GB buffer and consecutive launches walk through NSET different
    # row sets (>= 3.5 GB apart): each launch streams its algorithmic bytes from DRAM, nothing else.
    if rank == 0 and full:
        def reduce_cold(Pb, Rb, NSET):
            Pb4 = (Pb + 3) // 4 * 4
            cold = torch.randn(NSET * Rb * Pb4, device=dev)
            gb = torch.empty(Pb, device=dev)
            sb = ctx.zeros_bytes(lib.dfd_fd_reduce_scratch_bytes(ctx.handle, Pb, Rb))
            sets = []
            for c in range(NSET):
                rp = (cold.data_ptr() + 4 * Pb4 * (c * Rb + torch.arange(Rb, dtype=torch.int64))).to(dev)
                rc = torch.randn(Rb, device=dev)
                sets.append((rp, rc, _lib.DfdFdRows(rp.data_ptr(), rc.data_ptr(), Rb)))

            def k(r):
                _lib.check(lib.dfd_fd_reduce(ctx.handle, C.byref(sets[r % NSET][2]), Rb, Pb, ptr(gb), aligned_ptr(sb),
                                             sb.numel() - 256, ctx.stream))
            us = min(time_calls(k, 3 * NSET) for _ in range(3))
            # correctness of this launch shape against torch on the same rows (fp32 matmul, relative to the result's scale)
            k(0)
            want = sets[0][1].double() @ cold[:Rb * Pb4].view(Rb, Pb4)[:, :Pb].double()
            err = float((gb.double() - want).abs().max() / want.abs().max())
            nbytes = Rb * Pb * 4 + Pb * 4
            tr, src = ncu_traffic("cold", "fd_reduce") if Pb == 171042 else (None, None)
            return {"rows": Rb, "n_params": Pb, "algorithmic_bytes": nbytes, "us": us, "achieved": nbytes / us * 1e-3,
                    "frac": nbytes / us * 1e-3 / hbm_peak, "unit": "GB/s", "row_sets_cycled": NSET,
                    "buffer_GB": NSET * Rb * Pb4 * 4 / 1e9, "rel_err_vs_fp64_matvec": err, "traffic": tr, "traffic_source": src}
        for key, shape in (("C3_size", (171042, 1024, 6)), ("C5_size", (1158709, 256, 4))):
            try:
                roofline.setdefault("fd_reduce_cold", {})[key] = reduce_cold(*shape)
            except Exception as e:
                roofline.setdefault("fd_reduce_cold", {})[key] = {"error": str(e)}
            torch.cuda.empty_cache()

    # ---------------- e2e through the reference-facing objects, host buffers ---------------------------
    e2e = None
    if not args.no_e2e:
        TRACE = [] if os.environ.get("DFD_E2E_TRACE") else None

        class HostObsAgent(object):
            """obs from pinned host memory every call; returns come back to the host.  `prefetch`: the observations are
            double-buffered - step k+1's host->device copy is queued on a copy stream behind step k's, so it overlaps step
            k's forward, return read-back and learner step (an input pipeline; every step's copy still happens inside the
            timed region and the copy engine is the bound).  Without it the copy sits in front of the forward."""
            saved_states = []
            prefetch = True

            def __init__(self):
                self.copy_stream = torch.cuda.Stream(device=dev)
                self.bufs = [torch.empty_like(obs_d[0]) for _ in range(2)]
                self.ready = [torch.cuda.Event(), torch.cuda.Event()]
                self.staged = [None, None]            # which step's observations each buffer holds
                # member indices / signs go up through a pinned buffer read by a kernel (dfd_host_stage): a pageable
                # cudaMemcpyAsync would block the host behind the observation upload in flight on the copy engine
                self.small_host = torch.empty(M * 9 + 16, dtype=torch.uint8).pin_memory()
                self.small_dev = torch.empty(M * 9 + 16, dtype=torch.uint8, device=dev)
                self.reward_host = torch.empty(M, dtype=torch.float64).pin_memory()
                sh = self.small_host.numpy()
                self.idx_np, self.sign_np = sh[:M * 8].view(np.int64), sh[M * 8:M * 9].view(np.int8)
                self.i_d = self.small_dev[:M * 8].view(torch.int64)
                self.s_d = self.small_dev[M * 8:M * 9].view(torch.int8)
                self.reward_np = self.reward_host.numpy()

            def stage(self, k):
                b = k % 2
                if self.staged[b] == k:
                    return
                with torch.cuda.stream(self.copy_stream):
                    if TRACE is not None:
                        e = torch.cuda.Event(enable_timing=True); e.record(self.copy_stream); TRACE.append(("copy_start", k, e))
                    self.bufs[b].copy_(obs_host[(k % CYC) % n_obs_buf], non_blocking=True)
                    self.ready[b].record(self.copy_stream)
                    if TRACE is not None:
                        e = torch.cuda.Event(enable_timing=True); e.record(self.copy_stream); TRACE.append(("copy_end", k, e))
                self.staged[b] = k

            def collect_returns(self, pol, m_idx, m_sign, sigma):
                k = self.k
                if self.prefetch:
                    self.stage(k)
                    self.stage(k + 1)             # queued right behind step k's copy: its buffer was last read by step
                    #                               k-1's forward, which finished before step k-1's returns were read
                    torch.cuda.current_stream(dev).wait_event(self.ready[k % 2])
                    o = self.bufs[k % 2]
                else:
                    o = obs_host[(k % CYC) % n_obs_buf].to(dev, non_blocking=True)
                self.idx_np[:] = m_idx
                self.sign_np[:] = m_sign
                ctx.host_stage(self.small_host, self.small_dev)
                i_d, s_d = self.i_d, self.s_d
                if is_impala:
                    _lib.check(lib.dfd_impala_forward(ctx.handle, C.byref(pol.desc), table.device_table.ref(), ptr(pol.theta),
                                                      ptr(pol.buffers), ptr(i_d), ptr(s_d), M, sigma, ptr(o), ptr(zero_r),
                                                      ptr(zero_done), ptr(h_d), ptr(c_d), E, ptr(out_d), ptr(h1_d),
                                                      ptr(c1_d), None, 0, ctx.stream))
                else:
                    pol.forward_members(i_d, s_d, o, sigma, out=out_d)
                _lib.check(lib.dfd_synthetic_reward(ctx.handle, ptr(out_d), M, E, w["out_width"], ptr(target),
                                                    ptr(reward_d), ctx.stream))
                # the returns come back through SM writes into pinned memory (dfd_host_stage): a device-to-host
                # cudaMemcpyAsync completed only after the observation upload in flight (measured)
                if TRACE is not None:
                    e = torch.cuda.Event(enable_timing=True); e.record(); TRACE.append(("fwd_reward_end", k, e))
                ctx.host_stage(reward_d, self.reward_host)
                if TRACE is not None:
                    e = torch.cuda.Event(enable_timing=True); e.record(); TRACE.append(("reward_staged", k, e))
                    TRACE.append(("host_launched_all", k, time.perf_counter()))
                torch.cuda.current_stream(dev).synchronize()
                if TRACE is not None:
                    TRACE.append(("host_has_rewards", k, time.perf_counter()))
                rew = self.reward_np.copy()
                return {"reward": rew, "entropy": np.zeros(M), "timesteps": np.full(M, E), "states": None}
        agent = HostObsAgent()
        worker = D.Worker(policy, agent, table, None, sigma=SIGMA, eval_prob=0.0, random_seed=TABLE_SEED)
        learner._host_policy = True          # mirror theta to the host every step, as the drivers need it
        learner.policy = type("HostMirror", (), {"set_trainable_flat": staticmethod(lambda f: None)})()

        def e2e_step(k):
            agent.k = k
            if TRACE is not None:
                TRACE.append(("host_step_start", k, time.perf_counter()))
                e = torch.cuda.Event(enable_timing=True); e.record(); TRACE.append(("step_start", k, e))
            worker.epoch = learner.epoch
            flags = np.zeros(R, dtype=bool)
            idx = idx_sets[k % CYC]
            rets = worker.evaluate(flags, idx, antithetic=True)
            if is_impala:       # fd_state mode: spread the returns' epochs over the accepted window
                rets.epoch[:] = learner.epoch - (np.arange(len(rets)) % (H + 1))
            if world > 1 and xchg is not None:
                learner.step(rets, 0.0, 0.0, 0.0)           # the exchange kernel carries the statistics
            elif world > 1:
                allr = [None] * world
                dist.all_gather_object(allr, rets.reward)
                learner.step_arrays(rets.epoch, rets.idx, rets.sign, rets.reward, 0.0, all_rewards=np.concatenate(allr))
            else:
                learner.step(rets, 0.0, 0.0, 0.0)

        def e2e_run(k0, prefetch):
            agent.prefetch = prefetch
            agent.staged = [None, None]
            for k in range(3):
                e2e_step(k0 + k)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            for k in range(n_e2e):
                e2e_step(k0 + 3 + k)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if world > 1:
                t = torch.tensor([dt], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            return dt
        n_e2e = max(3, min(args.steps, 30 if full else 10))
        # a pinned page reaches full DMA speed only after the device has read it a few times (first reads of fresh pinned
        # memory ran at 34-47 GB/s on this box, 54.8 GB/s from the third pass on: scripts/h2d_probe.py); a long-running
        # worker reuses its staging buffers, so the staging buffers are read through before the timed loops
        for _ in range(4):
            for b in range(n_obs_buf):
                agent.bufs[0].copy_(obs_host[b], non_blocking=True)
        torch.cuda.synchronize()
        # the bound of the end-to-end step: its observations crossing PCIe (measured here, same buffers, copy engine)
        # N > 1: every rank copies at the same moment (barrier first), so this is the CONCURRENT host->device rate - the
        # floor the sharded end-to-end step can reach on this host - and the slowest rank's rate is reported beside it
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if world > 1:
            dist.barrier()
        ea.record()
        for b in range(8):
            agent.bufs[0].copy_(obs_host[b % n_obs_buf], non_blocking=True)
        eb.record()
        torch.cuda.synchronize()
        h2d_gbps = obs_host[0].numel() * 4 * 8 / (ea.elapsed_time(eb) * 1e-3) / 1e9
        h2d_gbps_min = h2d_gbps
        if world > 1:
            t = torch.tensor([h2d_gbps], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            h2d_gbps_min = float(t.item())
        dt_serial = e2e_run(n_done, False) if full else None
        if TRACE is not None:
            del TRACE[:]
        dt = e2e_run(n_done + 3 + n_e2e, True)
        if TRACE is not None and rank == 0:
            torch.cuda.synchronize()
            base_e = next(t for t in TRACE if t[0] == "step_start")
            base_h = next(t for t in TRACE if t[0] == "host_step_start")
            k_lo = base_e[1] + 8
            for name, k, v in TRACE:
                if k_lo <= k < k_lo + 4:
                    if isinstance(v, float):
                        sys.stderr.write("trace host   %-18s k=%d %9.1f us\n" % (name, k, (v - base_h[2]) * 1e6))
                    else:
                        sys.stderr.write("trace device %-18s k=%d %9.1f us\n" % (name, k, base_e[2].elapsed_time(v) * 1e3))
        h2d = obs_host[0].numel() * 4 + M * 8 + M + M * (8 + 8 + 4 + 1)
        d2h = M * 8 + 4 + P * 4
        e2e = {"value": M * E * world * n_e2e / dt, "unit": "env-steps/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": dt / n_e2e * 1e3, "steps": n_e2e,
               "serial_ms_per_step": None if dt_serial is None else dt_serial / n_e2e * 1e3,
               "pcie": {"h2d_GBps_measured": h2d_gbps, "h2d_GBps_slowest_rank_all_ranks_copying": h2d_gbps_min,
                        "floor_ms_per_step": h2d / (h2d_gbps_min * 1e9) * 1e3,
                        "frac_of_floor": (h2d / (h2d_gbps_min * 1e9)) / (dt / n_e2e)},
               "api": "Worker.evaluate -> ReturnBatch (sequence of FDReturn) -> FiniteDifferences.step (host observations, host "
                      "returns, theta mirrored to host); observations double-buffered: step k+1's pinned host->device copy is "
                      "issued on a copy stream while step k's returns are read back and the learner steps "
                      "(serial_ms_per_step: the same loop with the copy in front of each forward)"}

    # release what the workload held (graphs keep their memory pool alive) before the next one starts
    if xchg is not None and not full:
        try:
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            xchg.close()
        except Exception:
            pass
    value = M * E * world / (ms_step * 1e-3)
    tanh_txt = "tanh.approx.f32" if tc_level == 2 else "tanh to 1e-6"
    if not use_tc:
        dtype = "f32"
    elif w["kind"] == "impala":
        dtype = ("fp16 operands (10-bit mantissa as tf32), fp32 accumulate: convolutions as implicit GEMMs on tcgen05 kind::f16 (operand "
                 "maps in the UMMA layout, a filter tap = a shifted descriptor), dense tail (Linear + LSTM) on "
                 "tcgen05 kind::f16 with weight tiles by TMA from an fp16 repack of theta and the sigma-scaled fp16 table mirror; "
                 "fp32 BatchNorm folds, LSTM cell and head; f32 estimator")
    elif w["kind"] == "atari":
        dtype = ("fp16 operands (tcgen05 kind::f16, 10-bit mantissa as tf32; weight tiles by TMA from an fp16 copy of theta and "
                 "the sigma-scaled fp16 table mirror), fp32 accumulate, fp32 BatchNorm folds and head; f32 estimator")
    elif lib.dfd_policy_direct_supported(C.byref(policy.desc)):
        dtype = ("fp16 forward operands (tcgen05 kind::f16, 10-bit mantissa as tf32; weight tiles by TMA from an fp16 copy of "
                 "theta and the sigma-scaled fp16 table mirror), fp32 accumulate, fp32 biases, %s; f32 estimator" % tanh_txt)
    else:
        dtype = "tf32 forward operands, fp32 accumulate, %s; f32 estimator" % tanh_txt
    return {"value": value, "ms_per_step": ms_step, "dtype": dtype, "fd_estimates_per_s": 1e3 / ms_step,
            "launch_mode": "cuda-graph replay (one graph per history-ring position)" if graphs is not None else "plain launches",
            "roofline": roofline, "kernels": kernels, "e2e": e2e, "parity": parity,
            "gpu_launches": int(launches_plain if graphs is None else args.steps * launches_per_graph), "clocks": clocks,
            "E": E, "R": R}


def parity_step(env, w, table, learner, c, idx_sets, hist_row_d, reward_d, run_one_step):
    """One UNTIMED step whose gradient is compared with an independent fp64 evaluation of
    learner/finite_differences.py:40-49,80-112 over the UNSHARDED batch (all ranks' members):
        g = sum_i w_i * lam_i / ||lam_i||^2,  lam_i = fp32(s_i * sigma * eps_i) [+ dist_row(e_i)],  w = standardize(R - b)
    written here in numpy (bench.py never calls the oracle on the GPU arm).  Every rank's indices, rewards and epoch
    rows are gathered; rank 0 evaluates the closed form - on every coordinate for short parameter vectors, on a fixed
    sample of 16 384 coordinates (plus both ends) for long ones, the row norms always over the full rows - and all
    ranks' gradients are compared bit for bit."""
    import torch
    import torch.distributed as dist
    rank, world, pg = env.rank, env.world, env.pg
    dev = env.ctx.device
    P, R = learner.P, w["pairs"]
    fd_state = hist_row_d is not None
    dist_before = learner.dist[:, :P].clone() if fd_state else None
    run_one_step()
    if dev.type == "cuda":
        torch.cuda.synchronize()
    grad = learner.grad.clone()
    rew = reward_d.clone()
    idx_local = torch.from_numpy(np.ascontiguousarray(idx_sets[c])).to(dev)
    hr = hist_row_d.clone() if fd_state else None
    digest = (grad.view(torch.int32).to(torch.int64) * torch.arange(1, P + 1, device=dev)).sum().reshape(1)
    if world > 1:
        def gather(t):
            out = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(out, t.contiguous(), group=pg)
            return out
        idx_all, rew_all, dig_all = gather(idx_local), gather(rew), gather(digest)
        hr_all = gather(hr) if fd_state else None
    else:
        idx_all, rew_all, dig_all, hr_all = [idx_local], [rew], [digest], ([hr] if fd_state else None)
    identical = all(int(d.item()) == int(dig_all[0].item()) for d in dig_all)
    if rank != 0:
        return None
    t0 = time.perf_counter()
    tab = table._table
    sig32 = np.float32(SIGMA)
    g_dev = grad.double().cpu().numpy()
    full_cols = P <= 200_000
    cols = np.arange(P) if full_cols else np.unique(np.concatenate([
        np.arange(256), np.arange(P - 256, P), np.random.RandomState(5).randint(0, P, size=16384)]))
    rewards = np.concatenate([r.cpu().numpy() for r in rew_all])          # rank-major, each [plus | minus]
    x = rewards - 0.0                                                    # baseline b = 0 in the bench step
    s = x.std()
    wts = x if s == 0 else (x - x.mean()) / s
    g = np.zeros(cols.shape[0])
    drows = dist_before.cpu().numpy() if fd_state else None
    for r in range(world):
        ix = idx_all[r].cpu().numpy()
        wr = wts[r * 2 * R:(r + 1) * 2 * R]
        hrow = hr_all[r].cpu().numpy() if fd_state else None
        for j in range(R):
            eps = tab[ix[j]:ix[j] + P]
            lam_p = eps * sig32                                           # fp32 product, as finite_differences.py:89
            if not fd_state:
                # both members of the pair share ||lam||^2; lam- = -lam+
                n2 = float(np.dot(lam_p.astype(np.float64), lam_p.astype(np.float64)))
                g += ((wr[j] - wr[R + j]) / n2) * lam_p[cols]
                continue
            for member, sgn in ((j, 1.0), (R + j, -1.0)):
                lam = lam_p if sgn > 0 else -lam_p
                if hrow[member] >= 0:
                    lam = lam + drows[hrow[member]]                       # fp32 add (:89)
                l64 = lam.astype(np.float64)
                g += (wr[member] / float(np.dot(l64, l64))) * l64[cols]
    ref_max = float(np.max(np.abs(g)))
    rel = float(np.max(np.abs(g_dev[cols] - g)) / ref_max) if ref_max > 0 else float("nan")
    cos = float(np.dot(g_dev[cols], g) / (np.linalg.norm(g_dev[cols]) * np.linalg.norm(g) + 1e-300))
    return {"grad_rel_max": rel, "cosine": cos, "ranks_bit_identical": bool(identical), "ranks": world,
            "returns_checked": int(2 * R * world), "coordinates_checked": int(cols.shape[0]), "of": int(P),
            "tolerance": 1e-5, "ok": bool(rel <= 1e-5 and identical),
            "reference": "fp64 closed form of finite_differences.py:40-49,80-112 over the unsharded batch, evaluated in "
                         "bench.py (numpy) on rank 0", "seconds": time.perf_counter() - t0}


SUB_WORKLOADS = ("C3", "C4", "C5")
C2_POINTS = (("E1", dict(obs_per_member=1)), ("E16", dict(obs_per_member=16)), ("E128_fp32", dict(precision="fp32")))


def rng_noise_source_point(ctx, full=True, rows_only=False):
    """The reference drivers' DEFAULT noise source (run_sequential.py:89, run_server.py:78, run_client.py:123:
    RNGNoiseSource - numpy PCG64 keys, standard_normal(P) per member on the worker and again per return on the learner).
    Measured here: (1) one batch of member rows drawn on the device (RNGNoiseSource.sample_rows -> dfd_rng_normal_rows,
    bit-identical to numpy, tests/test_gpu_rng.py) at the C2 and C3 row shapes, CUDA events around the public call,
    next to numpy's rate on this host for the same draws (bounded sample); (2) the reference's own worker -> learner step
    (Worker.collect_returns(n) -> FiniteDifferences.step) with that noise source at the C2 shape, rows drawn on the
    device vs on the host."""
    import numpy as np
    import torch
    import dfd_starter_b200 as D
    dev = ctx.device
    out = {"source": "RNGNoiseSource (numpy Generator(PCG64).standard_normal, bit-exact on the device)", "rows": {}}
    for name, P, rows in (("C2", 6092, 2048), ("C3", 171042, 2048 if full else 128)):
        Ps = (P + 3) // 4 * 4
        src = D.RNGNoiseSource(P, TABLE_SEED)
        theta = torch.zeros(P, dtype=torch.float32, device=dev)
        buf = torch.empty(rows * Ps, dtype=torch.float32, device=dev)
        src.sample_rows(ctx, 4, buf, Ps, theta=theta, sigma=SIGMA)                     # warm-up
        times = []
        for _ in range(3):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            src.sample_rows(ctx, rows, buf, Ps, theta=theta, sigma=SIGMA)
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        ms = sorted(times)[1]
        n_cpu = min(rows * P, 4_000_000)
        g = np.random.default_rng(1)
        t0 = time.perf_counter()
        g.standard_normal(n_cpu)
        cpu_rate = n_cpu / (time.perf_counter() - t0)
        out["rows"][name] = {"rows": rows, "n_params": P, "normals": rows * P, "ms": ms, "normals_per_s": rows * P / ms * 1e3,
                             "rows_written_GBps": rows * P * 4 / ms * 1e-6, "numpy_normals_per_s_1core": cpu_rate,
                             "numpy_ms_extrapolated": rows * P / cpu_rate * 1e3}
        del buf
    if rows_only:
        return out
    # the reference's worker -> learner step with this noise source (C2 shape, 128 observations per member)
    wl = WORKLOADS["C2"]
    n, E = 2 * wl["pairs"], wl["E"]
    steps = {}
    for mode in ("device", "device_arrays", "host"):
        policy = D.MujocoPolicy(wl["n_in"], wl["n_act"], seed=TABLE_SEED, h1=wl["h1"], h2=wl["h2"], device=ctx.device_index, precision=0)
        P = policy.num_params
        src = D.RNGNoiseSource(P, TABLE_SEED, device=(mode != "host"))
        agent = D.SyntheticAgent(policy, E, seed=1)

        class Omega(object):
            omega, min_omega, max_omega = 0.0, 0.0, 1.0
        opt = D.DSGD([torch.nn.Parameter(torch.zeros(1))], lr=LR)
        opt.coef = np.sqrt(P)
        learner = D.FiniteDifferences(policy, opt, Omega(), src, noise_std=SIGMA, batch_size=n, max_delayed_return=3)
        worker = D.Worker(policy, agent, src, None, sigma=SIGMA, eval_prob=0.0, random_seed=TABLE_SEED)
        n_steps = 5 if mode != "host" else 2
        import contextlib
        import io
        for k in range(1 + n_steps):
            if k == 1:
                torch.cuda.synchronize()
                t0 = time.perf_counter()
            worker.epoch = learner.epoch
            rets = worker.collect_returns(n)
            with contextlib.redirect_stdout(io.StringIO()):
                learner.step(rets.non_eval() if mode == "device_arrays" else [r for r in rets if not r.is_eval], 0.0, 0.0, 0.0)
        torch.cuda.synchronize()
        steps[mode] = (time.perf_counter() - t0) / n_steps * 1e3
    out["worker_learner_step_C2"] = {"members": n, "obs_per_member": E, "ms_per_step_rows_on_device": steps["device"],
                                     "ms_per_step_rows_on_device_batch_as_arrays": steps["device_arrays"],
                                     "ms_per_step_rows_on_host": steps["host"],
                                     "env_steps_per_s_rows_on_device": n * E / steps["device"] * 1e3,
                                     "note": "per-return FDReturn objects as in the reference loop (rows_on_device / rows_on_host) or the Worker's ReturnBatch handed to the learner untouched (batch_as_arrays); exact fp32 forward"}
    return out


def _brief(res, w, world):
    """What a non-headline workload contributes to the JSON line."""
    rl = res["roofline"]
    out = {"workload": w["desc"], "obs_per_member": res["E"], "pairs_per_gpu": res["R"], "ms_per_step": res["ms_per_step"],
           "value": res["value"], "unit": "env-steps/s", "fd_estimates_per_s": res["fd_estimates_per_s"], "dtype": res["dtype"],
           "forward_us": res["kernels"]["policy_forward"]["us"], "reduce_us": res["kernels"]["fd_reduce"]["us"],
           "kernels": res["kernels"],
           "roofline": {k: rl.get(k) for k in ("kernel", "bound", "achieved", "peak", "unit", "frac", "traffic", "share_of_step",
                                               "fd_reduce")},
           "parity": res["parity"], "launch_mode": res["launch_mode"]}
    if res["e2e"] is not None:
        out["e2e"] = {k: res["e2e"].get(k) for k in ("value", "unit", "ms_per_step", "steps", "h2d_bytes_per_step",
                                                     "d2h_bytes_per_step", "pcie")}
    return out


def b200_main(args, w):
    import copy
    import gc
    import torch
    env = b200_env(args)
    rank, world = env.rank, env.world
    res = run_workload(env, args, w, full=True)
    if args.profile_mode:
        if rank == 0:
            print(json.dumps(res), flush=True)
        _finish(world)
        return
    line = None
    if rank == 0:
        line = {
            "metric": "perturbed-policy env-steps/sec", "value": res["value"], "unit": "env-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": res["dtype"], "data": "synthetic", "config": bench_config(args, w),
            "fd_estimates_per_s": res["fd_estimates_per_s"], "launch_mode": res["launch_mode"], "roofline": res["roofline"],
            "kernels": res["kernels"], "e2e": res["e2e"], "parity": res["parity"], "gpu_launches": res["gpu_launches"],
            "clocks": res["clocks"], "host": {"affinity": env.affinity},
        }
    # ---------------- the other named configs and operating points, same process group, same JSON line ----------------
    default_run = args.workload == "C2" and not args.obs_per_member and not args.pairs and args.precision == "auto"
    if default_run and not args.no_workloads:
        extra = []
        extra += [("operating_points", name, "C2", ov) for name, ov in C2_POINTS]
        extra += [("workloads", name, name, {}) for name in SUB_WORKLOADS]
        for group, name, wl, ov in extra:
            gc.collect()
            torch.cuda.empty_cache()
            a2 = copy.copy(args)
            a2.workload, a2.obs_per_member, a2.pairs = wl, ov.get("obs_per_member", 0), 0
            a2.precision = ov.get("precision", "auto")
            w2 = workload(a2)
            t0 = time.perf_counter()
            try:
                r2 = run_workload(env, a2, w2, full=False)
                entry = _brief(r2, w2, world) if rank == 0 else None
            except Exception as e:          # a failed side workload is reported, the headline stands
                entry = {"error": "%s: %s" % (type(e).__name__, e)}
            if rank == 0:
                entry["bench_seconds"] = time.perf_counter() - t0
                line.setdefault(group, {})[name] = entry
    if rank != 0:
        _xchg_profile(env.ctx, rank)
        _finish(world)
        return
    if default_run and not args.no_workloads and world == 1:
        gc.collect()
        torch.cuda.empty_cache()
        t0 = time.perf_counter()
        try:
            line["noise_sources"] = {"rng": rng_noise_source_point(env.ctx)}
        except Exception as e:
            line["noise_sources"] = {"rng": {"error": "%s: %s" % (type(e).__name__, e)}}
        line["noise_sources"]["rng"]["bench_seconds"] = time.perf_counter() - t0
    if not args.no_cpu_baseline and world == 1:
        try:
            cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload", args.workload,
                   "--obs-per-member", str(res["E"]), "--pairs", str(res["R"]), "--table-size", str(args.table_size),
                   "--steps", "2", "--warmup", "1"]
            outp = subprocess.run(cmd, capture_output=True, text=True, timeout=900,
                                  env={k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")})
            ref = json.loads(outp.stdout.strip().splitlines()[-1])
            line["cpu_baseline"] = ref["cpu_baseline"]
            line["cpu_baseline"]["fd_estimates_per_s"] = ref.get("fd_estimates_per_s")
        except Exception as e:
            line["cpu_baseline"] = {"value": None, "unit": "env-steps/s", "cores": None, "kind": "port",
                                    "sample": "failed: %s" % e}
    print(json.dumps(line), flush=True)
    _xchg_profile(env.ctx, rank)
    _finish(world)


def _xchg_profile(ctx, rank):
    if not os.environ.get("DFD_XCHG_PROF"):
        return
    import ctypes as C
    out = (C.c_double * 5)()
    if ctx.lib.dfd_xchg_profile(out) == 0:
        sys.stderr.write("[xchg prof] rank %d: push %.0f ns, publish %.0f ns, wait for peers %.0f ns, combine %.0f ns (mean of %d launches)\n"
                         % (rank, out[0], out[1], out[2], out[3], int(out[4])))


def _finish(world):
    """N > 1: leave together and skip communicator / IPC teardown - destroying an NCCL communicator while
    captured graphs that reference it are alive can block, and the driver only needs the JSON line."""
    if world <= 1:
        return
    import torch
    import torch.distributed as dist
    sys.stdout.flush()
    sys.stderr.flush()
    torch.cuda.synchronize()
    dist.barrier()
    os._exit(0)


# forward, synthetic return, fd_coef, fd_reduce, [sumsq,] dsgd_update  (fd_return mode: no dots pass); the value
# reported is counted from the library's launch counter during graph capture, this is only the fallback
LAUNCHES_PER_STEP = 6


def main():
    args = parse_args()
    w = workload(args)
    if args.impl == "reference":
        reference_main(args, w)
    else:
        b200_main(args, w)


if __name__ == "__main__":
    main()
