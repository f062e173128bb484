#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by EXECUTING THE REFERENCE.

Run in the build container only (needs /root/reference, which does not exist on
the GPU box):

    python tests/golden/make_golden.py

It imports the unmodified reference modules (utils.SharedNoiseTable,
policies.*, learner.FiniteDifferences, dsgd.DSGD, worker.Worker) through the
test-only `gym` shim in tests/golden/_shim and records inputs + reference
outputs as small .npz/.json fixtures.  The oracle (oracle/dfd_oracle.py) and the
CUDA path are then both checked against these files; nothing at test/bench time
reads /root/reference.
"""
import hashlib
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("DFD_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(HERE, "_shim"))
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)

from utils import SharedNoiseTable            # noqa: E402  (reference)
from policies import MujocoPolicy, DiscretePolicy, AtariPolicy, ImpalaPolicy  # noqa: E402
from learner import FiniteDifferences, FDReturn  # noqa: E402
from dsgd import DSGD                          # noqa: E402
from worker import Worker                      # noqa: E402
import torch.nn as nn                          # noqa: E402
from utils import torch_helpers                # noqa: E402

from oracle import dfd_oracle as O             # noqa: E402  (recipes for synthetic theta only)

torch.set_num_threads(1)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


# ---------------------------------------------------------------- noise tables
def gen_noise():
    out = {}
    for size, P, seed in [(1_000_000, 6092, 123), (200_000, 5197, 124), (25_000_000, 6092, 124),
                          (25_000_000, 678294, 124)]:
        t = SharedNoiseTable(size, P, seed)
        keys = [t.sample()[0] for _ in range(8)]
        i0 = int(keys[0])
        out["%d_%d_%d" % (size, P, seed)] = {
            "size": size, "n_params": P, "seed": seed,
            "sha256": sha(t._table),
            "head": [float(x) for x in t._table[:4]],
            "tail": [float(x) for x in t._table[-2:]],
            "keys": keys,
            "norm2_first": float(np.dot(t._table[i0:i0 + P].astype(np.float64), t._table[i0:i0 + P].astype(np.float64))),
        }
        del t
    return out


# ---------------------------------------------------------------- worker draws
class _StubStats(object):
    mean, std = 0.0, 1.0

    def serialize(self):
        return []


class _StubAgent(object):
    saved_states = []
    obs_stats = _StubStats()

    def collect_return(self, **kw):
        return 1.0, 0.0, 1


class _StubStrategy(object):
    def compute_novelty(self, policy):
        return 0.0


def gen_worker_draws():
    """Flags (worker stream) and keys (noise stream) of the reference Worker
    (worker/worker.py:19-38) driven like run_sequential.py:134-147."""
    torch.manual_seed(124)
    pol = MujocoPolicy(17, 6, seed=124)
    table = SharedNoiseTable(1_000_000, pol.num_params, 124)
    w = Worker(pol, _StubAgent(), table, _StubStrategy(), sigma=0.02, eval_prob=0.1, random_seed=124)
    w.epoch = 7
    flags, keys = [], []
    while len(keys) < 64:
        for ret in w.collect_returns():
            flags.append(bool(ret.is_eval))
            if not ret.is_eval:
                keys.append(ret.encoded_noise)
            else:
                assert ret.encoded_noise == "0"
    return {"size": 1_000_000, "seed": 124, "eval_prob": 0.1, "batch": 64, "flags": flags, "keys": keys}


# ---------------------------------------------------------------- MLP forwards
def member_outputs(pol, theta, table, idxs, signs, sigma, obs, fwd):
    thetas, outs = [], []
    for i, s in zip(idxs, signs):
        eps = table.decode(str(i))
        if s < 0:
            eps = -eps
        new_flat = theta + sigma * eps                 # worker/worker.py:28
        pol.set_trainable_flat(new_flat)
        thetas.append(pol.get_trainable_flat().copy())
        outs.append(fwd(pol, obs[len(outs)]))
    pol.set_trainable_flat(theta)
    return np.stack(thetas), outs


def gen_mujoco():
    torch.manual_seed(124)
    pol = MujocoPolicy(17, 6, seed=124)
    theta = pol.get_trainable_flat().copy()
    table = SharedNoiseTable(1_000_000, pol.num_params, 123)
    idxs = [int(table.sample()[0]) for _ in range(6)]
    signs = [1, 1, -1, 1, -1, 1]
    sigma = 0.02
    obs = torch.randn(6, 5, 17, generator=torch.Generator().manual_seed(0)).numpy()

    def fwd(p, x):
        with torch.no_grad():
            m, s = p.forward(x)
        return np.concatenate([m.numpy(), s.numpy()], -1)

    thetas, outs = member_outputs(pol, theta, table, idxs, signs, sigma, obs, fwd)
    np.savez(os.path.join(HERE, "mujoco_c2.npz"), theta=theta, idx=np.array(idxs), sign=np.array(signs, np.int8),
             sigma=sigma, obs=obs, theta_members=thetas, out=np.stack(outs), table_size=1_000_000, table_seed=123)


class _WideMujoco(MujocoPolicy):
    """Reference MujocoPolicy with the hidden widths of BASELINE config 3
    (the reference hard-codes 64x64, mujoco.py:33-34)."""

    def _build_model(self):
        h1 = h2 = 256
        self.model = nn.Sequential(nn.Linear(self.input_shape, h1), nn.Tanh(), nn.Linear(h1, h2), nn.Tanh(),
                                   nn.Linear(h2, self.output_shape * 2), torch_helpers.MapContinuousToAction())


def gen_humanoid():
    torch.manual_seed(124)
    pol = _WideMujoco(376, 17, seed=124)
    L = O.mujoco_layout(376, 17, 256, 256)
    assert L.num_params == pol.num_params == 171042
    theta = O.synthetic_theta(L, 11)
    table = SharedNoiseTable(1_000_000, pol.num_params, 123)
    idxs = [int(table.sample()[0]) for _ in range(3)]
    signs = [1, -1, 1]
    obs = torch.randn(3, 4, 376, generator=torch.Generator().manual_seed(1)).numpy()

    def fwd(p, x):
        with torch.no_grad():
            m, s = p.forward(x)
        return np.concatenate([m.numpy(), s.numpy()], -1)

    _, outs = member_outputs(pol, theta, table, idxs, signs, 0.02, obs, fwd)
    np.savez(os.path.join(HERE, "mujoco_c3.npz"), theta_seed=11, idx=np.array(idxs), sign=np.array(signs, np.int8),
             sigma=0.02, obs=obs, out=np.stack(outs), table_size=1_000_000, table_seed=123)


def load_buffers(pol, layout, buffers):
    sd = pol.state_dict()
    for e in layout.entries:
        if e.kind == "buffer":
            v = torch.from_numpy(buffers[e.offset:e.offset + e.numel].copy()).view(e.shape)
            sd[e.name] = v.to(sd[e.name].dtype)
    pol.load_state_dict(sd)


def gen_discrete():
    torch.manual_seed(124)
    pol = DiscretePolicy(2, 9, seed=124)
    L = O.discrete_layout(2, 9)
    assert L.num_params == pol.num_params
    # refresh running stats the way the drivers do (policy.py:31-34)
    vbn = torch.rand(64, 2, generator=torch.Generator().manual_seed(5))
    pol.compute_vbn(vbn)
    theta = pol.get_trainable_flat().copy()
    serialized = np.asarray(pol.serialize(), dtype=np.float32)
    table = SharedNoiseTable(200_000, pol.num_params, 124)
    idxs = [int(table.sample()[0]) for _ in range(5)]
    signs = [1, 1, -1, -1, 1]
    obs = torch.rand(5, 6, 2, generator=torch.Generator().manual_seed(2)).numpy()

    def fwd(p, x):
        with torch.no_grad():
            return p.forward(x).numpy()

    thetas, outs = member_outputs(pol, theta, table, idxs, signs, 0.02, obs, fwd)
    np.savez(os.path.join(HERE, "discrete_c1.npz"), theta=theta, serialized=serialized, idx=np.array(idxs),
             sign=np.array(signs, np.int8), sigma=0.02, obs=obs, theta_members=thetas, out=np.stack(outs),
             table_size=200_000, table_seed=124)


def gen_atari():
    torch.manual_seed(124)
    pol = AtariPolicy((84, 84), 6, seed=124)
    L = O.atari_layout(6)
    assert L.num_params == pol.num_params == 678294
    theta = O.synthetic_theta(L, 21)
    load_buffers(pol, L, O.synthetic_buffers(L, 22))
    table = SharedNoiseTable(1_000_000, pol.num_params, 123)
    idxs = [int(table.sample()[0]) for _ in range(3)]
    signs = [1, -1, 1]
    obs = torch.rand(3, 2, 4, 84, 84, generator=torch.Generator().manual_seed(3)).numpy()

    def fwd(p, x):
        with torch.no_grad():
            return p.forward(torch.from_numpy(x)).numpy()

    _, outs = member_outputs(pol, theta, table, idxs, signs, 0.02, obs, fwd)
    np.savez(os.path.join(HERE, "atari_c4.npz"), theta_seed=21, buffer_seed=22, idx=np.array(idxs),
             sign=np.array(signs, np.int8), sigma=0.02, obs_seed=3, out=np.stack(outs),
             table_size=1_000_000, table_seed=123)


def gen_impala():
    torch.manual_seed(124)
    pol = ImpalaPolicy((3, 64, 64), 15, seed=124)
    L = O.impala_layout(15)
    assert L.num_params == pol.num_params == 1158709
    # the reference's own registration order must match the oracle layout
    names = [k for k, _ in pol.named_parameters()]
    assert names == [e.name for e in L.entries if e.kind == "param"], "impala layout order"
    theta = O.synthetic_theta(L, 31)
    load_buffers(pol, L, O.synthetic_buffers(L, 32))
    table = SharedNoiseTable(2_000_000, pol.num_params, 123)
    idxs = [int(table.sample()[0]) for _ in range(2)]
    signs = [1, -1]
    g = torch.Generator().manual_seed(4)
    frames = torch.randint(0, 256, (2, 2, 3, 64, 64), generator=g).float().numpy()
    rewards = np.array([[0.5, -3.0], [2.0, 0.0]], np.float32)
    dones = np.array([[False, True], [False, False]])
    h0 = (0.3 * torch.randn(2, 2, 256, generator=g)).numpy()
    c0 = (0.3 * torch.randn(2, 2, 256, generator=g)).numpy()
    probs = np.zeros((2, 2, 15), np.float32)
    h1 = np.zeros((2, 2, 256), np.float32)
    c1 = np.zeros((2, 2, 256), np.float32)
    for m, (i, s) in enumerate(zip(idxs, signs)):
        eps = table.decode(str(i))
        if s < 0:
            eps = -eps
        pol.set_trainable_flat(theta + 0.02 * eps)
        for e in range(2):   # each env is one batch-1, T=1 call with its own carried state (impala.py:126,184)
            pol.model[0].state = (torch.from_numpy(h0[m, e]).view(1, 1, 256).clone(),
                                  torch.from_numpy(c0[m, e]).view(1, 1, 256).clone())
            inp = {"frame": torch.from_numpy(frames[m, e]).view(1, 1, 3, 64, 64),
                   "reward": torch.from_numpy(rewards[m, e:e + 1]).view(1, 1),
                   "done": torch.from_numpy(dones[m, e:e + 1]).view(1, 1)}
            with torch.no_grad():
                p = pol.forward(inp)
            probs[m, e] = p.view(-1).numpy()
            h1[m, e] = pol.model[0].state[0].view(-1).numpy()
            c1[m, e] = pol.model[0].state[1].view(-1).numpy()
    np.savez(os.path.join(HERE, "impala_c5.npz"), theta_seed=31, buffer_seed=32, idx=np.array(idxs),
             sign=np.array(signs, np.int8), sigma=0.02, frame_seed=4, reward=rewards, done=dones,
             probs=probs, h1=h1, c1=c1, table_size=2_000_000, table_seed=123)


# ---------------------------------------------------------------- estimator
class _Omega(object):
    def __init__(self, omega, lo=0.0, hi=1.0):
        self.omega, self.min_omega, self.max_omega = omega, lo, hi


class SignedTable(SharedNoiseTable):
    """Test-only antithetic decode: '+i' / '-i' keys."""

    def decode(self, key):
        key = str(key)
        if key[0] == "-":
            return -super().decode(key[1:])
        return super().decode(key)


def mkret(epoch, key, reward):
    r = FDReturn()
    r.epoch, r.encoded_noise, r.reward = epoch, key, reward
    return r


def gen_fd_steps():
    """Unmodified FiniteDifferences.step + DSGD on the C2 policy: 6 current-epoch
    steps (fd_return mode), then steps whose returns come from older epochs
    (fd_state mode, H=4 -> H+1 accepted epochs), including too-old and unknown
    epochs that must be discarded, an all-equal-reward batch (std == 0 branch),
    a None baseline, an empty batch, and one antithetic step."""
    import io
    import contextlib
    torch.manual_seed(124)
    pol = MujocoPolicy(17, 6, seed=124)
    P = pol.num_params
    table = SignedTable(1_000_000, P, 123)
    omega = _Omega(0.3)
    opt = DSGD(pol.parameters(), lr=0.01)
    H = 4
    fd = FiniteDifferences(pol, opt, omega, table, noise_std=0.02, batch_size=16, ent_coef=0.0, max_delayed_return=H)
    rrng = np.random.RandomState(0)
    erng = np.random.RandomState(1)
    rec = {"theta0": pol.get_trainable_flat().copy(), "H": H, "sigma": 0.02, "lr": 0.01, "omega": 0.3,
           "table_size": 1_000_000, "table_seed": 123}
    n_steps = 14
    for s in range(n_steps):
        N = 16
        keys = [table.sample()[0] for _ in range(N)]
        rewards = (rrng.randn(N) * 3.0 + 10.0).tolist()
        baseline = 0.1 * s
        if s < 6:
            epochs = [fd.epoch] * N
        elif s == 9:      # includes too-old, future and negative epochs -> discarded
            epochs = [fd.epoch - int(k) for k in erng.randint(0, H + 3, size=N)]
            epochs[3] = fd.epoch + 5
        elif s == 10:     # std == 0 branch: rewards pass through un-standardised
            epochs = [fd.epoch - int(k) for k in erng.randint(0, H + 1, size=N)]
            rewards = [2.5] * N
        elif s == 11:     # baseline None
            epochs = [fd.epoch - int(k) for k in erng.randint(0, H + 1, size=N)]
            baseline = None
        elif s == 12:     # antithetic pairs, mixed epochs per pair
            base = keys[:N // 2]
            keys = ["+" + k for k in base] + ["-" + k for k in base]
            ep = [fd.epoch - int(k) for k in erng.randint(0, 2, size=N // 2)]
            epochs = ep + ep
        else:
            epochs = [fd.epoch - int(k) for k in erng.randint(0, H + 1, size=N)]
        batch = [mkret(e, k, r) for e, k, r in zip(epochs, keys, rewards)]
        with contextlib.redirect_stdout(io.StringIO()):
            upd = fd.step(batch, baseline, 0.0, 0.0)
        rec["s%d_keys" % s] = np.array(keys)
        rec["s%d_epochs" % s] = np.array(epochs)
        rec["s%d_rewards" % s] = np.array(rewards)
        rec["s%d_baseline" % s] = np.nan if baseline is None else baseline
        rec["s%d_grad" % s] = fd.gradient_memory.copy()
        rec["s%d_theta" % s] = pol.get_trainable_flat().copy()
        rec["s%d_update" % s] = float(upd)
        rec["s%d_discarded" % s] = fd.discarded_returns
        rec["s%d_epoch_after" % s] = fd.epoch
    # empty batch: returns int 0, no update (finite_differences.py:30-31)
    before = pol.get_trainable_flat().copy()
    r0 = fd.step([], 0.0, 0.0, 0.0)
    assert r0 == 0 and np.array_equal(before, pol.get_trainable_flat()) and fd.epoch == n_steps
    rec["n_steps"] = n_steps
    np.savez_compressed(os.path.join(HERE, "fd_steps.npz"), **rec)


class _Numpy2Generator(object):
    """numpy >= 2 dropped the dict `Generator.__getstate__()` the reference RNGNoiseSource reads
    (utils/noise_sources.py:7,11,18).  This proxy gives the UNMODIFIED reference class the
    old protocol on top of `bit_generator.state`, so its own sample()/decode() code runs."""

    def __init__(self, gen):
        self._g = gen

    def __getstate__(self):
        return dict(self._g.bit_generator.state)

    def __setstate__(self, st):
        self._g.bit_generator.state = st

    def standard_normal(self, size=None):
        return self._g.standard_normal(size=size)


def gen_fd_steps_hostnoise():
    """Unmodified FiniteDifferences.step + DSGD driven by the two noise sources whose key is NOT a table index
    (utils/noise_sources.py:4-33): SimpleNoiseSource as shipped, RNGNoiseSource through the numpy-2 proxy above.
    Only seeds, keys (RNG), rewards and epochs are stored; the tests regenerate the noise from the seeds."""
    import io
    import contextlib
    from utils.noise_sources import RNGNoiseSource, SimpleNoiseSource
    rec = {}
    for name in ("simple", "rng"):
        torch.manual_seed(124)
        pol = MujocoPolicy(17, 6, seed=124)
        P = pol.num_params
        if name == "simple":
            src = SimpleNoiseSource(P, 321)
        else:
            src = RNGNoiseSource.__new__(RNGNoiseSource)
            src.rng = _Numpy2Generator(np.random.default_rng(np.random.SeedSequence(321)))
            src.base_state = src.rng.__getstate__()
            src.n_params = P
        H = 2
        opt = DSGD(pol.parameters(), lr=0.01)
        fd = FiniteDifferences(pol, opt, _Omega(0.3), src, noise_std=0.02, batch_size=12, ent_coef=0.0, max_delayed_return=H)
        rrng = np.random.RandomState(7)
        erng = np.random.RandomState(8)
        rec[name + "_theta0"] = pol.get_trainable_flat().copy()
        n_steps = 5
        for s in range(n_steps):
            N = 12
            keys = [src.sample()[0] for _ in range(N)]
            rewards = (rrng.randn(N) * 2.0 - 1.0).tolist()
            if s < 2:
                epochs = [fd.epoch] * N
            else:                      # delayed returns, one of them too old (discarded before it is decoded)
                epochs = [fd.epoch - int(k) for k in erng.randint(0, H + 1, size=N)]
                if s == 3:
                    epochs[5] = fd.epoch - H - 1
            batch = [mkret(e, k, r) for e, k, r in zip(epochs, keys, rewards)]
            with contextlib.redirect_stdout(io.StringIO()):
                upd = fd.step(batch, 0.05 * s, 0.0, 0.0)
            if name == "rng":
                rec["rng_s%d_keys" % s] = np.array(keys)
            rec["%s_s%d_epochs" % (name, s)] = np.array(epochs)
            rec["%s_s%d_rewards" % (name, s)] = np.array(rewards)
            rec["%s_s%d_grad" % (name, s)] = fd.gradient_memory.copy()
            rec["%s_s%d_theta" % (name, s)] = pol.get_trainable_flat().copy()
            rec["%s_s%d_update" % (name, s)] = float(upd)
            rec["%s_s%d_discarded" % (name, s)] = fd.discarded_returns
    rec.update(n_steps=5, H=2, sigma=0.02, lr=0.01, omega=0.3, seed=321, N=12)
    # RNGNoiseSource stream pins (SURVEY.md App. C uses seed 123)
    g = np.random.default_rng(np.random.SeedSequence(123))
    st = g.bit_generator.state["state"]
    rec["pcg_seed123_state"] = np.array(str(st["state"]))
    rec["pcg_seed123_inc"] = np.array(str(st["inc"]))
    rec["pcg_seed123_normals"] = g.standard_normal(8)
    np.savez_compressed(os.path.join(HERE, "fd_steps_hostnoise.npz"), **rec)



def gen_wire():
    """Bytes written by the reference's own message classes (learner/fd_return.py:25-39, networking/client.py:39-44,
    networking/server.py:128-142) and the outcome of its `ServerInterface.get_returns_batch` (server.py:64-95) on a
    scripted arrival sequence."""
    from networking.rpc_misc import client_server_interface_pb2 as pb
    from networking.server import ServerInterface
    from learner import FDState
    rng = np.random.RandomState(3)
    rets = []
    for j in range(40):
        r = FDReturn()
        r.epoch = int(rng.randint(-1, 9))
        kind = j % 5
        r.encoded_noise = ["%d" % rng.randint(0, 10 ** 7), "+%d" % rng.randint(0, 10 ** 7), "-%d" % rng.randint(0, 10 ** 7),
                           "%d,%d" % (rng.randint(1, 2 ** 62), rng.randint(1, 2 ** 62)), "0"][kind]
        r.reward = float(rng.randn() * 100)
        r.novelty = float(rng.rand()) if j % 3 else 0
        r.entropy = float(rng.randn()) if j % 4 else 0.0
        r.timesteps = int(rng.randint(0, 1000))
        r.is_eval = kind == 4
        if r.is_eval:
            r.eval_states = rng.randn(int(rng.randint(1, 4)), 3, 2).astype(np.float32) if j % 10 == 4 else []
        if j % 6 == 0:
            r.obs_stats_update = rng.randn(7).astype(np.float32).tolist()
        rets.append(r)
    rets[7].reward = -0.0
    rets[8].epoch = 0
    rets[8].timesteps = 0
    rec = {"n": len(rets)}
    msgs = [r.serialize_to_grpc() for r in rets]
    arr = pb.ReturnArray(rets=msgs)
    rec["array_bytes"] = np.frombuffer(arr.SerializeToString(), dtype=np.uint8)
    for j, m in enumerate(msgs):
        rec["ret%d_bytes" % j] = np.frombuffer(m.SerializeToString(), dtype=np.uint8)
    back = []
    for m in pb.ReturnArray.FromString(arr.SerializeToString()).rets:    # what the reference server reads back
        r = FDReturn()
        r.deserialize_from_grpc(m)
        back.append(r)
    rec["epoch"] = np.array([r.epoch for r in back])
    rec["key"] = np.array([r.encoded_noise for r in back])
    rec["reward"] = np.array([r.reward for r in back], dtype=np.float64)
    rec["novelty"] = np.array([r.novelty for r in back], dtype=np.float64)
    rec["entropy"] = np.array([r.entropy for r in back], dtype=np.float64)
    rec["timesteps"] = np.array([r.timesteps for r in back])
    rec["is_eval"] = np.array([r.is_eval for r in back])
    for j, r in enumerate(back):
        rec["ret%d_states" % j] = np.asarray(r.eval_states, dtype=np.float32)
        rec["ret%d_stats" % j] = np.asarray(list(r.obs_stats_update), dtype=np.float32)
    # an unpacked encoding of the repeated fields (legal proto3 input; exercises the general decoder)
    # ServerState as the servicer builds it
    st = FDState()
    st.strategy_frames = rng.randn(3, 2).astype(np.float32)
    st.strategy_history = rng.randn(2, 3, 4).astype(np.float32)
    st.policy_params = rng.randn(50).astype(np.float32).tolist()
    st.epoch = 17
    st.experiment_id = "exp-a1"
    st.obs_stats = rng.randn(5).astype(np.float32).tolist()
    st.cfg = {"env_id": "Walker2d-v2", "noise_std": 0.02, "normalize_obs": True, "random_seed": 124, "eval_prob": 0.05}
    si = ServerInterface(st)
    ss = si.server_state
    msg = pb.ServerState(strategy_frames=ss.strategy_frames, strategy_frames_shape=ss.strategy_frames_shape,
                         strategy_history=ss.strategy_history, strategy_history_shape=ss.strategy_history_shape,
                         policy_parameters=ss.policy_params, epoch=ss.epoch, experiment_id=ss.experiment_id,
                         obs_stats=ss.obs_stats)
    rec["state_bytes"] = np.frombuffer(msg.SerializeToString(), dtype=np.uint8)
    rec["state_frames"] = st.strategy_frames
    rec["state_history"] = st.strategy_history
    rec["state_params"] = np.asarray(st.policy_params, dtype=np.float32)
    rec["state_obs_stats"] = np.asarray(st.obs_stats, dtype=np.float32)
    si.grpc_cfg.params["random_seed"] += 1
    rec["config_bytes"] = np.frombuffer(si.grpc_cfg.SerializeToString(deterministic=True), dtype=np.uint8)
    # get_returns_batch scenarios: arrivals in order 0..39 (ids = position), then three pulls
    for r in back:
        si.submit_return(r)
    ids = {id(r): j for j, r in enumerate(back)}
    pulls = [(5, 8, 3), (6, 8, None), (4, None, None)]
    for q, (bs, cur, mdr) in enumerate(pulls):
        got, ts, n_del, n_disc = si.get_returns_batch(batch_size=bs, current_epoch=cur, max_delayed_return=mdr)
        rec["pull%d_ids" % q] = np.array([ids[id(r)] for r in got])
        rec["pull%d_stats" % q] = np.array([ts, n_del, n_disc])
    rec["pulls"] = np.array([[bs, -99 if cur is None else cur, -99 if mdr is None else mdr] for bs, cur, mdr in pulls])
    rec["left"] = len(si.waiting_returns)
    np.savez_compressed(os.path.join(HERE, "wire.npz"), **rec)



def gen_strategy():
    """The reference's StrategyHandler / SparseHistoryManager (strategy/*.py) driven through a scripted sequence with a
    small history (so replacements happen), for the two driver configurations (init_helper.py:12-30): MuJoCo MLP with
    the Gaussian Wasserstein distance and the discrete MLP with the categorical TVD; plus every distance function of
    utils/math_helpers.py:166-222 on random strategies."""
    from strategy import StrategyHandler
    from utils import math_helpers
    rec = {}
    for name in ("mujoco", "discrete"):
        torch.manual_seed(124)
        if name == "mujoco":
            pol = MujocoPolicy(17, 6, seed=124)
            fn = math_helpers.gaussian_wasserstein_dist_from_strategies
            zeta = np.random.RandomState(5).randn(7, 17).astype(np.float32)
        else:
            pol = DiscretePolicy(2, 9, seed=124)
            fn = math_helpers.categorical_tvd
            zeta = np.random.RandomState(5).rand(7, 2).astype(np.float32)
            rec["discrete_serialized"] = np.asarray(pol.serialize(), dtype=np.float32)
        P = pol.num_params
        rng = np.random.RandomState(6)
        handler = StrategyHandler(pol, fn, max_history_size=4)
        rec[name + "_theta0"] = pol.get_trainable_flat().copy()
        rec[name + "_zeta"] = zeta
        n_events = 12
        for t in range(n_events):
            flat = (pol.get_trainable_flat() + (0.02 if t % 3 else 0.2) * rng.randn(P)).astype(np.float32)
            pol.set_trainable_flat(flat)
            rec["%s_e%d_flat" % (name, t)] = flat
            n_before = len(handler.strategy_history_manager.strategy_points)
            res = handler.strategy_history_manager.submit_policy(pol)
            rec["%s_e%d_submit" % (name, t)] = -2 if res is None else int(res)
            if t in (1, 3, 6, 9):
                handler.set_zeta(zeta)
            rec["%s_e%d_tensor" % (name, t)] = np.asarray(handler.strategy_tensor, dtype=np.float32).copy()
            rec["%s_e%d_worst" % (name, t)] = handler.strategy_history_manager.worst_point_idx
            probe = (flat + 0.05 * rng.randn(P)).astype(np.float32)
            keep = pol.get_trainable_flat().copy()
            pol.set_trainable_flat(probe)
            rec["%s_e%d_probe" % (name, t)] = probe
            rec["%s_e%d_novelty" % (name, t)] = float(handler.compute_novelty(pol))
            pol.set_trainable_flat(keep)
        rec[name + "_n_events"] = n_events
    rng = np.random.RandomState(9)
    cat_a = rng.dirichlet(np.ones(9), size=(11,)).astype(np.float32)
    cat_b = rng.dirichlet(np.ones(9), size=(5, 11)).astype(np.float32)
    ga = np.concatenate([rng.randn(11, 6), 0.1 + rng.rand(11, 6)], -1).astype(np.float32)
    gb = np.concatenate([rng.randn(5, 11, 6), 0.1 + rng.rand(5, 11, 6)], -1).astype(np.float32)
    rec.update(cat_a=cat_a, cat_b=cat_b, gauss_a=ga, gauss_b=gb)
    rec["d_l2_dist"] = math_helpers.l2_dist(ga, gb)
    rec["d_categorical_tvd"] = math_helpers.categorical_tvd(cat_a, cat_b)
    rec["d_gaussian_wasserstein_dist_from_strategies"] = math_helpers.gaussian_wasserstein_dist_from_strategies(ga, gb)
    rec["d_categorical_bhattacharrya_dist"] = math_helpers.categorical_bhattacharrya_dist(cat_a, cat_b)
    rec["d_gaussian_bhattacharrya_dist"] = math_helpers.gaussian_bhattacharrya_dist(ga[None], gb)
    rec["novelty_tvd"] = math_helpers.compute_strategy_novelty(cat_a, cat_b, distance_fn=math_helpers.categorical_tvd)
    np.savez_compressed(os.path.join(HERE, "strategy.npz"), **rec)



def gen_obs_stats():
    """WelfordRunningStat (utils/math_helpers.py:7-105) and the rollout loop's normalisation (worker/agent.py:37-41) run by
    the reference on fp32 observations: per-member sequential updates, the learner-wide merge, mean / std, and the
    normalised observations."""
    from utils.math_helpers import WelfordRunningStat
    rng = np.random.RandomState(11)
    M, E, K = 6, 23, 17
    obs = (rng.randn(M, E, K) * np.linspace(0.1, 30, K) + np.linspace(-5, 5, K)).astype(np.float32)
    obs[:, :, 3] = 2.5                                    # a constant feature: variance 0 -> std 1 (math_helpers.py:62)
    select = rng.uniform(0, 1, size=(M, E)) < 0.4
    select[1, :] = False                                  # a member that drew nothing: count 0, merge skipped
    select[2, :] = False
    select[2, 5] = True                                   # a single sample: count 1
    rows = []
    for m in range(M):
        st = WelfordRunningStat(K)
        for e in range(E):
            if select[m, e]:
                st.increment(obs[m, e], 1)
        rows.append(st.serialize())
    glob = WelfordRunningStat(K)
    snaps = []
    for r in rows:
        glob.increment_from_obs_stats_update(r)
        snaps.append(np.asarray(glob.serialize(), dtype=np.float64))
    mean, std = np.asarray(glob.mean), np.asarray(glob.std)
    normed = np.clip(np.subtract(obs, mean) / std, -10, 10)
    assert normed.dtype == np.float32
    back = WelfordRunningStat(K)
    back.deserialize(glob.serialize())
    # the worker's own path: statistics deserialised from the state (float64 arrays), fp64 normalisation, one rounding
    # to fp32 when the observation enters the policy (worker.py:43,47; agent.py:40-41; policy.py:28)
    wire_stats = np.asarray(glob.serialize(), dtype=np.float32).tolist()      # `repeated float` on the wire (proto:27)
    wst = WelfordRunningStat(K)
    wst.deserialize(wire_stats)
    normed_worker = torch.as_tensor(np.clip(np.subtract(obs, wst.mean) / wst.std, -10, 10), dtype=torch.float32).numpy()
    assert np.asarray(wst.std).dtype == np.float64
    np.savez_compressed(os.path.join(HERE, "obs_stats.npz"), obs=obs, select=select, rows=np.asarray(rows, dtype=np.float64),
                        wire_stats=np.asarray(wire_stats, dtype=np.float64), normed_worker=normed_worker,
                        worker_mean=np.asarray(wst.mean), worker_std=np.asarray(wst.std),
                        merged=np.asarray(snaps), mean=mean, std=std, normed=normed,
                        back_mean=np.asarray(back.mean, dtype=np.float64), back_std=np.asarray(back.std, dtype=np.float64))


# ---------------------------------------------------------------- round 2: compute_vbn, non-DSGD optimizer, default init
def gen_vbn():
    """policy.py:31-34 / impala.py:13-17: train-mode forward of the unperturbed policy refreshes every BatchNorm's
    running statistics.  Records the buffers before / after for Atari (tensor batch) and IMPALA (list of dicts)."""
    rec = {}
    torch.manual_seed(124)
    pol = AtariPolicy((84, 84), 6, seed=124)
    L = O.atari_layout(6)
    pol.set_trainable_flat(O.synthetic_theta(L, 21))
    load_buffers(pol, L, O.synthetic_buffers(L, 22))
    x = torch.rand(6, 4, 84, 84, generator=torch.Generator().manual_seed(41))
    with torch.no_grad():
        pol.compute_vbn(x)
        pol.compute_vbn(x[:3])
    assert not pol.training
    rec.update(atari_theta_seed=21, atari_buffer_seed=22, atari_obs_seed=41,
               atari_serialized_after=np.asarray(pol.serialize(), dtype=np.float32))
    torch.manual_seed(124)
    pol = ImpalaPolicy((3, 64, 64), 15, seed=124)
    L = O.impala_layout(15)
    pol.set_trainable_flat(O.synthetic_theta(L, 31))
    load_buffers(pol, L, O.synthetic_buffers(L, 32))
    g = torch.Generator().manual_seed(42)
    n = 4
    frames = torch.randint(0, 256, (n, 1, 1, 3, 64, 64), generator=g).float()
    rewards = torch.tensor([0.5, -3.0, 2.0, 0.0]).view(n, 1, 1)
    dones = torch.tensor([False, True, False, False]).view(n, 1, 1)
    buf = [{"frame": frames[i], "reward": rewards[i], "done": dones[i]} for i in range(n)]
    pol.reset()
    with torch.no_grad():
        pol.compute_vbn(buf)
    rec.update(impala_state_h=pol.model[0].state[0].reshape(-1).numpy().copy(),
               impala_state_c=pol.model[0].state[1].reshape(-1).numpy().copy())
    rec.update(impala_theta_seed=31, impala_buffer_seed=32, impala_frame_seed=42, impala_reward=rewards.view(-1).numpy(),
               impala_done=dones.view(-1).numpy(), impala_serialized_after=np.asarray(pol.serialize(), dtype=np.float32))
    # keep the fixture small: only the buffer values (state_dict order) are compared
    for k, LL in (("atari", O.atari_layout(6)), ("impala", O.impala_layout(15))):
        full = rec.pop(k + "_serialized_after")
        off, vals = 0, []
        for e in LL.entries:
            if e.kind == "buffer":
                vals.append(full[off:off + e.numel])
            off += e.numel
        rec[k + "_buffers_after"] = np.concatenate(vals)
    np.savez_compressed(os.path.join(HERE, "vbn.npz"), **rec)


def gen_adam_steps():
    """finite_differences.py:54-57 with a stock torch optimizer (not DSGD): policy.set_grad_from_flat(-g)
    (policy.py:63-70) then optimizer.step().  4 steps of the unmodified learner with torch.optim.Adam."""
    import io
    import contextlib
    torch.manual_seed(124)
    pol = MujocoPolicy(17, 6, seed=124)
    P = pol.num_params
    table = SharedNoiseTable(1_000_000, P, 123)
    opt = torch.optim.Adam(pol.parameters(), lr=0.01)
    fd = FiniteDifferences(pol, opt, _Omega(0.3), table, noise_std=0.02, batch_size=16, ent_coef=0.0, max_delayed_return=3)
    rrng = np.random.RandomState(7)
    rec = {"theta0": pol.get_trainable_flat().copy(), "H": 3, "sigma": 0.02, "lr": 0.01, "table_size": 1_000_000,
           "table_seed": 123, "n_steps": 4}
    for s in range(4):
        keys = [table.sample()[0] for _ in range(16)]
        rewards = (rrng.randn(16) * 2.0 + 1.0).tolist()
        epochs = [fd.epoch] * 16 if s < 2 else [fd.epoch - (i % 3) for i in range(16)]
        with torch.enable_grad(), contextlib.redirect_stdout(io.StringIO()):
            upd = fd.step([mkret(e, k, r) for e, k, r in zip(epochs, keys, rewards)], 0.0, 0.0, 0.0)
        rec["s%d_keys" % s] = np.array(keys)
        rec["s%d_epochs" % s] = np.array(epochs)
        rec["s%d_rewards" % s] = np.array(rewards)
        rec["s%d_grad" % s] = fd.gradient_memory.copy()
        rec["s%d_theta" % s] = pol.get_trainable_flat().copy()
        rec["s%d_update" % s] = float(upd)
    np.savez_compressed(os.path.join(HERE, "fd_steps_adam.npz"), **rec)


def gen_default_init():
    """Initial parameter vectors of the reference constructors under torch.manual_seed(124) (IMPALA: modules are
    CONSTRUCTED stage by stage - impala.py:62-107 - but registered list by list, :109-111)."""
    rec = {}
    for name, ctor in (("impala", lambda: ImpalaPolicy((3, 64, 64), 15, seed=124)),
                       ("atari", lambda: AtariPolicy((84, 84), 6, seed=124))):
        torch.manual_seed(124)
        th = ctor().get_trainable_flat()
        rec[name + "_sha256"] = sha(th)
        rec[name + "_head"] = th[:8].copy()
        rec[name + "_probe_idx"] = np.linspace(0, th.shape[0] - 1, 64).astype(np.int64)
        rec[name + "_probe"] = th[rec[name + "_probe_idx"]].copy()
    np.savez(os.path.join(HERE, "default_init.npz"), **rec)


def main():
    only = sys.argv[1:]
    if only:       # regenerate just the named fixtures: python make_golden.py vbn adam_steps default_init
        for name in only:
            globals()["gen_" + name]()
        print("golden fixtures written to", HERE, only)
        return
    with open(os.path.join(HERE, "noise.json"), "w") as f:
        json.dump({"tables": gen_noise(), "worker": gen_worker_draws()}, f, indent=1)
    gen_mujoco()
    gen_humanoid()
    gen_discrete()
    gen_atari()
    gen_impala()
    gen_fd_steps()
    gen_fd_steps_hostnoise()
    gen_wire()
    gen_strategy()
    gen_obs_stats()
    gen_vbn()
    gen_adam_steps()
    gen_default_init()
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
