"""CPU oracle for the dfd-starter finite-difference learner hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs may import it, and only as the checker / the timed CPU baseline.  The
product package (`dfd_starter_b200/`) never imports anything from `oracle/`
and has no CPU fallback.

It is a plain numpy / torch-CPU *restatement* (no code copied) of what the
reference does on this path; every function cites the reference file:line it
follows (paths relative to the reference root).  The reference is 100 % Python
(numpy + torch CPU), so the restatement is numpy for the integer / estimator
arithmetic and torch-CPU functional ops for the fp32 network forwards.

Parity pinning: the reference's own tests pin nothing on this path
(SURVEY.md §4), so the oracle is pinned against outputs of the *reference
itself*, executed in the build container by `tests/golden/make_golden.py`
(committed) and stored as fixtures under `tests/golden/`; see
`tests/test_oracle_golden.py`.
"""
from __future__ import annotations

import hashlib
import math
from dataclasses import dataclass, field

import numpy as np

# --------------------------------------------------------------------------
# a1-a3  noise table            reference: utils/noise_sources.py:36-51
# --------------------------------------------------------------------------


class RNGNoiseSourceOracle(object):
    """`RNGNoiseSource` restated (utils/noise_sources.py:4-20) for numpy >= 2: key = PCG64 `state,inc` before the
    draw, noise = `standard_normal(P)` (fp64); `decode` rewinds the same generator and redraws.  The reference reads
    the words through `Generator.__getstate__()`, whose layout changed in numpy 2; the generator
    words and the normal stream are pinned by SURVEY.md App. C."""

    def __init__(self, n_params, random_seed=123):
        self.rng = np.random.default_rng(np.random.SeedSequence(random_seed))
        self.n_params = n_params

    def sample(self):
        st = self.rng.bit_generator.state["state"]
        return "{},{}".format(st["state"], st["inc"]), self.rng.standard_normal(size=self.n_params)

    def decode(self, key):
        s, i = str(key).split(",")
        self.rng.bit_generator.state = {"bit_generator": "PCG64", "state": {"state": int(s), "inc": int(i)},
                                        "has_uint32": 0, "uinteger": 0}
        return self.rng.standard_normal(size=self.n_params)


class SimpleNoiseSourceOracle(object):
    """`SimpleNoiseSource` restated (utils/noise_sources.py:23-33): `RandomState(seed).randn(P)` per draw (fp64), the key
    is the vector itself and `decode` is the identity."""

    def __init__(self, n_params, random_seed=123):
        self.rng = np.random.RandomState(random_seed)
        self.n_params = n_params

    def sample(self):
        noise = self.rng.randn(self.n_params)
        return noise, noise

    def decode(self, noise):
        return noise


class NoiseTableOracle(object):
    """`SharedNoiseTable` restated (utils/noise_sources.py:36-51).

    One legacy `RandomState(seed)` first fills the table (`randn(size)` cast to
    fp32, :39-40) and then serves every index draw (`randint(0, size - P)`,
    :45), so indices depend on `size` as well as `seed`.
    """

    def __init__(self, size, n_params, random_seed=123):
        assert size > n_params
        self.rng = np.random.RandomState(random_seed)
        self.table = self.rng.randn(size).astype(np.float32)
        self.n_params = n_params
        self.max_idx = size - n_params

    def sample(self):
        idx = self.rng.randint(0, self.max_idx)
        return "{}".format(idx), self.table[idx:idx + self.n_params]

    def decode(self, key):
        """Keys are decimal strings (:49-51).  A leading '+'/'-' is the
        antithetic extension (SURVEY.md §8c `SignedTable`): '-123' means
        `-table[123:123+P]`.  Plain keys behave exactly as the reference."""
        key = str(key)
        if key[0] == "-":
            i = int(key[1:])
            return -self.table[i:i + self.n_params]
        i = int(key)
        return self.table[i:i + self.n_params]

    def sha256(self):
        return hashlib.sha256(self.table.tobytes()).hexdigest()


def draw_flags_and_indices(worker_rng, noise, eval_prob, batch_size):
    """Pre-draw what `batch_size` non-eval `collect_returns()` calls would draw
    (worker/worker.py:23,27 driven by run_sequential.py:134-147): one
    `uniform(0,1)` flag per call from the worker stream and, only for non-eval
    calls, one `randint` from the noise-source stream."""
    flags, idx = [], []
    while len(idx) < batch_size:
        is_eval = worker_rng.uniform(0, 1) < eval_prob
        flags.append(bool(is_eval))
        if not is_eval:
            idx.append(int(noise.sample()[0]))
    return flags, idx


# --------------------------------------------------------------------------
# a5  perturbation               reference: worker/worker.py:28
# --------------------------------------------------------------------------


def perturb(flat, sigma, eps, sign=1):
    """`new_flat = flat + sigma * eps` in fp32 with two roundings (product,
    then sum; numpy never fuses).  `sigma` is a Python float and becomes fp32
    against the fp32 array (NEP 50).  sign=-1 is the antithetic extension; the
    negation is exact so `flat + sigma * (-eps)` == `flat - sigma * eps`."""
    flat = np.asarray(flat, dtype=np.float32)
    eps = np.asarray(eps, dtype=np.float32)
    if sign < 0:
        eps = -eps
    return flat + sigma * eps


# --------------------------------------------------------------------------
# a6  flat layouts               reference: policies/policy.py:36-61
# --------------------------------------------------------------------------


@dataclass
class Entry:
    name: str
    shape: tuple
    kind: str            # "param" | "buffer"
    offset: int = 0      # params: offset in the trainable flat vector; buffers: offset in the buffer vector
    sd_offset: int = 0   # offset in the serialized state_dict (Policy.serialize order)

    @property
    def numel(self):
        return int(np.prod(self.shape)) if len(self.shape) else 1


@dataclass
class Layout:
    entries: list = field(default_factory=list)
    num_params: int = 0
    num_buffer: int = 0
    num_state: int = 0

    def add(self, name, shape, kind="param"):
        e = Entry(name, tuple(shape), kind)
        e.sd_offset = self.num_state
        if kind == "param":
            e.offset = self.num_params
            self.num_params += e.numel
        else:
            e.offset = self.num_buffer
            self.num_buffer += e.numel
        self.num_state += e.numel
        self.entries.append(e)
        return e

    def add_bn(self, prefix, c):
        """BatchNorm state_dict order: weight, bias, running_mean, running_var,
        num_batches_tracked (torch; the reference serialises all five,
        policy.py:44-49)."""
        self.add(prefix + ".weight", (c,))
        self.add(prefix + ".bias", (c,))
        self.add(prefix + ".running_mean", (c,), "buffer")
        self.add(prefix + ".running_var", (c,), "buffer")
        self.add(prefix + ".num_batches_tracked", (), "buffer")

    def add_wb(self, prefix, wshape):
        self.add(prefix + ".weight", wshape)
        self.add(prefix + ".bias", (wshape[0],))

    def get(self, name):
        for e in self.entries:
            if e.name == name:
                return e
        raise KeyError(name)

    def param(self, theta, name):
        e = self.get(name)
        return theta[e.offset:e.offset + e.numel].reshape(e.shape)

    def buffer(self, buffers, name):
        e = self.get(name)
        return buffers[e.offset:e.offset + e.numel].reshape(e.shape)

    def split_state(self, serialized):
        """Policy.serialize() list (policy.py:44-49) -> (theta, buffers)."""
        s = np.asarray(serialized, dtype=np.float32)
        assert s.shape[0] == self.num_state
        theta = np.empty(self.num_params, np.float32)
        buf = np.empty(self.num_buffer, np.float32)
        for e in self.entries:
            dst = theta if e.kind == "param" else buf
            dst[e.offset:e.offset + e.numel] = s[e.sd_offset:e.sd_offset + e.numel]
        return theta, buf

    def join_state(self, theta, buffers):
        s = np.empty(self.num_state, np.float32)
        for e in self.entries:
            src = theta if e.kind == "param" else buffers
            s[e.sd_offset:e.sd_offset + e.numel] = src[e.offset:e.offset + e.numel]
        return s


def mujoco_layout(n_in, n_act, h1=64, h2=64):
    """policies/mujoco.py:32-41 (h1=h2=64 hard-coded there; widths are
    parameters here for the 256x256 Humanoid config)."""
    L = Layout()
    L.add_wb("model.0", (h1, n_in))
    L.add_wb("model.2", (h2, h1))
    L.add_wb("model.4", (2 * n_act, h2))
    return L


def discrete_layout(n_in, n_act, h1=64, h2=64):
    """policies/discrete.py:34-48."""
    L = Layout()
    L.add_bn("model.0", n_in)
    L.add_wb("model.1", (h1, n_in))
    L.add_bn("model.3", h1)
    L.add_wb("model.4", (h2, h1))
    L.add_bn("model.6", h2)
    L.add_wb("model.7", (n_act, h2))
    return L


def atari_layout(n_act):
    """policies/atari.py:34-51 (the 2-conv DQN-2013 net)."""
    L = Layout()
    L.add_wb("model.0", (16, 4, 8, 8))
    L.add_bn("model.1", 16)
    L.add_wb("model.3", (32, 16, 4, 4))
    L.add_bn("model.4", 32)
    L.add_wb("model.7", (256, 2592))
    L.add_bn("model.8", 256)
    L.add_wb("model.10", (n_act, 256))
    return L


def impala_layout(n_act):
    """policies/impala.py:50-126.  Registration order (feat_convs, resnet1,
    resnet2, fc, core, policy: :109-122) is NOT execution order."""
    L = Layout()
    chans = [(3, 16), (16, 32), (32, 32)]
    for s, (cin, cout) in enumerate(chans):
        p = "model.0.feat_convs.%d" % s
        L.add_bn(p + ".0", cin)
        L.add_wb(p + ".1", (cout, cin, 3, 3))
    for blk in ("resnet1", "resnet2"):
        for s, (_, c) in enumerate(chans):
            p = "model.0.%s.%d" % (blk, s)
            L.add_bn(p + ".0", c)
            L.add_wb(p + ".2", (c, c, 3, 3))
            L.add_bn(p + ".3", c)
            L.add_wb(p + ".5", (c, c, 3, 3))
    L.add_bn("model.0.fc.0", 2048)
    L.add_wb("model.0.fc.1", (256, 2048))
    L.add("model.0.core.weight_ih_l0", (1024, 257))
    L.add("model.0.core.weight_hh_l0", (1024, 256))
    L.add("model.0.core.bias_ih_l0", (1024,))
    L.add("model.0.core.bias_hh_l0", (1024,))
    L.add_bn("model.0.policy.0", 256)
    L.add_wb("model.0.policy.1", (n_act, 256))
    return L


def synthetic_theta(layout, seed):
    """Seeded, layout-aware parameter vector used by golden fixtures for the
    big networks (so fixtures need not store millions of floats): weights
    N(0,1)/sqrt(fan_in), biases 0.1 N(0,1), BN gamma 1+0.1 N, BN beta 0.1 N."""
    rng = np.random.RandomState(seed)
    theta = np.empty(layout.num_params, np.float32)
    bn_prefixes = {e.name.rsplit(".", 1)[0] for e in layout.entries if e.name.endswith("running_mean")}
    for e in layout.entries:
        if e.kind != "param":
            continue
        pre, leaf = e.name.rsplit(".", 1)
        z = rng.randn(e.numel).astype(np.float32)
        if pre in bn_prefixes:
            v = 1.0 + 0.1 * z if leaf == "weight" else 0.1 * z
        elif len(e.shape) >= 2:
            fan_in = int(np.prod(e.shape[1:]))
            v = z / np.float32(math.sqrt(fan_in))
        else:
            v = 0.1 * z
        theta[e.offset:e.offset + e.numel] = v.astype(np.float32)
    return theta


def synthetic_buffers(layout, seed):
    """Seeded BN running stats: mean 0.1 N, var 1 + 0.1 |N|, nbt = 3."""
    rng = np.random.RandomState(seed)
    buf = np.empty(layout.num_buffer, np.float32)
    for e in layout.entries:
        if e.kind != "buffer":
            continue
        if e.name.endswith("running_mean"):
            v = 0.1 * rng.randn(e.numel)
        elif e.name.endswith("running_var"):
            v = 1.0 + 0.1 * np.abs(rng.randn(e.numel))
        else:
            v = np.full(e.numel, 3.0)
        buf[e.offset:e.offset + e.numel] = v.astype(np.float32)
    return buf


# --------------------------------------------------------------------------
# a7-a10  policy forwards (torch CPU fp32 functional restatement)
# --------------------------------------------------------------------------


def _t(x):
    import torch
    return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))


def _bn_eval(x, gamma, beta, mean, var, eps=1e-5):
    """Eval-mode BatchNorm over channel dim 1 (torch semantics; BN layers are in
    eval mode: discrete.py:12, atari.py:13)."""
    import torch
    shape = [1, -1] + [1] * (x.dim() - 2)
    return (x - mean.view(shape)) / torch.sqrt(var.view(shape) + eps) * gamma.view(shape) + beta.view(shape)


def mujoco_forward(layout, theta, obs):
    """policies/mujoco.py:35-41 + utils/torch_helpers.py:20-25.
    obs (E, K) -> mean (E, A), std (E, A)."""
    import torch
    th = _t(theta)
    x = _t(obs).view(-1, layout.get("model.0.weight").shape[1])
    P = lambda n: layout.param(th, n)
    x = torch.tanh(torch.nn.functional.linear(x, P("model.0.weight"), P("model.0.bias")))
    x = torch.tanh(torch.nn.functional.linear(x, P("model.2.weight"), P("model.2.bias")))
    x = torch.tanh(torch.nn.functional.linear(x, P("model.4.weight"), P("model.4.bias")))
    n = x.shape[-1] // 2
    return x[..., :n].numpy().copy(), (0.55 + 0.45 * x[..., n:]).numpy().copy()


def discrete_forward(layout, theta, buffers, obs):
    """policies/discrete.py:37-48: BN-Lin-ReLU-BN-Lin-ReLU-BN-Lin-Softmax,
    eval-mode BN with shared running stats and per-member gamma/beta."""
    import torch
    F = torch.nn.functional
    th, bf = _t(theta), _t(buffers)
    P = lambda n: layout.param(th, n)
    B = lambda n: layout.buffer(bf, n)
    x = _t(obs).view(-1, layout.get("model.1.weight").shape[1])
    for bn, lin, act in (("model.0", "model.1", True), ("model.3", "model.4", True), ("model.6", "model.7", False)):
        x = _bn_eval(x, P(bn + ".weight"), P(bn + ".bias"), B(bn + ".running_mean"), B(bn + ".running_var"))
        x = F.linear(x, P(lin + ".weight"), P(lin + ".bias"))
        if act:
            x = torch.relu(x)
    return torch.softmax(x, dim=-1).numpy().copy()


def atari_forward(layout, theta, buffers, obs_nchw):
    """policies/atari.py:35-51.  obs (E,4,84,84) float in [0,1]."""
    import torch
    F = torch.nn.functional
    th, bf = _t(theta), _t(buffers)
    P = lambda n: layout.param(th, n)
    B = lambda n: layout.buffer(bf, n)
    bn = lambda x, p: _bn_eval(x, P(p + ".weight"), P(p + ".bias"), B(p + ".running_mean"), B(p + ".running_var"))
    x = _t(obs_nchw).view(-1, 4, 84, 84)
    x = torch.relu(bn(F.conv2d(x, P("model.0.weight"), P("model.0.bias"), stride=4), "model.1"))
    x = torch.relu(bn(F.conv2d(x, P("model.3.weight"), P("model.3.bias"), stride=2), "model.4"))
    x = x.flatten(1)
    x = torch.relu(bn(F.linear(x, P("model.7.weight"), P("model.7.bias")), "model.8"))
    x = F.linear(x, P("model.10.weight"), P("model.10.bias"))
    return torch.softmax(x, dim=-1).numpy().copy()


def impala_forward(layout, theta, buffers, frame, reward, done, h, c):
    """policies/impala.py:136-186 for E independent single-step environments
    (reference call shape is E = 1: frame (1,1,3,64,64)).
    frame (E,3,64,64) 0..255; reward (E,); done (E,) bool; h,c (E,256).
    Returns probs (E,A), h' (E,256), c' (E,256)."""
    import torch
    F = torch.nn.functional
    th, bf = _t(theta), _t(buffers)
    P = lambda n: layout.param(th, n)
    B = lambda n: layout.buffer(bf, n)
    bn = lambda x, p: _bn_eval(x, P(p + ".weight"), P(p + ".bias"), B(p + ".running_mean"), B(p + ".running_var"))
    conv = lambda x, p: F.conv2d(x, P(p + ".weight"), P(p + ".bias"), stride=1, padding=1)
    x = _t(frame).view(-1, 3, 64, 64) / 255.0
    for s in range(3):
        p = "model.0.feat_convs.%d" % s
        x = F.max_pool2d(conv(bn(x, p + ".0"), p + ".1"), kernel_size=3, stride=2, padding=1)
        for blk in ("resnet1", "resnet2"):
            q = "model.0.%s.%d" % (blk, s)
            y = conv(torch.relu(bn(x, q + ".0")), q + ".2")
            y = conv(torch.relu(bn(y, q + ".3")), q + ".5")
            x = x + y
    x = torch.relu(x).flatten(1)
    x = torch.relu(F.linear(bn(x, "model.0.fc.0"), P("model.0.fc.1.weight"), P("model.0.fc.1.bias")))
    r = torch.clamp(_t(reward).view(-1, 1), -1, 1)
    core_in = torch.cat([x, r], dim=-1)
    nd = (~torch.from_numpy(np.asarray(done, dtype=bool))).float().view(-1, 1)
    h0, c0 = _t(h) * nd, _t(c) * nd
    gates = (F.linear(core_in, P("model.0.core.weight_ih_l0"), P("model.0.core.bias_ih_l0"))
             + F.linear(h0, P("model.0.core.weight_hh_l0"), P("model.0.core.bias_hh_l0")))
    i, f, g, o = gates.chunk(4, dim=-1)
    c1 = torch.sigmoid(f) * c0 + torch.sigmoid(i) * torch.tanh(g)
    h1 = torch.sigmoid(o) * torch.tanh(c1)
    logits = F.linear(bn(h1, "model.0.policy.0"), P("model.0.policy.1.weight"), P("model.0.policy.1.bias"))
    return torch.softmax(logits, dim=-1).numpy().copy(), h1.numpy().copy(), c1.numpy().copy()


# --------------------------------------------------------------------------
# a17  helpers                   reference: utils/math_helpers.py:127-144
# --------------------------------------------------------------------------


def standardize_arr(arr):
    x = np.asarray(arr)
    m = x.mean()
    s = x.std()
    if s == 0:
        return x
    return (x - m) / s


def affine_transform(value, from_min, from_max, to_min, to_max):
    if from_max == from_min or to_max == to_min:
        return to_min
    return (value - from_min) * (to_max - to_min) / (from_max - from_min) + to_min


# --------------------------------------------------------------------------
# a19  DSGD                      reference: dsgd/dynamic_sgd.py:18-51
# --------------------------------------------------------------------------


def dsgd_lr_scale(omega, min_omega, max_omega, min_scale=0.23, max_scale=1.0):
    """dynamic_sgd.py:41-44."""
    return affine_transform(omega, min_omega, max_omega, min_scale, max_scale)


def dsgd_step(theta, grad, lr, lr_scale):
    """dynamic_sgd.py:18-39 on the flat vector: fp32 grad, fp32 norm,
    theta -= (lr*sqrt(P)*lr_scale/norm) * grad, all in fp32 as torch does."""
    import torch
    g = torch.from_numpy(np.asarray(grad, dtype=np.float32))
    norm = g.norm().item()
    assert norm > 0
    coef = lr * np.sqrt(theta.shape[0]) * lr_scale / norm
    t = torch.from_numpy(np.array(theta, dtype=np.float32))
    t.sub_(coef * g)
    return t.numpy()


# --------------------------------------------------------------------------
# a13-a16  the estimator         reference: learner/finite_differences.py
# --------------------------------------------------------------------------


@dataclass
class Ret:
    """The FDReturn fields the estimator reads (learner/fd_return.py:5-16)."""
    epoch: int
    encoded_noise: str
    reward: float


class FiniteDifferencesOracle(object):
    """`FiniteDifferences` restated on flat vectors (finite_differences.py:6-114)
    with DSGD as the optimizer (the only one the drivers use)."""

    def __init__(self, theta0, noise, noise_std, lr, max_delayed_return=10,
                 omega=0.0, min_omega=0.0, max_omega=1.0):
        self.theta = np.array(theta0, dtype=np.float32)
        self.noise = noise
        self.noise_std = noise_std
        self.lr = lr
        self.omega, self.min_omega, self.max_omega = omega, min_omega, max_omega
        self.max_delayed_return = max_delayed_return
        self.policy_history = [(self.theta.copy(), 0)]       # :16
        self.epoch = 0
        self.discarded_returns = 0
        self.dist_map = {0: 0}                               # :19
        self.gradient_memory = np.zeros(self.theta.shape[0])

    def _rows(self, batch):
        """_process_returns + _adjust_return (:80-114)."""
        rewards, rows = [], []
        for ret in batch:
            if ret.epoch not in self.dist_map:               # :82-85
                self.discarded_returns += 1
                continue
            eps = self.noise.decode(ret.encoded_noise)       # :87
            lam = eps * self.noise_std + self.dist_map[ret.epoch]   # :89 fp32
            norm = np.linalg.norm(lam)                       # :107 fp32
            rewards.append(ret.reward)
            rows.append(lam / (norm * norm))                 # :112
        return rewards, rows

    def step(self, batch, policy_reward=None):
        rewards, rows = self._rows(batch)
        if policy_reward is None:
            policy_reward = 0
        if len(rewards) == 0:
            return 0
        w = standardize_arr(np.subtract(rewards, policy_reward))   # :40,43
        np.dot(w, rows, out=self.gradient_memory)            # :49 (the /len(batch) is discarded, G10)
        lr_scale = dsgd_lr_scale(self.omega, self.min_omega, self.max_omega)   # :51-52
        flat = self.theta
        grad32 = (-self.gradient_memory).astype(np.float32)  # policy.py:63-70
        self.theta = dsgd_step(flat, grad32, self.lr, lr_scale)
        update_size = np.linalg.norm(flat - self.theta)      # :59
        self.epoch += 1
        self.dist_map = {self.epoch: 0}                      # :66-73
        for params, e in self.policy_history:
            self.dist_map[e] = params - self.theta
        self.policy_history.append((self.theta.copy(), self.epoch))   # :75-78
        while len(self.policy_history) > self.max_delayed_return:
            self.policy_history.pop(0)
        return update_size


def fd_gradient_closed_form(table, idx, sign, rewards, sigma, n_params, baseline=0.0,
                            dist_rows=None, epoch_row=None):
    """fp64 closed form of the same estimator, used for full-size property
    tests where the python loop above would take minutes:
        g = sum_i w_i * lam_i / ||lam_i||^2,  lam_i = s_i*sigma*eps_i + d_{e_i}
    (reproduces gradient_memory to rel. err ~1e-7)."""
    w = standardize_arr(np.asarray(rewards, dtype=np.float64) - baseline)
    g = np.zeros(n_params)
    sig32 = np.float32(sigma)
    for i in range(len(idx)):
        eps = table[idx[i]:idx[i] + n_params]
        lam = (eps * sig32 * np.float32(sign[i])).astype(np.float32)
        if dist_rows is not None and epoch_row is not None and epoch_row[i] >= 0:
            lam = lam + dist_rows[epoch_row[i]]
        lam = lam.astype(np.float64)
        g += w[i] * lam / np.dot(lam, lam)
    return g


def fd_partial_gradient(table, idx, sign, rewards, all_rewards, sigma, n_params, baseline=0.0):
    """A rank's share of the estimator when the population is sharded: weights are standardised with
    the mean / std of ALL ranks' rewards, the sum runs over this rank's rows only.  Summing the partial
    gradients over ranks gives `fd_gradient_closed_form` of the whole batch (linearity of :49)."""
    allr = np.asarray(all_rewards, dtype=np.float64) - baseline
    m, s = allr.mean(), allr.std()
    x = np.asarray(rewards, dtype=np.float64) - baseline
    w = x if s == 0 else (x - m) / s
    g = np.zeros(n_params)
    sig32 = np.float32(sigma)
    for i in range(len(idx)):
        lam = (table[idx[i]:idx[i] + n_params] * sig32 * np.float32(sign[i])).astype(np.float64)
        g += w[i] * lam / np.dot(lam, lam)
    return g


# --------------------------------------------------------------------------
# Strategy distances / novelty / history
# --------------------------------------------------------------------------
def strategy_distance(name, a, b):
    """utils/math_helpers.py:166-222 restated (numpy, the input dtype's arithmetic): a (Z, W) or (n, Z, W) against
    b (m, Z, W) with numpy broadcasting; returns one distance per leading entry."""
    a, b = np.asarray(a), np.asarray(b)
    if name == "l2_dist":                                               # :166-170
        return np.linalg.norm(b - a, axis=-1).mean(axis=-1)
    if name == "categorical_tvd":                                       # :218-221
        return np.abs(np.subtract(a, b)).sum(axis=-1).mean(axis=-1)
    if name == "categorical_bhattacharrya_dist":                        # :194-197
        return (-np.log(np.sum(np.sqrt(a * b), axis=-1) + 1e-12)).mean(axis=-1)
    n = a.shape[-1] // 2
    m1, s1, m2, s2 = a[..., :n], a[..., n:], b[..., :n], b[..., n:]
    if name == "gaussian_wasserstein_dist_from_strategies":             # :200-216
        inside = s1 + s2 - 2 * np.sqrt(s1 * s2)
        return (np.square(np.linalg.norm(m1 - m2, axis=-1)) + inside.sum(axis=-1)).mean(axis=-1)
    if name == "gaussian_bhattacharrya_dist":                           # :173-191 (as written: no log on the "log term")
        s3 = (s1 + s2) / 2
        d = m1 - m2
        return ((d * d / s3).sum(axis=-1) / 8 + (s3.prod(axis=-1) / np.sqrt(s1.prod(axis=-1) * s2.prod(axis=-1))) / 4).mean(axis=-1)
    raise ValueError(name)


def strategy_novelty(name, strategy, others):
    """compute_strategy_novelty (math_helpers.py:147-155): distance to the nearest history strategy."""
    return float(np.min(strategy_distance(name, strategy, others)))


class StrategyHistoryOracle(object):
    """StrategyHandler + SparseHistoryManager + StrategyPoint restated (strategy/strategy_handler.py:6-30,
    sparse_history_manager.py:6-149, strategy_point.py:6-39).  `evaluate(flat, zeta)` is the policy's `get_strategy`
    with the given parameter vector (one of the forward restatements above)."""

    def __init__(self, evaluate, distance_name, max_history_size=200):
        self.evaluate, self.name, self.max = evaluate, distance_name, max_history_size
        self.flats, self.strategies = [], []
        self.closest, self.second = [], []
        self.strategy_tensor = np.zeros(0)
        self.zeta = None
        self.known = {}
        self.worst_point_idx = 0

    def add_policy(self, flat):
        flat = np.array(flat, dtype=np.float32)
        if len(self.flats) >= self.max and self.zeta is not None and len(self.zeta) > 0:   # manager :25-26
            return self._replace(flat)
        self.flats.append(flat)
        self.strategies.append(None)
        return None

    def set_zeta(self, zeta):
        if zeta is None or len(zeta) == 0:                               # handler :17-19
            return
        self.zeta = zeta
        self.strategies = [self.evaluate(f, zeta) for f in self.flats]   # manager :40-44
        n = len(self.flats)
        self.known = {(i, j): float(strategy_distance(self.name, self.strategies[i], self.strategies[j]))
                      for i in range(n) for j in range(i + 1, n)}       # :51-66
        self._update()
        self.strategy_tensor = np.asarray(self.strategies)

    def compute_novelty(self, flat):
        if self.zeta is None or len(self.zeta) == 0 or self.strategy_tensor is None or len(self.strategy_tensor) < 2:
            return 0                                                     # handler :26-27
        return strategy_novelty(self.name, self.evaluate(np.asarray(flat, np.float32), self.zeta), self.strategy_tensor)

    def _replace(self, flat):                                            # manager :72-109
        strategy = self.evaluate(flat, self.zeta)
        dists = strategy_distance(self.name, strategy, self.strategy_tensor)
        novelty = float(np.min(dists))
        idx = self.worst_point_idx
        current_worst = self.closest[idx][1]
        if novelty > current_worst or current_worst == np.inf:
            self.flats[idx] = flat
            self.strategies[idx] = strategy
            self.strategy_tensor[idx] = strategy
            for pair in self.known:
                if idx in pair:
                    self.known[pair] = float(dists[pair[1 - pair.index(idx)]])
            self._update()
            return idx
        return -1

    def _update(self):                                                   # manager :111-149, strategy_point.py:27-39
        n = len(self.flats)
        self.closest = [[None, np.inf] for _ in range(n)]
        self.second = [[None, np.inf] for _ in range(n)]
        for i in range(n):
            for key, val in self.known.items():
                if i in key:
                    if val < self.closest[i][1]:
                        self.second[i] = self.closest[i][:]
                        self.closest[i] = [key, val]
                    elif val < self.second[i][1] and key != self.closest[i][0]:
                        self.second[i] = [key, val]
        worst = np.inf
        for i in range(n):
            c = self.closest[i]
            if c[1] < worst:
                if c[0] is None:
                    self.worst_point_idx = i
                    continue
                j = c[0][1 - c[0].index(i)]
                worst = c[1]
                self.worst_point_idx = i if self.second[i][1] < self.second[j][1] else j
