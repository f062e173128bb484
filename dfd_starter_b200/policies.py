"""Policies with the reference's `Policy` interface (policies/policy.py:8-153),
evaluated by the sm_100a kernels behind the C ABI.

A policy object owns the flat parameter vector theta (device, fp32), the shared
BatchNorm running statistics (device) and a layout table that maps the flat
vector to tensors in `parameters()` order.  Besides the
reference's single-observation methods it has the batched entry point

    forward_members(idx, sign, obs, sigma)  ->  [members, obs_per_member, out_width]

which evaluates member m with theta + sign[m]*sigma*table[idx[m]:idx[m]+P]
(worker/worker.py:28) without ever materialising the perturbed vector in HBM.
"""
import ctypes as C
import math

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .device import get_context, ptr, aligned_ptr

KIND = {"mujoco": 0, "discrete": 1, "atari": 2, "impala": 3}


# ----------------------------------------------------------------------------
# layouts: (name, shape, is_param) in state_dict order
# ----------------------------------------------------------------------------
class Layout(object):
    def __init__(self):
        self.entries = []      # dicts: name, shape, numel, param(bool), off (in theta or buffer vec), sd_off
        self.num_params = self.num_buffers = self.num_state = 0

    def _add(self, name, shape, param=True):
        n = int(np.prod(shape)) if len(shape) else 1
        e = dict(name=name, shape=tuple(shape), numel=n, param=param, sd_off=self.num_state,
                 off=self.num_params if param else self.num_buffers)
        if param:
            self.num_params += n
        else:
            self.num_buffers += n
        self.num_state += n
        self.entries.append(e)

    def dense(self, name, n_out, *in_shape):
        self._add(name + ".weight", (n_out,) + tuple(in_shape))
        self._add(name + ".bias", (n_out,))

    def norm(self, name, c):
        self._add(name + ".weight", (c,))
        self._add(name + ".bias", (c,))
        self._add(name + ".running_mean", (c,), False)
        self._add(name + ".running_var", (c,), False)
        self._add(name + ".num_batches_tracked", (), False)

    def split_state(self, serialized):
        s = np.asarray(serialized, dtype=np.float32)
        if s.shape[0] != self.num_state:
            raise ValueError("serialized state has %d values, layout needs %d" % (s.shape[0], self.num_state))
        theta = np.empty(self.num_params, np.float32)
        buf = np.empty(self.num_buffers, np.float32)
        for e in self.entries:
            (theta if e["param"] else buf)[e["off"]:e["off"] + e["numel"]] = s[e["sd_off"]:e["sd_off"] + e["numel"]]
        return theta, buf

    def join_state(self, theta, buf):
        s = np.empty(self.num_state, np.float32)
        for e in self.entries:
            s[e["sd_off"]:e["sd_off"] + e["numel"]] = (theta if e["param"] else buf)[e["off"]:e["off"] + e["numel"]]
        return s


def build_layout(kind, n_in, n_act, h1=64, h2=64):
    L = Layout()
    if kind == "mujoco":          # policies/mujoco.py:32-41
        L.dense("model.0", h1, n_in)
        L.dense("model.2", h2, h1)
        L.dense("model.4", 2 * n_act, h2)
    elif kind == "discrete":      # policies/discrete.py:34-48
        L.norm("model.0", n_in)
        L.dense("model.1", h1, n_in)
        L.norm("model.3", h1)
        L.dense("model.4", h2, h1)
        L.norm("model.6", h2)
        L.dense("model.7", n_act, h2)
    elif kind == "atari":         # policies/atari.py:34-51
        L.dense("model.0", 16, 4, 8, 8)
        L.norm("model.1", 16)
        L.dense("model.3", 32, 16, 4, 4)
        L.norm("model.4", 32)
        L.dense("model.7", 256, 2592)
        L.norm("model.8", 256)
        L.dense("model.10", n_act, 256)
    elif kind == "impala":        # policies/impala.py:56-126, registration order
        widths = ((3, 16), (16, 32), (32, 32))
        for s, (ci, co) in enumerate(widths):
            L.norm("model.0.feat_convs.%d.0" % s, ci)
            L.dense("model.0.feat_convs.%d.1" % s, co, ci, 3, 3)
        for blk in ("resnet1", "resnet2"):
            for s, (_, c) in enumerate(widths):
                base = "model.0.%s.%d" % (blk, s)
                L.norm(base + ".0", c)
                L.dense(base + ".2", c, c, 3, 3)
                L.norm(base + ".3", c)
                L.dense(base + ".5", c, c, 3, 3)
        L.norm("model.0.fc.0", 2048)
        L.dense("model.0.fc.1", 256, 2048)
        L._add("model.0.core.weight_ih_l0", (1024, 257))
        L._add("model.0.core.weight_hh_l0", (1024, 256))
        L._add("model.0.core.bias_ih_l0", (1024,))
        L._add("model.0.core.bias_hh_l0", (1024,))
        L.norm("model.0.policy.0", 256)
        L.dense("model.0.policy.1", n_act, 256)
    else:
        raise ValueError(kind)
    return L


# ----------------------------------------------------------------------------
# initial parameters: torch default init drawn in registration order, then the
# reference's normc re-initialisation of Sequential layers (policy.py:88-115)
# ----------------------------------------------------------------------------
def _construction_order(kind, layout):
    """Entries in the order the reference CONSTRUCTS the modules (= the order torch's default init consumes the global
    RNG).  For every policy but IMPALA that is the registration order.  ImpalaCNN builds stage by stage - feat_convs[s],
    then the stage's two residual blocks (impala.py:62-107) - and only afterwards wraps the three lists in ModuleLists
    (:109-111), so its draws run feat0, res1_0, res2_0, feat1, ... while parameters() lists feat0..2, res1_0..2, res2_0..2."""
    ents = layout.entries
    if kind != "impala":
        return list(ents)
    order, taken = [], set()
    for s in range(3):
        for blk in ("feat_convs", "resnet1", "resnet2"):
            prefix = "model.0.%s.%d." % (blk, s)
            for i, e in enumerate(ents):
                if e["name"].startswith(prefix):
                    order.append(e)
                    taken.add(i)
    order += [e for i, e in enumerate(ents) if i not in taken]      # fc, core, policy: constructed last, in this order
    return order


def _default_init(layout, kind=None):
    """Draw torch's default initialisation for every tensor, consuming the global torch RNG in the order the reference's
    constructor creates the modules (see _construction_order); values land at their state_dict offsets."""
    theta = np.zeros(layout.num_params, np.float32)
    buf = np.zeros(layout.num_buffers, np.float32)
    ents = _construction_order(kind, layout)
    i = 0
    while i < len(ents):
        e = ents[i]
        name = e["name"]
        if name.endswith("running_mean") or name.endswith("num_batches_tracked"):
            i += 1
            continue
        if name.endswith("running_var"):
            buf[e["off"]:e["off"] + e["numel"]] = 1.0
            i += 1
            continue
        nxt = ents[i + 2]["name"] if i + 2 < len(ents) else ""
        if name.endswith(".weight") and nxt.endswith("running_mean"):      # BatchNorm affine
            theta[e["off"]:e["off"] + e["numel"]] = 1.0
            i += 2
            continue
        if "core.weight_ih" in name:                                        # nn.LSTM: U(-1/sqrt(H), 1/sqrt(H)) x4
            k = 1.0 / math.sqrt(256)
            for j in range(4):
                ee = ents[i + j]
                t = torch.empty(ee["shape"]).uniform_(-k, k)
                theta[ee["off"]:ee["off"] + ee["numel"]] = t.numpy().ravel()
            i += 4
            continue
        shape = e["shape"]
        if len(shape) == 2:
            m = nn.Linear(shape[1], shape[0])
        else:
            m = nn.Conv2d(shape[1], shape[0], kernel_size=(shape[2], shape[3]))
        b = ents[i + 1]
        theta[e["off"]:e["off"] + e["numel"]] = m.weight.detach().numpy().ravel()
        theta[b["off"]:b["off"] + b["numel"]] = m.bias.detach().numpy().ravel()
        i += 2
    return theta, buf


def _normc(layout, theta, rng):
    """policy.py:88-115: for every layer of `self.model` that has a weight (Linear,
    Conv2d AND BatchNorm), w += (normc_sample - w), b += -b; the last such layer uses
    gain 0.01.  Applied in fp32 with the same two roundings."""
    groups = []
    for e in layout.entries:
        if e["param"] and e["name"].endswith(".weight"):
            groups.append(e)
    by_name = {e["name"]: e for e in layout.entries}
    for n, e in enumerate(groups):
        std = 0.01 if n == len(groups) - 1 else 1.0
        out = rng.randn(*e["shape"]).astype(np.float32)
        out *= std / np.sqrt(np.square(out).sum(axis=0, keepdims=True))
        w = theta[e["off"]:e["off"] + e["numel"]].reshape(e["shape"])
        w += (out - w)
        b = by_name[e["name"][:-6] + "bias"]
        bv = theta[b["off"]:b["off"] + b["numel"]]
        bv += -bv


def initial_parameters(kind, n_in, n_act, seed=124, h1=64, h2=64):
    """(theta, buffers) exactly as constructing the reference policy would leave them:
    torch default init from the global torch RNG, then normc from RandomState(seed)
    (IMPALA: normc finds no layer with a weight in `self.model`, policy.py:94-100)."""
    layout = build_layout(kind, n_in, n_act, h1, h2)
    theta, buf = _default_init(layout, kind)
    if kind != "impala":
        _normc(layout, theta, np.random.RandomState(seed))
    return theta, buf


class Policy(object):
    """Host-side mirror of policies/policy.py:8-153 backed by device state."""
    kind = None

    def __init__(self, n_inputs, n_actions, seed=124, h1=64, h2=64, device=None, precision=0):
        self.input_shape = n_inputs
        self.output_shape = n_actions
        self.rng = np.random.RandomState(seed)
        self.h1, self.h2 = h1, h2
        n_in = int(np.prod(n_inputs)) if self.kind in ("mujoco", "discrete") else 0
        self.layout = build_layout(self.kind, n_in, int(n_actions), h1, h2)
        self.num_params = self.layout.num_params
        self.desc = _lib.DfdPolicyDesc(KIND[self.kind], n_in, h1, h2, int(n_actions), int(precision))
        self.ctx = get_context(device)
        lib = self.ctx.lib
        assert lib.dfd_policy_num_params(C.byref(self.desc)) == self.num_params
        assert lib.dfd_policy_num_buffers(C.byref(self.desc)) == self.layout.num_buffers
        self.out_width = int(lib.dfd_policy_out_width(C.byref(self.desc)))
        theta, buf = initial_parameters(self.kind, n_in, int(n_actions), seed, h1, h2)
        dev = self.ctx.device
        self.theta = torch.from_numpy(theta).to(dev)
        self.buffers = torch.from_numpy(buf).to(dev) if self.layout.num_buffers else None
        self._one_idx = torch.zeros(1, dtype=torch.int64, device=dev)
        self._one_sign = torch.zeros(1, dtype=torch.int8, device=dev)
        self._table = None
        self._flat_param = None
        self._entry = {e["name"]: e for e in self.layout.entries}

    # ---- flat parameter access (policy.py:36-61) -------------------------------
    def get_trainable_flat(self):
        return self.theta.cpu().numpy()

    def set_trainable_flat(self, flat):
        self.theta.copy_(torch.as_tensor(np.asarray(flat), dtype=torch.float32), non_blocking=False)

    def serialize(self):
        buf = self.buffers.cpu().numpy() if self.buffers is not None else np.zeros(0, np.float32)
        return self.layout.join_state(self.get_trainable_flat(), buf).tolist()

    def deserialize(self, serialized_state_dict):
        theta, buf = self.layout.split_state(serialized_state_dict)
        self.set_trainable_flat(theta)
        if self.buffers is not None:
            self.buffers.copy_(torch.from_numpy(buf))

    def set_buffers(self, buf):
        self.buffers.copy_(torch.as_tensor(np.asarray(buf), dtype=torch.float32))

    def reset(self):
        pass

    # ---- optimizer-facing surface (policy.py:63-84): ONE flat parameter that aliases theta ----------
    def parameters(self):
        """What a stock torch optimizer is built from (`torch.optim.Adam(policy.parameters())`, as the drivers build
        theirs from the reference nn.Module): a single flat `nn.Parameter` sharing theta's device storage, so the
        optimizer's in-place update IS the update of the vector every kernel reads."""
        if self._flat_param is None:
            self._flat_param = nn.Parameter(self.theta, requires_grad=True)
        return [self._flat_param]

    def set_grad_from_flat(self, gradient):
        """policy.py:63-70: hand a flat gradient (host array or device tensor, any float dtype) to the optimizer; it is
        cast to fp32 and ACCUMULATED into .grad the way `p.backward(grad)` does."""
        p = self.parameters()[0]
        g = torch.as_tensor(gradient).to(device=self.theta.device, dtype=torch.float32).reshape(-1)
        if g.numel() != self.num_params:
            raise ValueError("set_grad_from_flat: %d values for %d parameters" % (g.numel(), self.num_params))
        p.grad = g.clone() if p.grad is None else p.grad.add_(g)

    def get_grad_as_flat(self):
        """policy.py:72-83 (float64 host vector; zeros where no gradient has been set)."""
        p = self._flat_param
        if p is None or p.grad is None:
            return np.zeros(self.num_params)
        return p.grad.double().cpu().numpy()

    # ---- train-mode pass of the UNPERTURBED policy (compute_vbn, policy.py:31-34) ---------------------
    def _tensor(self, name):
        e = self._entry[name]
        src = self.theta if e["param"] else self.buffers
        return src[e["off"]:e["off"] + e["numel"]].view(e["shape"])

    def _norm_train(self, x, name):
        """nn.BatchNorm*d.forward in training mode: normalise with the batch statistics and move the shared running
        statistics towards them (momentum 0.1, unbiased variance), in place in `self.buffers`."""
        y = nn.functional.batch_norm(x, self._tensor(name + ".running_mean"), self._tensor(name + ".running_var"),
                                     self._tensor(name + ".weight"), self._tensor(name + ".bias"), True, 0.1, 1e-5)
        self._tensor(name + ".num_batches_tracked").add_(1)
        return y

    def _dense(self, x, name):
        return nn.functional.linear(x, self._tensor(name + ".weight"), self._tensor(name + ".bias"))

    def _conv(self, x, name, stride=1, padding=0):
        return nn.functional.conv2d(x, self._tensor(name + ".weight"), self._tensor(name + ".bias"), stride=stride,
                                    padding=padding)

    # ---- batched forward ---------------------------------------------------------
    def bind_table(self, noise_source):
        """The table whose rows perturb the members (needed even for the unperturbed
        M=1 wrappers: the kernel signature always carries a table)."""
        if not hasattr(noise_source, "device_table"):       # RNG / Simple noise sources: nothing to bind, rows are staged per batch
            return self
        self._table = noise_source.device_table
        return self

    def forward_members(self, idx, sign, obs, sigma, out=None, theta=None, table=None):
        """idx int64 [M], sign int8 [M] (0 = unperturbed), obs float32 [M, E, ...] — device tensors.
        theta / table override the policy's own parameter vector and bound table (used for noise sources that are
        not a shared table: the members' perturbed vectors are staged as a RowTable and evaluated against theta = 0)."""
        table = table if table is not None else self._table
        theta = theta if theta is not None else self.theta
        if table is None:
            raise _lib.DfdError("policy.bind_table(noise_source) must be called before forward_members")
        M = idx.shape[0]
        E = obs.shape[1]
        if out is None:
            out = torch.empty(M, E, self.out_width, dtype=torch.float32, device=self.ctx.device)
        obs = obs.contiguous()
        if (self.desc.precision >= 1 and hasattr(table, "ensure_scaled16")
                and (self.kind == "atari" or (self.kind == "mujoco" and self.ctx.lib.dfd_policy_direct_supported(C.byref(self.desc))))
                and not torch.cuda.is_current_stream_capturing()):
            table.ensure_scaled16(sigma, self.num_params)       # wide MLPs: weight tiles straight from the table by TMA
        _lib.check(self.ctx.lib.dfd_policy_forward(
            self.ctx.handle, C.byref(self.desc), table.ref(), ptr(theta), ptr(self.buffers), ptr(idx),
            ptr(sign), M, float(sigma), ptr(obs), E, ptr(out), self.ctx.stream), "dfd_policy_forward")
        return out

    # ---- reference single-policy wrappers (M = 1, sign = 0) ----------------------
    def _obs_tensor(self, x):
        x = torch.as_tensor(np.asarray(x) if not torch.is_tensor(x) else x, dtype=torch.float32)
        return x.reshape((1, -1) + self._obs_shape()).to(self.ctx.device)

    def _obs_shape(self):
        return (int(np.prod(self.input_shape)),)

    def forward(self, x):
        return self.forward_members(self._one_idx, self._one_sign, self._obs_tensor(x), 0.0)[0]


class MujocoPolicy(Policy):
    """policies/mujoco.py:8-41 (hidden widths are parameters: SURVEY.md G3)."""
    kind = "mujoco"

    def forward(self, x):
        y = super().forward(x)
        a = self.output_shape
        return y[..., :a], y[..., a:]

    def get_action(self, x, deterministic=False):
        mean, std = self.forward(x)
        if deterministic:
            return mean.flatten().tolist()
        return torch.normal(mean, std).flatten().tolist()

    def get_entropy(self, x):
        _, std = self.forward(x)
        ent = 0.5 + 0.5 * math.log(2 * math.pi) + torch.log(std)
        return ent.sum(dim=-1).mean().item()

    def get_strategy(self, x):
        return super().forward(x).cpu().numpy()


class DiscretePolicy(Policy):
    """policies/discrete.py:8-48."""
    kind = "discrete"

    def get_action(self, x, deterministic=False):
        probs = self.forward(x)
        if deterministic:
            return probs.argmax().item()
        return torch.multinomial(probs.reshape(-1, probs.shape[-1]), 1).reshape(-1)[0].item()

    def get_entropy(self, x):
        p = self.forward(x)
        logp = torch.log(p.clamp_min(torch.finfo(torch.float32).tiny))
        return (-(p * logp).sum(-1)).mean().item()

    def get_strategy(self, x):
        return self.forward(x).cpu().numpy()

    @torch.no_grad()
    def compute_vbn(self, buffer):
        """policy.py:31-34: a train-mode forward of the UNPERTURBED policy refreshes the shared BN running statistics
        (momentum 0.1, unbiased variance).  Once per epoch, off the hot path: plain torch ops on the device over views
        of theta / the buffer vector."""
        x = torch.as_tensor(np.asarray(buffer) if not torch.is_tensor(buffer) else buffer, dtype=torch.float32)
        x = x.reshape(-1, int(np.prod(self.input_shape))).to(self.ctx.device)
        x = torch.relu(self._dense(self._norm_train(x, "model.0"), "model.1"))
        x = torch.relu(self._dense(self._norm_train(x, "model.3"), "model.4"))
        self._dense(self._norm_train(x, "model.6"), "model.7")


class AtariPolicy(DiscretePolicy):
    """policies/atari.py:7-51; n_inputs = (84, 84), observations NCHW (E,4,84,84)."""
    kind = "atari"

    def __init__(self, n_inputs, n_actions, seed=124, device=None, precision=0):
        super().__init__(n_inputs, n_actions, seed=seed, device=device, precision=precision)
        self.input_shape = (4, n_inputs[0], n_inputs[1])

    def _obs_shape(self):
        return (4, 84, 84)

    @torch.no_grad()
    def compute_vbn(self, buffer):
        """policy.py:31-34 over atari.py:35-51: conv0 -> BN1 -> ReLU -> conv3 -> BN4 -> ReLU -> flatten -> L7 -> BN8 in
        training mode on the unperturbed parameters; `buffer`: frames (N,4,84,84) in [0,1] (tensor or array)."""
        x = torch.as_tensor(np.asarray(buffer) if not torch.is_tensor(buffer) else buffer, dtype=torch.float32)
        x = x.reshape((-1,) + self._obs_shape()).to(self.ctx.device)
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
            x = torch.relu(self._norm_train(self._conv(x, "model.0", stride=4), "model.1"))
            x = torch.relu(self._norm_train(self._conv(x, "model.3", stride=2), "model.4"))
        self._norm_train(self._dense(x.flatten(1), "model.7"), "model.8")


class ImpalaPolicy(DiscretePolicy):
    """policies/impala.py:8-186: IMPALA ResNet trunk + FC + LSTM(257->256) + softmax head.
    n_inputs = (3, 64, 64).  The carried LSTM state lives on the device, one (h, c) pair of 256
    floats per (member, environment); `reset()` zeroes the M = 1 wrapper state (impala.py:29-30)."""
    kind = "impala"

    def __init__(self, n_inputs, n_actions, seed=124, device=None, precision=0):
        super().__init__(n_inputs, n_actions, seed=seed, device=device, precision=precision)
        self.input_shape = tuple(n_inputs)
        self.state = None
        self.reset()

    def _obs_shape(self):
        return (3, 64, 64)

    def reset(self):
        dev = self.ctx.device
        self.state = (torch.zeros(1, 1, 256, device=dev), torch.zeros(1, 1, 256, device=dev))

    def forward_members_impala(self, idx, sign, frame, reward, done, h, c, sigma):
        """idx/sign [M]; frame [M,E,3,64,64] float 0..255; reward [M,E] float; done [M,E] bool/uint8;
        h, c [M,E,256].  Returns probs [M,E,A], h', c' — E independent environments per member."""
        if self._table is None:
            raise _lib.DfdError("policy.bind_table(noise_source) must be called before forward_members_impala")
        M, E = frame.shape[0], frame.shape[1]
        dev = self.ctx.device
        frame = frame.contiguous().float()
        reward = reward.contiguous().float()
        done8 = done.contiguous().to(torch.uint8)
        h, c = h.contiguous(), c.contiguous()
        probs = torch.empty(M, E, self.out_width, device=dev)
        h1, c1 = torch.empty_like(h), torch.empty_like(c)
        if (self.desc.precision >= 2 and hasattr(self._table, "ensure_scaled16")
                and not torch.cuda.is_current_stream_capturing()):
            self._table.ensure_scaled16(sigma, self.num_params)     # tcgen05 trunk + TMA-fed dense tail
        _lib.check(self.ctx.lib.dfd_impala_forward(
            self.ctx.handle, C.byref(self.desc), self._table.ref(), ptr(self.theta), ptr(self.buffers), ptr(idx),
            ptr(sign), M, float(sigma), ptr(frame), ptr(reward), ptr(done8), ptr(h), ptr(c), E, ptr(probs), ptr(h1),
            ptr(c1), None, 0, self.ctx.stream), "dfd_impala_forward")
        return probs, h1, c1

    def forward(self, x):
        """Reference call shape: dict(frame (1,1,3,64,64) 0..255, reward (1,1), done (1,1)); one step of the
        carried state (impala.py:136-186)."""
        dev = self.ctx.device
        frame = torch.as_tensor(x["frame"]).float().reshape(1, 1, 3, 64, 64).to(dev)
        reward = torch.as_tensor(x["reward"]).float().reshape(1, 1).to(dev)
        done = torch.as_tensor(x["done"]).reshape(1, 1).to(dev)
        probs, h1, c1 = self.forward_members_impala(self._one_idx, self._one_sign, frame, reward, done,
                                                    self.state[0], self.state[1], 0.0)
        self.state = (h1, c1)
        return probs[0]

    @torch.no_grad()
    def compute_vbn(self, buffer):
        """impala.py:13-17: `buffer` is a list of the dicts the environment wrapper emits (frame (1,1,3,64,64) 0..255,
        reward (1,1), done (1,1)); they are stacked into one batch of N single-step sequences and run through the
        UNPERTURBED network in training mode (:136-186), which refreshes the running statistics of all 17 BatchNorm
        layers.  Like the reference, the pass also advances the carried LSTM state (by N steps, see below)."""
        dev = self.ctx.device
        frame = torch.cat([torch.as_tensor(b["frame"]).float().reshape(-1, 3, 64, 64) for b in buffer]).to(dev)
        reward = torch.cat([torch.as_tensor(b["reward"]).float().reshape(-1) for b in buffer]).to(dev)
        done = torch.cat([torch.as_tensor(b["done"]).reshape(-1) for b in buffer]).to(dev)
        n = frame.shape[0]
        x = frame / 255.0
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
            for s in range(3):
                fc = "model.0.feat_convs.%d" % s
                x = nn.functional.max_pool2d(self._conv(self._norm_train(x, fc + ".0"), fc + ".1", padding=1), 3, 2, 1)
                for blk in ("resnet1", "resnet2"):
                    b = "model.0.%s.%d" % (blk, s)
                    y = self._conv(torch.relu(self._norm_train(x, b + ".0")), b + ".2", padding=1)
                    y = self._conv(torch.relu(self._norm_train(y, b + ".3")), b + ".5", padding=1)
                    x = x + y
        x = torch.relu(x).reshape(n, -1)
        x = torch.relu(self._dense(self._norm_train(x, "model.0.fc.0"), "model.0.fc.1"))
        core_in = torch.cat([x, reward.clamp(-1, 1).reshape(n, 1)], dim=-1)
        w_ih, b_ih = self._tensor("model.0.core.weight_ih_l0"), self._tensor("model.0.core.bias_ih_l0")
        w_hh, b_hh = self._tensor("model.0.core.weight_hh_l0"), self._tensor("model.0.core.bias_hh_l0")
        # the stacked batch is (B = N, T = 1); the LSTM is batch_first and is handed inp.unsqueeze(0) = (1, N, 257)
        # (impala.py:166-173), so the N entries run through it as ONE sequence of N steps from the carried state, and
        # only the first entry's `done` flag is applied (zip over T = 1 items, :166)
        keep = 0.0 if bool(done.reshape(-1)[0]) else 1.0
        h = self.state[0].reshape(1, 256) * keep
        c = self.state[1].reshape(1, 256) * keep
        x_part = nn.functional.linear(core_in, w_ih, b_ih)
        outs = []
        for t in range(n):
            gi, gf, gg, go = (x_part[t:t + 1] + nn.functional.linear(h, w_hh, b_hh)).chunk(4, dim=-1)
            c = torch.sigmoid(gf) * c + torch.sigmoid(gi) * torch.tanh(gg)
            h = torch.sigmoid(go) * torch.tanh(c)
            outs.append(h)
        self._norm_train(torch.cat(outs), "model.0.policy.0")
        self.state = (h.reshape(1, 1, 256).contiguous(), c.reshape(1, 1, 256).contiguous())


POLICY_CLASSES = {"mujoco": MujocoPolicy, "discrete": DiscretePolicy, "atari": AtariPolicy, "impala": ImpalaPolicy}
