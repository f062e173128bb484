"""`Worker` with the reference's interface (worker/worker.py:8-57), batched.

The reference evaluates one perturbed member at a time (`flat + sigma*eps` ->
`set_trainable_flat` -> rollout -> restore).  Here `collect_returns(n)` draws the
n eval flags and noise indices from the SAME two RandomState streams in the SAME
order (worker.py:23,27), then evaluates all n members in ONE batched call: the
perturbation is generated in-kernel from the table offsets and never written to
HBM.  Rollouts are delegated to a batched agent:

    agent.collect_returns(policy, idx, sign, sigma) -> dict(reward[n], entropy[n], timesteps[n], states)

(`SyntheticAgent` below evaluates synthetic observations; environment stepping
itself is outside this path)
"""
import math

import numpy as np
import torch

from .fd_return import FDReturn, ReturnBatch  # noqa: F401


class Worker(object):
    def __init__(self, policy, agent, noise_source, strategy_handler=None, sigma=0.02, eval_prob=0.1, random_seed=123):
        self.policy = policy
        self.agent = agent
        self.noise_source = noise_source
        self.strategy_handler = strategy_handler
        self.sigma = sigma
        self.epoch = -1
        self.rng = np.random.RandomState(random_seed)
        self.eval_prob = eval_prob
        from .obs_stats import WelfordRunningStat
        self.fixed_obs_stats = WelfordRunningStat(policy.input_shape)      # worker.py:17
        if hasattr(policy, "bind_table"):
            policy.bind_table(noise_source)

    def draw(self, n):
        """n calls' worth of (is_eval, idx): one uniform per call from the worker stream and,
        for non-eval calls only, one randint from the noise stream (worker.py:23-27)."""
        flags = np.zeros(n, dtype=bool)
        idx = np.zeros(n, dtype=np.int64)
        for i in range(n):
            flags[i] = self.rng.uniform(0, 1) < self.eval_prob
            if not flags[i]:
                idx[i] = int(self.noise_source.sample()[0])
        return flags, idx

    def draw_batch(self, batch_size):
        """The drivers' loop (run_sequential.py:134-147): keep calling until `batch_size`
        NON-eval members exist.  Returns flags/idx over all calls made."""
        flags, idx = [], []
        n_train = 0
        while n_train < batch_size:
            f, i = self.draw(1)
            flags.append(bool(f[0]))
            idx.append(int(i[0]))
            n_train += 0 if f[0] else 1
        return np.array(flags, dtype=bool), np.array(idx, dtype=np.int64)

    @torch.no_grad()
    def collect_returns(self, n=1, antithetic=False):
        if not hasattr(self.noise_source, "device_table"):
            return self._collect_returns_host_noise(n)
        flags, idx = self.draw(n)
        return self.evaluate(flags, idx, antithetic=antithetic)

    @torch.no_grad()
    def _collect_returns_host_noise(self, n):
        """`RNGNoiseSource` / `SimpleNoiseSource` (utils/noise_sources.py:4-33): the noise is generated on the host, so
        the members' perturbed vectors are formed exactly as the reference does it - `flat + sigma * eps` with fp64
        noise, cast to fp32 by `set_trainable_flat` (worker/worker.py:28-29, policies/policy.py:40-42) - staged on the
        device as a RowTable and evaluated in ONE batched launch against theta = 0 (0 + 1 * x == x bit for bit).
        Same two draws per member, in the same order, as the reference loop."""
        from .noise_sources import RowTable
        if getattr(self.noise_source, "device_rows", False):
            return self._collect_returns_device_rng(n)
        flat = np.asarray(self.policy.get_trainable_flat(), dtype=np.float32)
        rows = np.empty((n, flat.shape[0]), dtype=np.float32)
        keys, is_eval = [], np.zeros(n, dtype=bool)
        for j in range(n):
            is_eval[j] = self.rng.uniform(0, 1) < self.eval_prob
            if is_eval[j]:
                rows[j] = flat
                keys.append("0")
            else:
                key, eps = self.noise_source.sample()
                rows[j] = (flat + self.sigma * eps).astype(np.float32)
                keys.append(key)
        ctx = self.policy.ctx
        rt = RowTable(ctx, rows)
        view = _RowPolicyView(self.policy, rt, torch.zeros(flat.shape[0], dtype=torch.float32, device=ctx.device))
        res = self.agent.collect_returns(view, rt.idx, np.ones(n, dtype=np.int8), 1.0)
        return ReturnBatch(self.epoch, rt.idx, np.ones(n, dtype=np.int8), res["reward"], res["entropy"], res["timesteps"],
                           is_eval, states=res.get("states"), keys=keys)

    @torch.no_grad()
    def _collect_returns_device_rng(self, n):
        """`RNGNoiseSource` with the rows drawn on the device (csrc/rng_normal.cu): the same draws in the same order as
        the loop of worker/worker.py:19-38 - the eval coin of member j comes from the worker's own generator, the
        noise of the non-eval members from ONE pass over the noise source's PCG64 stream (their keys are the stream
        states at the row boundaries) - and the same members: fp32(fp64(flat) + sigma * eps) built in-kernel, eval
        members = flat.  Evaluated in one batched launch against theta = 0 like the host-drawn form."""
        from .noise_sources import RowTable
        ctx = self.policy.ctx
        theta = self.policy.theta
        P = int(theta.numel())
        is_eval = np.array([self.rng.uniform(0, 1) < self.eval_prob for _ in range(n)], dtype=bool)
        train = np.nonzero(~is_eval)[0]
        rt = RowTable(ctx, shape=(n, P))
        rows = rt.raw[:n * rt.Ps].view(n, rt.Ps)
        from .noise_sources import LazyKeys
        streams = np.zeros((n, 4), dtype=np.uint64)
        if len(train):
            streams[train] = self.noise_source.sample_rows(ctx, len(train), rt.raw, rt.Ps, dest_row=train, theta=theta,
                                                           sigma=float(self.sigma), as_streams=True)
        keys = LazyKeys(streams, ~is_eval)        # "state,inc" strings on demand; eval members: "0"
        if is_eval.any():
            rows[torch.from_numpy(np.nonzero(is_eval)[0]).to(ctx.device), :P] = theta
        rt.build()
        view = _RowPolicyView(self.policy, rt, torch.zeros(P, dtype=torch.float32, device=ctx.device))
        res = self.agent.collect_returns(view, rt.idx, np.ones(n, dtype=np.int8), 1.0)
        return ReturnBatch(self.epoch, rt.idx, np.ones(n, dtype=np.int8), res["reward"], res["entropy"], res["timesteps"],
                           is_eval, states=res.get("states"), keys=keys)

    @torch.no_grad()
    def evaluate(self, flags, idx, antithetic=False):
        """Evaluate the drawn members in one batched launch and wrap them as FDReturns.
        antithetic=True (extension): every non-eval draw yields a +/- pair
        with keys '+i' / '-i'; returns are ordered [evals..., plus..., minus...]."""
        flags = np.asarray(flags, dtype=bool)
        idx = np.asarray(idx, dtype=np.int64)
        ev = np.nonzero(flags)[0]
        tr = np.nonzero(~flags)[0]
        if antithetic:
            m_idx = np.concatenate([idx[ev], idx[tr], idx[tr]])
            m_sign = np.concatenate([np.zeros(len(ev)), np.ones(len(tr)), -np.ones(len(tr))]).astype(np.int8)
            is_eval = np.concatenate([np.ones(len(ev), bool), np.zeros(2 * len(tr), bool)])
        else:
            m_idx = idx.copy()
            m_sign = np.where(flags, 0, 1).astype(np.int8)
            is_eval = flags.copy()                                           # worker.py:34: eval key is "0"
        if getattr(self.agent, "normalize_obs", False):
            # worker.py:47-48: the rollout normalises with the learner-wide statistics of the last FDState
            self.agent.obs_mean, self.agent.obs_std = self.fixed_obs_stats.mean, self.fixed_obs_stats.std
        res = self.agent.collect_returns(self.policy, m_idx, m_sign, self.sigma)
        novelty = None
        if self.strategy_handler is not None and hasattr(self.strategy_handler, "compute_novelty_members"):
            # worker.py:53: the novelty of every member's (perturbed) policy, here in one batched evaluation
            novelty = self.strategy_handler.compute_novelty_members(m_idx, m_sign, self.sigma)
        # records are built lazily (ReturnBatch): the learner consumes the arrays, any other caller
        # still sees a sequence of FDReturn objects
        return ReturnBatch(self.epoch, m_idx, m_sign, res["reward"], res["entropy"], res["timesteps"], is_eval,
                           states=res.get("states"), novelty=novelty, obs_stats_updates=res.get("obs_stats_updates"))

    def update(self, state):
        """worker.py:40-43: load the learner's snapshot (flattened state_dict incl. BN buffers)."""
        self.policy.deserialize(state.policy_params)
        self.epoch = state.epoch
        if state.obs_stats is not None and len(state.obs_stats):
            self.fixed_obs_stats.deserialize(np.asarray(state.obs_stats))


class _RowPolicyView(object):
    """What a batched agent sees when the members' parameter vectors were staged as rows: the policy's attributes,
    with `forward_members` evaluating row j (theta = 0, sigma = 1) instead of theta + sigma * table[idx]."""

    def __init__(self, policy, row_table, zero_theta):
        self._policy, self._rows, self._zero = policy, row_table, zero_theta

    def __getattr__(self, name):
        return getattr(self._policy, name)

    def forward_members(self, idx, sign, obs, sigma, out=None):
        return self._policy.forward_members(idx, sign, obs, 1.0, out=out, theta=self._zero, table=self._rows)


class SyntheticAgent(object):
    """Evaluates every member on `obs_per_member` synthetic observations (no
    environment): reward = - mean squared distance of the policy head to a fixed
    target head, so a learner can be seen to improve.  Observations are resident on
    the device; `host_obs=True` re-uploads them from pinned host memory each call
    (the end-to-end path)."""

    def __init__(self, policy, obs_per_member, seed=0, shared_obs=True, host_obs=False, members_hint=1,
                 normalize_obs=False, obs_stats_update_chance=0.01):
        self.E = int(obs_per_member)
        # worker/agent.py:26-41: observations are normalised with the learner-wide statistics (set by the Worker before
        # every evaluation) and every member folds the observations it draws (probability obs_stats_update_chance, from
        # the agent's own RandomState stream) into its own statistics, which travel as FDReturn.obs_stats_update
        self.normalize_obs = bool(normalize_obs)
        self.obs_stats_update_chance = obs_stats_update_chance
        self.obs_mean, self.obs_std = None, None
        self.rng = np.random.RandomState(seed)
        g = torch.Generator().manual_seed(seed)
        shape = policy._obs_shape()
        self.shape = shape
        self.shared = shared_obs
        n_sets = 1 if shared_obs else members_hint
        self.obs_host = torch.randn((n_sets, self.E) + shape, generator=g).pin_memory() \
            if torch.cuda.is_available() else torch.randn((n_sets, self.E) + shape, generator=g)
        self.target = torch.tanh(torch.randn(policy.out_width, generator=g)) * 0.5
        self.host_obs = host_obs
        self.obs_dev = None
        self.saved_states = []

    def _obs(self, policy, M):
        dev = policy.ctx.device
        if self.obs_dev is None or self.host_obs:
            self.obs_dev = self.obs_host.to(dev, non_blocking=True)
            self.target_dev = self.target.to(dev)
        if self.obs_dev.shape[0] == M:
            return self.obs_dev
        return self.obs_dev[:1].expand((M, self.E) + self.shape).contiguous()

    def collect_returns(self, policy, idx, sign, sigma):
        dev = policy.ctx.device
        M = len(idx)
        idx_d = torch.from_numpy(np.ascontiguousarray(idx)).to(dev)
        sign_d = torch.from_numpy(np.ascontiguousarray(sign)).to(dev)
        obs = self._obs(policy, M)
        stats = None
        if self.normalize_obs:
            from .obs_stats import normalize_obs, member_obs_stats
            select = torch.from_numpy(self.rng.uniform(0, 1, size=(M, self.E)) < self.obs_stats_update_chance)
            stats = member_obs_stats(policy.ctx, obs, select)                    # raw observations (agent.py:38-39)
            if self.obs_mean is not None:
                obs = normalize_obs(policy.ctx, obs, self.obs_mean, self.obs_std)    # agent.py:40-41
        out = policy.forward_members(idx_d, sign_d, obs, sigma)
        err = (out - self.target_dev) ** 2
        reward = -err.mean(dim=(1, 2))
        if policy.kind == "mujoco":
            a = policy.output_shape
            ent = (0.5 + 0.5 * math.log(2 * math.pi) + torch.log(out[..., a:])).sum(-1).mean(-1)
        else:
            ent = -(out * torch.log(out.clamp_min(1e-30))).sum(-1).mean(-1)
        return {"reward": reward.double().cpu().numpy(), "entropy": ent.double().cpu().numpy(),
                "timesteps": np.full(M, self.E), "states": None,
                "obs_stats_updates": None if stats is None else [r.tolist() for r in stats.cpu().numpy()]}
