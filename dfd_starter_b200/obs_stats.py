"""Observation statistics and normalisation (second half).

`WelfordRunningStat` mirrors utils/math_helpers.py:7-105 on the host (it is 2K+1 numbers of scalar work: the
learner-wide merge `increment_from_obs_stats_update` and the (de)serialisation that travels in `FDState.obs_stats` /
`FDReturn.obs_stats_update`).  The per-observation work of the rollout loop (worker/agent.py:37-41) is batched on the
device: `normalize_obs` applies `clip((obs - mean) / std, -10, 10)` to all members' observations in one launch and
`member_obs_stats` folds every member's drawn observations into that member's own statistics - both bit-identical to
numpy on fp32 observations (csrc/obs_stats.cu)."""
import numpy as np
import torch

from . import _lib
from .device import ptr


class WelfordRunningStat(object):
    """Running (count, mean, sum of squared deviations) of observations behind the reference's interface
    (utils/math_helpers.py:7-101; attribute and method names are what workers, drivers and `FDState.obs_stats` use).

    Everything funnels into two primitives: `_absorb` folds raw observations in one at a time (Welford's recurrence,
    evaluated in the accumulator's dtype and in the reference's operation order, so fp32 results are bit-identical to
    its `update`), and `_merge` combines with another accumulator's moments (Chan's pairwise formula, the reference's
    `increment_from_obs_stats_update`).  `running_variance` keeps the reference's name although it holds M2, the sum of
    squared deviations - the variance is M2 / (count - 1)."""

    def __init__(self, shape):
        self.shape = shape
        self.count = 0
        self.running_mean = np.zeros(shape, dtype=np.float32)
        self.running_variance = np.zeros(shape, dtype=np.float32)

    # ---- primitives ------------------------------------------------------------------------------------------
    def _absorb(self, observations):
        mean, m2 = self.running_mean, self.running_variance
        for x in observations:
            seen = self.count
            self.count = seen + 1
            dev = np.reshape(x - mean, mean.shape)              # deviation from the mean BEFORE this observation
            shift = np.reshape(dev / self.count, mean.shape)
            mean += shift
            m2 += dev * shift * seen

    def _merge(self, mean_b, m2_b, n_b):
        if n_b == 0:
            return
        n_a, mean_a = self.count, self.running_mean
        n = n_a + n_b
        gap = mean_b - mean_a
        self.running_variance = self.running_variance + m2_b + gap * gap * n_a * n_b / n
        self.running_mean = (n_a * mean_a + n_b * mean_b) / n
        self.count = n

    def _split(self, flat):
        """[mean | M2 | count] (the `serialize` layout) -> the three parts, mean / M2 still flat."""
        k = int(np.prod(self.shape))
        return flat[:k], flat[k:-1], flat[-1]

    # ---- the reference's surface -----------------------------------------------------------------------------
    def update(self, sample):
        self._absorb((sample["frame"] if isinstance(sample, dict) else sample,))

    def increment(self, samples, num):
        self._absorb([samples[i] for i in range(num)] if num > 1 else (samples,))

    def reset(self):
        self.__init__(self.shape)

    @property
    def mean(self):
        return self.running_mean if self.count >= 2 else np.zeros(self.shape, dtype=np.float32)

    @property
    def std(self):
        if self.count < 2:
            return np.ones(self.shape, dtype=np.float32)
        var = self.running_variance / (self.count - 1)
        # a constant feature gets variance 1, so (x - mean) / std maps it to zero instead of dividing by zero
        return np.sqrt(np.where(var == 0, 1.0, var))

    def increment_from_obs_stats_update(self, obs_stats_update):
        """Fold in what a worker shipped as `FDReturn.obs_stats_update` (math_helpers.py:68-87)."""
        if len(obs_stats_update) == 0:
            return
        mean_b, m2_b, n_b = self._split(obs_stats_update)
        as_f32 = lambda v: np.asarray(v, dtype=np.float32).reshape(self.running_mean.shape)   # noqa: E731
        self._merge(as_f32(mean_b), as_f32(m2_b), n_b)

    def serialize(self):
        return self.running_mean.ravel().tolist() + self.running_variance.ravel().tolist() + [self.count]

    def deserialize(self, other):
        """Adopt serialized moments as they are: the arrays take the dtype numpy infers from `other` (float64 for the
        lists `FDState.obs_stats` carries - which is why a worker that loaded them normalises in fp64, worker.py:43)."""
        mean, m2, n = self._split(other)
        self.running_mean = np.reshape(mean, self.shape)
        self.running_variance = np.reshape(m2, self.shape)
        self.count = n


def normalize_obs(ctx, obs, mean, std, clip=10.0, out=None):
    """obs: device fp32 [..., K]; mean / std: [K] host arrays or device tensors -> clip((obs - mean) / std, -clip, clip).
    The arithmetic follows the statistics' dtype, as numpy's does: float32 statistics -> fp32 subtract / divide;
    float64 statistics (what a Worker holds after `update(state)`: worker.py:43) -> fp64, rounded to fp32 once."""
    def dev(x):
        t = x if torch.is_tensor(x) else torch.from_numpy(np.ascontiguousarray(np.asarray(x).ravel()))
        if t.dtype not in (torch.float32, torch.float64):
            t = t.double()
        return t.to(ctx.device).contiguous()
    obs = obs.contiguous()
    mean_d, std_d = dev(mean), dev(std)
    f64 = mean_d.dtype == torch.float64 or std_d.dtype == torch.float64
    if f64:
        mean_d, std_d = mean_d.double(), std_d.double()
    K = int(mean_d.numel())
    if obs.numel() % K or std_d.numel() != K:
        raise _lib.DfdError("normalize_obs: %d observation values / %d std values do not match %d features"
                            % (obs.numel(), std_d.numel(), K))
    out = torch.empty_like(obs) if out is None else out
    _lib.check(ctx.lib.dfd_normalize_obs(ctx.handle, ptr(obs), obs.numel() // K, K, ptr(mean_d), ptr(std_d), 1 if f64 else 0,
                                         float(clip), ptr(out), ctx.stream), "dfd_normalize_obs")
    return out


def member_obs_stats(ctx, obs, select):
    """obs: device fp32 [M, E, ...]; select: [M, E] bool / uint8 (which observations each member's agent drew for its
    statistics, agent.py:38) -> device fp32 [M, 2K+1] rows laid out like `WelfordRunningStat.serialize()`."""
    M, E = obs.shape[0], obs.shape[1]
    obs = obs.contiguous()
    K = obs.numel() // max(M * E, 1)
    sel = select.to(device=ctx.device, dtype=torch.uint8).contiguous()
    out = torch.empty(M, 2 * K + 1, dtype=torch.float32, device=ctx.device)
    _lib.check(ctx.lib.dfd_member_obs_stats(ctx.handle, ptr(obs), ptr(sel), M, E, K, ptr(out), ctx.stream),
               "dfd_member_obs_stats")
    return out
