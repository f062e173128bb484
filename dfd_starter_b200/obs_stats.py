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
    def __init__(self, shape):
        self.ones = np.ones(shape=shape, dtype=np.float32)
        self.zeros = np.zeros(shape=shape, dtype=np.float32)
        self.running_mean = np.zeros(shape=shape, dtype=np.float32)
        self.running_variance = np.zeros(shape=shape, dtype=np.float32)
        self.count = 0
        self.shape = shape

    def increment(self, samples, num):
        if num > 1:
            for i in range(num):
                self.update(samples[i])
        else:
            self.update(samples)

    def update(self, sample):
        if type(sample) == dict:
            sample = sample["frame"]
        current_count = self.count
        self.count += 1
        delta = (sample - self.running_mean).reshape(self.running_mean.shape)
        delta_n = (delta / self.count).reshape(self.running_mean.shape)
        self.running_mean += delta_n
        self.running_variance += delta * delta_n * current_count

    def reset(self):
        self.__init__(self.shape)

    @property
    def mean(self):
        return self.zeros if self.count < 2 else self.running_mean

    @property
    def std(self):
        if self.count < 2:
            return self.ones
        var = self.running_variance / (self.count - 1)
        var = np.where(var == 0, 1.0, var)         # a constant feature normalises to zero instead of dividing by zero
        return np.sqrt(var)

    def increment_from_obs_stats_update(self, obs_stats_update):
        """math_helpers.py:68-87: pairwise merge of (mean, M2, count)."""
        if len(obs_stats_update) == 0:
            return
        n = int(np.prod(self.shape))
        other_mean = np.asarray(obs_stats_update[:n], dtype=np.float32).reshape(self.running_mean.shape)
        other_var = np.asarray(obs_stats_update[n:-1], dtype=np.float32).reshape(self.running_variance.shape)
        other_count = obs_stats_update[-1]
        if other_count == 0:
            return
        count = self.count + other_count
        mean_delta = other_mean - self.running_mean
        mean_delta_squared = mean_delta * mean_delta
        combined_mean = (self.count * self.running_mean + other_count * other_mean) / count
        combined_variance = self.running_variance + other_var + mean_delta_squared * self.count * other_count / count
        self.running_mean = combined_mean
        self.running_variance = combined_variance
        self.count = count

    def serialize(self):
        return self.running_mean.ravel().tolist() + self.running_variance.ravel().tolist() + [self.count]

    def deserialize(self, other):
        self.reset()
        n = int(np.prod(self.shape))
        self.running_mean = np.reshape(other[:n], self.shape)
        self.running_variance = np.reshape(other[n:-1], self.shape)
        self.count = other[-1]


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
