"""`FDReturn` record (learner/fd_return.py:5-23): what a worker hands the learner."""
import numpy as np


class FDReturn(object):
    def __init__(self):
        self.epoch = -1
        self.encoded_noise = "-1"
        self.perturbation = None
        self.reward = 0
        self.novelty = 0
        self.entropy = 0
        self.timesteps = 0
        self.is_eval = False
        self.eval_states = []
        self.obs_stats_update = []

    def serialize(self):
        return self.reward, self.novelty, self.entropy, self.timesteps, self.encoded_noise, self.perturbation, \
               self.epoch, self.is_eval, self.eval_states, self.obs_stats_update

    def deserialize(self, other):
        self.reward, self.novelty, self.entropy, self.timesteps, self.encoded_noise, self.perturbation, self.epoch, \
            self.is_eval, self.eval_states, self.obs_stats_update = other


class ReturnBatch(object):
    """What the batched `Worker.evaluate` hands back: a read-only SEQUENCE of `FDReturn` records
    (len / indexing / iteration build the reference's record objects on demand) that also keeps the
    structure-of-arrays form the device learner consumes (`epoch`, `idx`, `sign`, `reward`, ... numpy
    arrays).  `FiniteDifferences.step` takes the arrays directly when it is given the untouched batch
    and falls back to the per-record path (learner/finite_differences.py:94-114 semantics) as soon as
    any record object has been handed out, because callers may have edited it."""

    def __init__(self, epoch, idx, sign, reward, entropy, timesteps, is_eval, states=None, keys=None):
        n = len(idx)
        self.epoch = np.full(n, int(epoch), dtype=np.int64) if np.isscalar(epoch) else np.asarray(epoch, dtype=np.int64)
        self.idx = np.asarray(idx, dtype=np.int64)
        self.sign = np.asarray(sign, dtype=np.int8)
        self.reward = np.asarray(reward, dtype=np.float64)
        self.entropy = np.asarray(entropy, dtype=np.float64)
        self.timesteps = np.asarray(timesteps, dtype=np.int64)
        self.is_eval = np.asarray(is_eval, dtype=bool)
        self.states = states
        self.keys = keys            # explicit `encoded_noise` values (noise sources whose key is not a table index)
        self.antithetic = bool((self.sign < 0).any())
        self._records = None

    def __len__(self):
        return len(self.idx)

    def key(self, j):
        """`encoded_noise` of record j (utils/noise_sources.py:46; eval members carry "0", worker.py:34;
        antithetic extension: '+i' / '-i')."""
        if self.is_eval[j]:
            return "0"
        if self.keys is not None:
            return self.keys[j]
        if self.antithetic:
            return ("+%d" if self.sign[j] > 0 else "-%d") % self.idx[j]
        return "%d" % self.idx[j]

    def _materialise(self):
        if self._records is None:
            recs = []
            for j in range(len(self)):
                ret = FDReturn()
                ret.is_eval = bool(self.is_eval[j])
                ret.timesteps = int(self.timesteps[j])
                ret.encoded_noise = self.key(j)
                ret.reward = float(self.reward[j])
                ret.novelty = 0
                ret.entropy = float(self.entropy[j])
                ret.epoch = int(self.epoch[j])
                ret.obs_stats_update = []
                if ret.is_eval and self.states is not None:
                    ret.eval_states = self.states
                recs.append(ret)
            self._records = recs
        return self._records

    @property
    def soa(self):
        """(epoch, idx, sign, reward) when no record object has been handed out, else None."""
        if self._records is not None or self.keys is not None:
            return None
        return self.epoch, self.idx, self.sign, self.reward

    def __getitem__(self, j):
        return self._materialise()[j]

    def __iter__(self):
        return iter(self._materialise())
