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

    def __init__(self, epoch, idx, sign, reward, entropy, timesteps, is_eval, states=None, keys=None, novelty=None,
                 eval_states=None, obs_stats_updates=None, wire_keys=None):
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
        self.novelty = None if novelty is None else np.asarray(novelty, dtype=np.float64)
        self.eval_states = eval_states              # per-return eval_states (ingested batches); `states` is shared
        self.obs_stats_updates = obs_stats_updates  # per-return obs_stats_update lists, or None
        self.wire_keys = wire_keys                  # j -> the key exactly as it arrived (ingested batches)
        self.antithetic = bool((self.sign < 0).any())
        self._records = None

    def __len__(self):
        return len(self.idx)

    def key(self, j):
        """`encoded_noise` of record j (utils/noise_sources.py:46; eval members carry "0", worker.py:34;
        antithetic extension: '+i' / '-i')."""
        if self.wire_keys is not None:
            return self.wire_keys(j)
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
                ret.novelty = 0 if self.novelty is None else float(self.novelty[j])
                ret.entropy = float(self.entropy[j])
                ret.epoch = int(self.epoch[j])
                ret.obs_stats_update = [] if self.obs_stats_updates is None else self.obs_stats_updates[j]
                if ret.is_eval and self.eval_states is not None:
                    if self.eval_states[j] is not None:
                        ret.eval_states = self.eval_states[j]
                elif ret.is_eval and self.states is not None:
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

    def select(self, which):
        """A new batch holding the returns `which` picks (boolean mask or index array), arrays only: the SoA
        counterpart of the drivers' `non_eval_returns.append(ret)` loop (run_server.py:143-158)."""
        w = np.asarray(which)
        w = np.nonzero(w)[0] if w.dtype == bool else w.astype(np.int64)

        def pick(x):
            return None if x is None else [x[j] for j in w]
        wk = self.wire_keys
        return ReturnBatch(self.epoch[w], self.idx[w], self.sign[w], self.reward[w], self.entropy[w], self.timesteps[w],
                           self.is_eval[w], states=self.states, keys=self.keys.take(w) if hasattr(self.keys, "take") else pick(self.keys),
                           novelty=None if self.novelty is None else self.novelty[w], eval_states=pick(self.eval_states),
                           obs_stats_updates=pick(self.obs_stats_updates),
                           wire_keys=None if wk is None else (lambda j, w=w, wk=wk: wk(int(w[j]))))

    @staticmethod
    def concat(batches):
        """One batch from several, in order (arrays only)."""
        batches = [b for b in batches if len(b)]
        if not batches:
            z = np.zeros(0)
            return ReturnBatch(z, z, z, z, z, z, z)
        if len(batches) == 1:
            return batches[0]

        def cat(name):
            return np.concatenate([getattr(b, name) for b in batches])

        def lists(name, fill):
            if all(getattr(b, name) is None for b in batches):
                return None
            out = []
            for b in batches:
                out += list(getattr(b, name)) if getattr(b, name) is not None else fill(b)
            return out
        wire_keys = None
        if any(b.wire_keys is not None for b in batches):
            owner = np.concatenate([np.full(len(b), i) for i, b in enumerate(batches)])
            local = np.concatenate([np.arange(len(b)) for b in batches])
            wire_keys = lambda j: batches[owner[j]].key(int(local[j]))            # noqa: E731
        return ReturnBatch(cat("epoch"), cat("idx"), cat("sign"), cat("reward"), cat("entropy"), cat("timesteps"), cat("is_eval"),
                           states=batches[0].states, keys=lists("keys", lambda b: [b.key(j) for j in range(len(b))]),
                           novelty=None if all(b.novelty is None for b in batches) else
                           np.concatenate([b.novelty if b.novelty is not None else np.zeros(len(b)) for b in batches]),
                           eval_states=lists("eval_states", lambda b: [b.states if e else None for e in b.is_eval]),
                           obs_stats_updates=lists("obs_stats_updates", lambda b: [[] for _ in range(len(b))]),
                           wire_keys=wire_keys)

    def non_eval(self):
        return self.select(~self.is_eval)

    def __getitem__(self, j):
        return self._materialise()[j]

    def __iter__(self):
        return iter(self._materialise())
