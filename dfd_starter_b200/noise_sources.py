"""Noise sources with the reference's interface (utils/noise_sources.py).

`SharedNoiseTable(size, n_params, random_seed=123)`: `sample() -> (key, noise)`,
`decode(key) -> noise` exactly as utils/noise_sources.py:36-51.  The fp32 table
and every index come from numpy's legacy `RandomState` on the host (bit-exact
requirement; the same generator first fills the table, then serves the draws),
and the table is mirrored once per GPU as four 16-byte aligned shifted replicas
plus an fp64 prefix sum of squares (include/dfd_b200.h, dfd_table)."""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .device import get_context, ptr, aligned_ptr


class SharedNoiseTable(object):
    def __init__(self, size, n_params, random_seed=123, device=None, upload=None):
        assert size > n_params, "!ATTEMPTED TO MAKE NOISE TABLE WITH SIZE {} FOR {} PARAMETERS!".format(size, n_params)
        self._rng = np.random.RandomState(random_seed)
        self._table = self._rng.randn(size).astype(np.float32)
        self._n_params = n_params
        self._max_sample_idx = size - n_params
        self.size = size
        self._device = device
        self._dev = None
        if upload is None:
            upload = torch.cuda.is_available()
        if upload:
            self.to_device(device)

    # ---- reference interface -------------------------------------------------
    def sample(self):
        noise_idx = self._rng.randint(0, self._max_sample_idx)
        return "{}".format(noise_idx), self._table[noise_idx:noise_idx + self._n_params]

    def decode(self, noise_idx):
        """Plain decimal keys as in the reference.  '+i' / '-i' keys are the
        antithetic extension: '-i' decodes to -table[i:i+P]."""
        i, s = parse_key(noise_idx)
        v = self._table[i:i + self._n_params]
        return -v if s < 0 else v

    # ---- batched draws (same stream, same order as repeated sample()) ----------
    def sample_indices(self, n):
        """n successive `sample()` index draws as an int64 array."""
        return np.array([self._rng.randint(0, self._max_sample_idx) for _ in range(n)], dtype=np.int64)

    @property
    def n_params(self):
        return self._n_params

    # ---- device mirror -------------------------------------------------------
    def to_device(self, device=None):
        if self._dev is not None:
            return self._dev
        ctx = get_context(device if device is not None else self._device)
        lib = ctx.lib
        size = self.size
        stride = int(lib.dfd_table_replica_stride(size))
        with torch.cuda.device(ctx.device):
            raw = torch.from_numpy(self._table).to(ctx.device)
            replicas = torch.empty(4 * stride, dtype=torch.float32, device=ctx.device)
            prefix = torch.empty(size + 1, dtype=torch.float64, device=ctx.device)
            scratch = ctx.zeros_bytes(lib.dfd_table_scratch_bytes(size))
            _lib.check(lib.dfd_table_build(ctx.handle, ptr(raw), size, ptr(replicas), stride, ptr(prefix),
                                           aligned_ptr(scratch), scratch.numel() - 256, ctx.stream), "dfd_table_build")
            torch.cuda.current_stream(ctx.device).synchronize()
        del raw, scratch
        self._dev = DeviceTable(ctx, replicas, stride, prefix, size)
        return self._dev

    @property
    def device_table(self):
        return self.to_device()


class RNGNoiseSource(object):
    """utils/noise_sources.py:4-20 restated for numpy >= 2 (the reference reads `rng.__getstate__()['state']`, which
    newer numpy no longer lays out that way; the PCG64 words themselves are the same).  The key is the generator's
    `state,inc` BEFORE the draw, the noise is `standard_normal(P)` in fp64; `decode` rewinds the SAME generator to the
    key and redraws (so, as in the reference, it also moves the stream `sample()` continues from).
    Noise is generated on the host; the learner and the worker stage the vectors on the device per batch
    (`RowTable`).  In-kernel PCG64 + ziggurat generation is SURVEY.md §8(f) row N4."""

    def __init__(self, n_params, random_seed=123):
        self.rng = np.random.default_rng(np.random.SeedSequence(random_seed))
        self.n_params = n_params

    def sample(self):
        st = self.rng.bit_generator.state["state"]
        state = "{},{}".format(st["state"], st["inc"])
        noise = self.rng.standard_normal(size=self.n_params)
        return state, noise

    def decode(self, state):
        state_data = str(state).split(",")
        self.rng.bit_generator.state = {"bit_generator": "PCG64",
                                        "state": {"state": int(state_data[0]), "inc": int(state_data[1])},
                                        "has_uint32": 0, "uinteger": 0}
        return self.rng.standard_normal(size=self.n_params)


class SimpleNoiseSource(object):
    """utils/noise_sources.py:23-33: the key IS the noise vector."""

    def __init__(self, n_params, random_seed=123):
        self.rng = np.random.RandomState(random_seed)
        self.n_params = n_params

    def sample(self):
        noise = self.rng.randn(self.n_params)
        return noise, noise

    def decode(self, noise):
        return noise


class RowTable(object):
    """N host vectors staged as a throw-away device table: row j lives at entries [j*Ps, j*Ps + P) with Ps = P rounded
    up to 4, so every row starts 16-byte aligned in replica 0 and all kernels (forward, prepare, reduce, one-kernel
    step) run unchanged with idx[j] = j*Ps.  Used for noise sources that are not a shared table."""

    def __init__(self, ctx, rows):
        rows = np.ascontiguousarray(rows, dtype=np.float32)
        n, P = rows.shape
        self.Ps = (P + 3) // 4 * 4
        size = n * self.Ps + self.Ps + 8          # strictly larger than any row end; the tail is zero
        host = np.zeros(size, dtype=np.float32)
        host[:n * self.Ps].reshape(n, self.Ps)[:, :P] = rows
        lib = ctx.lib
        stride = int(lib.dfd_table_replica_stride(size))
        with torch.cuda.device(ctx.device):
            raw = torch.from_numpy(host).to(ctx.device)
            replicas = torch.empty(4 * stride, dtype=torch.float32, device=ctx.device)
            prefix = torch.empty(size + 1, dtype=torch.float64, device=ctx.device)
            scratch = ctx.zeros_bytes(lib.dfd_table_scratch_bytes(size))
            _lib.check(lib.dfd_table_build(ctx.handle, ptr(raw), size, ptr(replicas), stride, ptr(prefix),
                                           aligned_ptr(scratch), scratch.numel() - 256, ctx.stream), "dfd_table_build")
        self._keep = (raw, scratch)               # stream-ordered: freed with the object
        self.table = DeviceTable(ctx, replicas, stride, prefix, size)
        self.idx = np.arange(n, dtype=np.int64) * self.Ps

    def ref(self):
        return self.table.ref()


class DeviceTable(object):
    def __init__(self, ctx, replicas, stride, prefix, size):
        self.ctx, self.replicas, self.stride, self.prefix, self.size = ctx, replicas, stride, prefix, size
        self.c = _lib.DfdTable(replicas.data_ptr(), stride, prefix.data_ptr(), size)

    def ref(self):
        return C.byref(self.c)

    def ensure_scaled16(self, sigma, n_params):
        """Register the sigma-scaled fp16 mirror of this table with the context (include/dfd_b200.h:
        dfd_table_build_scaled16) unless the context already holds one for this table, this sigma and at least
        n_params parameters.  One-off cost: a pass over the table and 16 bytes per table entry; call it outside CUDA-graph
        capture (the first eager forward does)."""
        ctx = self.ctx
        sig = float(np.float32(sigma))
        have = getattr(ctx, "_scaled16", None)
        if have is not None and have[0] == self.replicas.data_ptr() and have[1] == sig and have[2] >= int(n_params):
            return
        lib = ctx.lib
        nbytes = int(lib.dfd_table_scaled16_bytes(self.size, int(n_params)))
        ctx._scaled16 = None
        _lib.check(lib.dfd_table_drop_scaled16(ctx.handle), "dfd_table_drop_scaled16")
        buf = torch.empty(nbytes + 256, dtype=torch.uint8, device=ctx.device)
        _lib.check(lib.dfd_table_build_scaled16(ctx.handle, self.ref(), sig, int(n_params), aligned_ptr(buf), nbytes,
                                                ctx.stream), "dfd_table_build_scaled16")
        ctx._scaled16 = (self.replicas.data_ptr(), sig, int(n_params), buf)


def parse_key(key):
    """'123' -> (123, +1); '+123' -> (123, +1); '-123' -> (123, -1)."""
    k = str(key)
    if k[0] == "-":
        return int(k[1:]), -1
    return int(k), 1
