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


# arguments on which glibc's two builds of log1p (plain / -mfma, csrc/rng_normal_core.h) round differently:
# (u, log1p(-u) of the plain build, log1p(-u) of the fused build)
_LOG1P_PROBES = (
    ("0x1.cea9fc931f76cp-3", "-0x1.0636846356c73p-2", "-0x1.0636846356c74p-2"),
    ("0x1.59f5ba805f9c0p-5", "-0x1.617a3c45aff1ap-5", "-0x1.617a3c45aff1bp-5"),
    ("0x1.3965b6984e77ep-2", "-0x1.76207f9db2642p-2", "-0x1.76207f9db2641p-2"),
    ("0x1.a840cb21997efp-1", "-0x1.c38cdbde6e899p+0", "-0x1.c38cdbde6e898p+0"),
)


# the same for exp: (x, exp(x) of the plain build, exp(x) of the fused build)
_EXP_PROBES = (
    ("-0x1.558f2e05ccc8ap-3", "0x1.b159cfd428588p-1", "0x1.b159cfd428587p-1"),
    ("-0x1.b1d883ba3443dp+1", "0x1.144d3b88dd75ap-5", "0x1.144d3b88dd759p-5"),
    ("-0x1.0371da37ef5a9p-3", "0x1.c3143d528ab2ap-1", "0x1.c3143d528ab2bp-1"),
    ("-0x1.a47b45f17bd8cp-1", "0x1.c26ff070c5ac8p-2", "0x1.c26ff070c5ac9p-2"),
)


def libm_fused():
    """Which builds of log1p and exp this process's libm resolved to (numpy's ziggurat calls them in the tail and in the
    wedge test): 1 = glibc's -mfma builds, 0 = the plain ones.  Anything else - or a mixed answer - is a libm the device
    restatement was not pinned against: raise rather than draw normals that might differ from numpy's."""
    import math
    got = [math.log1p(-float.fromhex(u)).hex() for u, _, _ in _LOG1P_PROBES] + \
          [math.exp(float.fromhex(x)).hex() for x, _, _ in _EXP_PROBES]
    if got == [f for _, _, f in _LOG1P_PROBES + _EXP_PROBES]:
        return 1
    if got == [p for _, p, _ in _LOG1P_PROBES + _EXP_PROBES]:
        return 0
    raise _lib.DfdError("this libm's log1p / exp match neither glibc build the device generator was pinned against "
                        "(%s): RNGNoiseSource(device=False) draws on the host" % ", ".join(got))


def pcg64_advance(state, inc, words):
    """State of numpy's PCG64 `words` 64-bit outputs after (state, inc)."""
    bg = np.random.PCG64()
    bg.state = {"bit_generator": "PCG64", "state": {"state": int(state), "inc": int(inc)}, "has_uint32": 0, "uinteger": 0}
    bg.advance(int(words))
    return int(bg.state["state"]["state"])


def device_normal_rows(ctx, streams, rows_per_stream, n_params, out=None, row_stride=None, theta=None, sigma=0.0,
                       dest_row=None, want_f64=False, force_serial=False, margin=0.0):
    """dfd_rng_normal_rows (include/dfd_b200.h): numpy's Generator(PCG64).standard_normal drawn on the device, bit-exact.
    streams: [n, 2] Python ints or an [n, 4] uint64 array {state_lo, state_hi, inc_lo, inc_hi}.
    Returns (rows fp32 device tensor [rows, row_stride] or None, rows_f64 or None, RowMarks, status).  Retries with a larger word budget on DFD_RNG_SHORT."""
    lib = ctx.lib
    if not (isinstance(streams, np.ndarray) and streams.dtype == np.uint64):
        m64 = (1 << 64) - 1
        streams = np.array([[int(s) & m64, int(s) >> 64, int(i) & m64, int(i) >> 64] for s, i in streams], dtype=np.uint64)
    streams = np.ascontiguousarray(streams.reshape(-1, 4))
    n = streams.shape[0]
    if n > 65535:                     # one launch serves up to 65 535 streams (grid.y): split, rows land in the same buffers
        assert out is not None and not want_f64, "more than 65 535 streams: pass the output buffer"
        marks = []
        for lo in range(0, n, 65535):
            hi = min(lo + 65535, n)
            dr = (np.arange(lo, hi) if dest_row is None else np.asarray(dest_row)[lo * int(rows_per_stream):hi * int(rows_per_stream)])
            if rows_per_stream != 1 and dest_row is None:
                dr = np.arange(lo * int(rows_per_stream), hi * int(rows_per_stream))
            _, _, m, status = device_normal_rows(ctx, streams[lo:hi], rows_per_stream, n_params, out=out, row_stride=row_stride,
                                                 theta=theta, sigma=sigma, dest_row=dr, force_serial=force_serial, margin=margin)
            marks.append(m)
        return out, None, RowMarks(np.concatenate([m.words for m in marks]), np.concatenate([m._states for m in marks])), status
    n_rows = n * int(rows_per_stream)
    row_stride = int(row_stride or n_params)
    fused = libm_fused()
    with torch.cuda.device(ctx.device):
        d_streams = torch.from_numpy(streams.view(np.int64)).to(ctx.device)
        if out is None:
            n_out = n_rows if dest_row is None else int(np.max(dest_row)) + 1
            out = torch.empty(n_out * row_stride, dtype=torch.float32, device=ctx.device)
        out64 = torch.empty(out.numel(), dtype=torch.float64, device=ctx.device) if want_f64 else None
        d_dest = None if dest_row is None else torch.from_numpy(np.ascontiguousarray(dest_row, dtype=np.int32)).to(ctx.device)
        d_words = torch.zeros(n * (int(rows_per_stream) + 1), dtype=torch.int64, device=ctx.device)
        d_states = torch.zeros(2 * n * (int(rows_per_stream) + 1), dtype=torch.int64, device=ctx.device)
        d_status = torch.zeros(1, dtype=torch.int32, device=ctx.device)
        for margin in ((margin or 1.04), 1.25, 2.0, 8.0):
            nbytes = int(lib.dfd_rng_scratch_bytes(n, int(rows_per_stream), int(n_params), float(margin)))
            scratch = torch.empty(nbytes + 256, dtype=torch.uint8, device=ctx.device)
            _lib.check(lib.dfd_rng_normal_rows(ctx.handle, ptr(d_streams), n, int(rows_per_stream), int(n_params), ptr(theta),
                                               float(sigma), ptr(d_dest), ptr(out), ptr(out64), row_stride, ptr(d_words),
                                               ptr(d_states), ptr(d_status), fused, 1 if force_serial else 0, float(margin),
                                               aligned_ptr(scratch), nbytes, ctx.stream), "dfd_rng_normal_rows")
            status = int(d_status.item()) & 0xffffffff
            del scratch
            if not status & 4:
                break
        else:
            raise _lib.DfdError("dfd_rng_normal_rows: word budget still short at 8 words per normal")
        if status & 2:
            raise _lib.DfdError("dfd_rng_normal_rows: a ziggurat tail loop ran past 60 rounds (status 0x%x)" % status)
        words = d_words.cpu().numpy().reshape(n, int(rows_per_stream) + 1)
        states = d_states.cpu().numpy().view(np.uint64).reshape(n, int(rows_per_stream) + 1, 2)
    return out, out64, RowMarks(words, states), status


class RowMarks(object):
    """Where the rows of a device draw begin in their streams: `words[s, r]` 64-bit words consumed, `state(s, r)` the
    PCG64 state there (row r's key; r = rows_per_stream: where the stream stands after the last row)."""

    def __init__(self, words, states):
        self.words, self._states = words, states

    def state(self, s, r):
        return int(self._states[s, r, 0]) | (int(self._states[s, r, 1]) << 64)

    def __getitem__(self, i):
        return self.words[i]


class LazyKeys(object):
    """The `encoded_noise` keys of a batch drawn on the device, held as arrays: streams uint64 [n, 4] = {state_lo, state_hi,
    inc_lo, inc_hi} of every member, is_key False for members without a draw (eval members: key "0", worker.py:34).
    Behaves like the list of key strings the reference carries (len, indexing, iteration) and formats a key only when
    somebody asks for it; the learner takes `streams` as they are (no 128-bit decimal round trip per return)."""

    def __init__(self, streams, is_key):
        self.streams = np.ascontiguousarray(streams, dtype=np.uint64).reshape(-1, 4)
        self.is_key = np.asarray(is_key, dtype=bool)

    def __len__(self):
        return len(self.is_key)

    def __getitem__(self, j):
        if isinstance(j, slice):
            return [self[i] for i in range(*j.indices(len(self)))]
        j = int(j)
        if not self.is_key[j]:
            return "0"
        a = self.streams[j]
        return "{},{}".format(int(a[0]) | (int(a[1]) << 64), int(a[2]) | (int(a[3]) << 64))

    def __iter__(self):
        return (self[j] for j in range(len(self)))

    def take(self, which):
        w = np.asarray(which)
        return LazyKeys(self.streams[w], self.is_key[w])


class RNGNoiseSource(object):
    """utils/noise_sources.py:4-20 restated for numpy >= 2 (the reference reads `rng.__getstate__()['state']`, which
    newer numpy no longer lays out that way; the PCG64 words themselves are the same).  The key is the generator's
    `state,inc` BEFORE the draw, the noise is `standard_normal(P)` in fp64; `decode` rewinds the SAME generator to the
    key and redraws (so, as in the reference, it also moves the stream `sample()` continues from).

    `sample` / `decode` are the reference's per-member host calls.  The batched worker and the learner do not call them:
    `sample_rows` / `decode_rows` draw the rows of a whole batch on the device (csrc/rng_normal.cu, bit-identical to
    numpy: SURVEY.md §8(f) row N4) and leave this object's generator where the per-member calls would have left it.
    device=False keeps the host path (`Worker` / `FiniteDifferences` then stage host-drawn rows)."""

    def __init__(self, n_params, random_seed=123, device=True):
        self.rng = np.random.default_rng(np.random.SeedSequence(random_seed))
        self.n_params = n_params
        self.device_rows = bool(device)

    def _state(self):
        st = self.rng.bit_generator.state["state"]
        return int(st["state"]), int(st["inc"])

    def _set_state(self, state, inc):
        self.rng.bit_generator.state = {"bit_generator": "PCG64", "state": {"state": int(state), "inc": int(inc)},
                                        "has_uint32": 0, "uinteger": 0}

    def sample(self):
        state = "{},{}".format(*self._state())
        noise = self.rng.standard_normal(size=self.n_params)
        return state, noise

    def decode(self, state):
        state_data = str(state).split(",")
        self._set_state(state_data[0], state_data[1])
        return self.rng.standard_normal(size=self.n_params)

    # ---- batched device forms ------------------------------------------------
    def sample_rows(self, ctx, n, out, row_stride, dest_row=None, theta=None, sigma=0.0, as_streams=False):
        """n successive `sample()` calls: returns their keys; row j goes to row dest_row[j] of `out` (fp32 device
        buffer, row_stride apart) as fp32(eps) or, with theta, as fp32(fp64(theta) + sigma * eps) (worker.py:28)."""
        s0, inc = self._state()
        _, _, marks, _ = device_normal_rows(ctx, [(s0, inc)], n, self.n_params, out=out, row_stride=row_stride, theta=theta,
                                            sigma=sigma, dest_row=dest_row)
        self._set_state(marks.state(0, n), inc)
        if as_streams:              # [n, 4] uint64: the keys as numbers (LazyKeys formats them on demand)
            m64 = (1 << 64) - 1
            out4 = np.empty((n, 4), dtype=np.uint64)
            out4[:, :2] = marks._states[0, :n]
            out4[:, 2], out4[:, 3] = inc & m64, inc >> 64
            return out4
        return ["{},{}".format(marks.state(0, r), inc) for r in range(n)]

    def decode_rows(self, ctx, keys, out, row_stride):
        """`decode(key)` for every key, in order: row j of `out` = fp32(noise_j)."""
        if hasattr(keys, "streams"):            # LazyKeys: the numbers themselves
            streams = keys.streams
            last_inc = int(streams[-1, 2]) | (int(streams[-1, 3]) << 64)
        else:
            streams = [tuple(int(v) for v in str(k).split(",")) for k in keys]
            last_inc = streams[-1][1]
        _, _, marks, _ = device_normal_rows(ctx, streams, 1, self.n_params, out=out, row_stride=row_stride)
        self._set_state(marks.state(len(streams) - 1, 1), last_inc)    # where the last decode() leaves the generator


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
    """N vectors staged as a throw-away device table: row j lives at entries [j*Ps, j*Ps + P) with Ps = P rounded
    up to 4, so every row starts 16-byte aligned in replica 0 and all kernels (forward, prepare, reduce, one-kernel
    step) run unchanged with idx[j] = j*Ps.  Used for noise sources that are not a shared table.
    rows: host array [n, P] - or shape=(n, P): `raw` is left zeroed for a device producer (RNGNoiseSource.sample_rows /
    decode_rows write rows at stride Ps), which then calls build()."""

    def __init__(self, ctx, rows=None, shape=None):
        self.ctx = ctx
        if rows is not None:
            rows = np.ascontiguousarray(rows, dtype=np.float32)
            n, P = rows.shape
        else:
            n, P = int(shape[0]), int(shape[1])
        self.n, self.P = n, P
        self.Ps = (P + 3) // 4 * 4
        self.size = n * self.Ps + self.Ps + 8     # strictly larger than any row end; the tail is zero
        with torch.cuda.device(ctx.device):
            if rows is not None:
                host = np.zeros(self.size, dtype=np.float32)
                host[:n * self.Ps].reshape(n, self.Ps)[:, :P] = rows
                self.raw = torch.from_numpy(host).to(ctx.device)
            else:
                self.raw = torch.zeros(self.size, dtype=torch.float32, device=ctx.device)
        self.idx = np.arange(n, dtype=np.int64) * self.Ps
        self.table = None
        if rows is not None:
            self.build()

    def build(self):
        ctx, lib, size = self.ctx, self.ctx.lib, self.size
        stride = int(lib.dfd_table_replica_stride(size))
        with torch.cuda.device(ctx.device):
            replicas = torch.empty(4 * stride, dtype=torch.float32, device=ctx.device)
            prefix = torch.empty(size + 1, dtype=torch.float64, device=ctx.device)
            scratch = ctx.zeros_bytes(lib.dfd_table_scratch_bytes(size))
            _lib.check(lib.dfd_table_build(ctx.handle, ptr(self.raw), size, ptr(replicas), stride, ptr(prefix),
                                           aligned_ptr(scratch), scratch.numel() - 256, ctx.stream), "dfd_table_build")
        self._keep = (self.raw, scratch)          # stream-ordered: freed with the object
        self.table = DeviceTable(ctx, replicas, stride, prefix, size)
        return self

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
