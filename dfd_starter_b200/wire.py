"""proto3 wire codec of the worker <-> learner messages.

Schema: networking/rpc_misc/proto/client_server_interface.proto:13-49 (`Config`, `ServerState`, `Return`,
`ReturnArray`, `Null`).  The reference goes bytes -> generated pb2 message -> one `FDReturn` object per return
(learner/fd_return.py:41-56, networking/server.py:151-162); here `decode_returns` goes bytes -> `ReturnBatch`
(structure of arrays) through the C ABI (`dfd_wire_decode_returns`, csrc/wire_ingest.cu), and the encoders write
the exact bytes the reference's pb2 classes write (fields in number order, zero-valued scalars and empty repeated
fields omitted, repeated scalars packed), so either side of the RPC loop can be swapped independently.
`tests/golden/wire.npz` holds bytes produced by the reference's own classes.

The codec needs the shared library (host code only, no GPU); like the rest of the package it has no fallback
for a missing library.  The pure-Python decoder below serves only the encodings the C decoder declines
(unpacked repeated scalars, which no proto3 writer emits by default)."""
import ctypes as C
import struct

import numpy as np

from . import _lib
from .fd_return import ReturnBatch
from .fd_state import FDState


# ---------------------------------------------------------------- primitives
def _varint(v):
    v &= (1 << 64) - 1                       # negative int32 / int64 are sign-extended to 64 bits on the wire
    out = bytearray()
    while True:
        b = v & 0x7f
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _tag(field, wtype):
    return _varint(field << 3 | wtype)


def _f32_field(field, x):
    b = struct.pack("<f", x)
    return b"" if b == b"\0\0\0\0" else _tag(field, 5) + b        # -0.0 is written, +0.0 is the default and omitted


def _int_field(field, x):
    x = int(x)
    return b"" if x == 0 else _tag(field, 0) + _varint(x)


def _len_field(field, payload):
    return b"" if len(payload) == 0 else _tag(field, 2) + _varint(len(payload)) + payload


def _packed_f32(field, values):
    return _len_field(field, np.ravel(np.asarray(values, dtype=np.float64)).astype("<f4").tobytes())


def _packed_i32(field, values):
    return _len_field(field, b"".join(_varint(int(v)) for v in values))


def _read_varint(buf, p):
    v, shift = 0, 0
    while True:
        b = buf[p]
        p += 1
        v |= (b & 0x7f) << shift
        shift += 7
        if not b & 0x80:
            return v, p


def _signed(v, bits=64):
    v &= (1 << bits) - 1
    return v - (1 << bits) if v >> (bits - 1) else v


def _fields(buf):
    """(field, wire type, value) triples of one message; value is an int (varint), bytes (fixed / delimited)."""
    p, n = 0, len(buf)
    while p < n:
        tag, p = _read_varint(buf, p)
        f, t = tag >> 3, tag & 7
        if t == 0:
            v, p = _read_varint(buf, p)
        elif t == 1:
            v, p = buf[p:p + 8], p + 8
        elif t == 2:
            ln, p = _read_varint(buf, p)
            v, p = buf[p:p + ln], p + ln
            if len(v) != ln:
                raise ValueError("truncated proto3 message")
        elif t == 5:
            v, p = buf[p:p + 4], p + 4
        else:
            raise ValueError("unsupported wire type %d" % t)
        yield f, t, v
    if p != n:
        raise ValueError("truncated proto3 message")


def _unpack_varints(payload):
    out, p = [], 0
    while p < len(payload):
        v, p = _read_varint(payload, p)
        out.append(_signed(v, 64))
    return out


# ---------------------------------------------------------------- Return / ReturnArray
def encode_return(ret):
    """`FDReturn.serialize_to_grpc().SerializeToString()` (learner/fd_return.py:25-39), byte for byte."""
    states = np.asarray(ret.eval_states)
    return b"".join([
        _int_field(1, ret.epoch),
        _len_field(2, str(ret.encoded_noise).encode("utf-8")),
        _f32_field(3, ret.reward), _f32_field(4, ret.novelty), _f32_field(5, ret.entropy),
        _int_field(6, ret.timesteps),
        _int_field(7, 1 if ret.is_eval else 0),
        _packed_f32(8, np.ravel(states)),
        _packed_i32(9, np.shape(ret.eval_states)),
        _packed_f32(10, np.ravel(np.asarray(ret.obs_stats_update))),
    ])


def encode_return_array(rets):
    """`ReturnArray(rets=[...])` (networking/client.py:39-44)."""
    out = []
    for r in rets:
        b = encode_return(r)
        out.append(_tag(1, 2) + _varint(len(b)) + b)          # an empty sub-message is still written
    return b"".join(out)


def _eval_states(buf, so, sl, ho, hl):
    """learner/fd_return.py:52-56: reshape to eval_states_shape, fp32."""
    if sl == 0:
        return []
    vals = np.frombuffer(buf, dtype="<f4", count=sl // 4, offset=so)
    shape = _unpack_varints(bytes(buf[ho:ho + hl]))
    return vals.reshape(shape).astype(np.float32)


def decode_returns(buf, is_array=True):
    """Serialized `ReturnArray` (or one `Return`) -> `ReturnBatch`, in wire order."""
    buf = bytes(buf)
    lib = _lib.load()
    n = int(lib.dfd_wire_count_returns(buf, len(buf))) if is_array else 1
    if n < 0:
        raise _lib.DfdError("malformed ReturnArray")
    a = {"epoch": np.empty(n, np.int64), "idx": np.empty(n, np.int64), "sign": np.empty(n, np.int8),
         "reward": np.empty(n, np.float64), "novelty": np.empty(n, np.float32), "entropy": np.empty(n, np.float32),
         "timesteps": np.empty(n, np.int32), "is_eval": np.empty(n, np.uint8)}
    for k in ("key", "states", "shape", "stats"):
        a[k + "_off"] = np.empty(n, np.int64)
        a[k + "_len"] = np.empty(n, np.int32)
    soa = _lib.DfdReturnSoa(**{k: v.ctypes.data for k, v in a.items()})
    rc = int(lib.dfd_wire_decode_returns(buf, len(buf), 1 if is_array else 0, n, C.byref(soa)))
    if rc == -2:
        return _decode_returns_py(buf, is_array)
    if rc != n:
        raise _lib.DfdError("dfd_wire_decode_returns: %s" % lib.dfd_last_error().decode("utf-8", "replace"))
    is_eval = a["is_eval"].astype(bool)
    keys = None
    if bool(((a["idx"] < 0) & ~is_eval).any()):                   # keys that are not table indices (RNGNoiseSource ...)
        keys = [buf[o:o + l].decode("utf-8") for o, l in zip(a["key_off"], a["key_len"])]
    states = [None] * n
    for j in np.nonzero(is_eval)[0]:                             # fd_return.py:52: eval_states only read for eval returns
        states[j] = _eval_states(buf, int(a["states_off"][j]), int(a["states_len"][j]), int(a["shape_off"][j]),
                                 int(a["shape_len"][j]))
    stats = [np.frombuffer(buf, dtype="<f4", count=l // 4, offset=o) if l else [] for o, l in zip(a["stats_off"], a["stats_len"])] \
        if bool(a["stats_len"].any()) else None
    return ReturnBatch(a["epoch"], a["idx"], a["sign"], a["reward"], a["entropy"], a["timesteps"], is_eval, keys=keys,
                       novelty=a["novelty"], eval_states=states, obs_stats_updates=stats,
                       wire_keys=lambda j, b=buf, o=a["key_off"], l=a["key_len"]: b[o[j]:o[j] + l[j]].decode("utf-8"))


def _decode_one_py(msg):
    d = {"epoch": 0, "key": "", "reward": 0.0, "novelty": 0.0, "entropy": 0.0, "timesteps": 0, "is_eval": False,
         "states": [], "shape": [], "stats": []}
    rep = {8: ("states", "f"), 9: ("shape", "i"), 10: ("stats", "f")}
    for f, t, v in _fields(msg):
        if f == 1 and t == 0:
            d["epoch"] = _signed(v)
        elif f == 2 and t == 2:
            d["key"] = bytes(v).decode("utf-8")
        elif f in (3, 4, 5) and t == 5:
            d[{3: "reward", 4: "novelty", 5: "entropy"}[f]] = struct.unpack("<f", v)[0]
        elif f == 6 and t == 0:
            d["timesteps"] = _signed(v, 32)
        elif f == 7 and t == 0:
            d["is_eval"] = v != 0
        elif f in rep:
            name, kind = rep[f]
            if t == 2:
                d[name] += list(np.frombuffer(bytes(v), dtype="<f4")) if kind == "f" else _unpack_varints(bytes(v))
            else:
                d[name].append(struct.unpack("<f", v)[0] if kind == "f" else _signed(v, 64))
    return d


def _decode_returns_py(buf, is_array):
    from .noise_sources import parse_key
    msgs = [bytes(v) for f, t, v in _fields(buf) if f == 1 and t == 2] if is_array else [buf]
    ds = [_decode_one_py(m) for m in msgs]
    idx, sign, table_keys = [], [], True
    for d in ds:
        try:
            i, s = parse_key(d["key"])
        except ValueError:
            i, s, table_keys = -1, 0, table_keys and d["is_eval"]
        idx.append(i)
        sign.append(s)
    states = [(np.asarray(d["states"], np.float32).reshape(d["shape"]) if len(d["states"]) else []) if d["is_eval"] else None
              for d in ds]
    stats = [np.asarray(d["stats"], np.float32) for d in ds]
    return ReturnBatch([d["epoch"] for d in ds], idx, sign, [d["reward"] for d in ds], [d["entropy"] for d in ds],
                       [d["timesteps"] for d in ds], [d["is_eval"] for d in ds],
                       keys=None if table_keys else [d["key"] for d in ds],
                       novelty=[d["novelty"] for d in ds], eval_states=states, obs_stats_updates=stats,
                       wire_keys=lambda j, ks=[d["key"] for d in ds]: ks[j])


# ---------------------------------------------------------------- ServerState / Config
def encode_server_state(state):
    """`ServerState(...)` as networking/server.py:128-142 fills it from the learner's `FDState`
    (`policy_parameters` = `Policy.serialize()`, i.e. parameters AND BatchNorm buffers)."""
    def flat(x):
        return [] if x is None else np.ravel(np.asarray(x, dtype=np.float64))

    def shape(x, given):
        return list(given) if given is not None else ([] if x is None else list(np.shape(x)))
    return b"".join([
        _packed_f32(1, flat(state.strategy_history)),
        _packed_f32(2, flat(state.strategy_frames)),
        _packed_f32(3, flat(state.policy_params)),
        _packed_i32(4, shape(state.strategy_history, state.strategy_history_shape)),
        _packed_i32(5, shape(state.strategy_frames, state.strategy_frames_shape)),
        _int_field(6, state.epoch if state.epoch is not None else 0),
        _len_field(7, ("" if state.experiment_id is None else str(state.experiment_id)).encode("utf-8")),
        _packed_f32(8, flat(state.obs_stats)),
    ])


def decode_server_state(buf):
    """Bytes -> `FDState` with the fields networking/client.py:58-65 sets (arrays reshaped, params as an array)."""
    rep = {1: [], 2: [], 3: [], 4: [], 5: [], 8: []}
    st = FDState()
    st.epoch, st.experiment_id = 0, ""
    for f, t, v in _fields(bytes(buf)):
        if f in rep:
            if f in (4, 5):
                rep[f] += _unpack_varints(bytes(v)) if t == 2 else [_signed(v)]
            else:
                rep[f].append(np.frombuffer(bytes(v), dtype="<f4"))
        elif f == 6:
            st.epoch = _signed(v)
        elif f == 7:
            st.experiment_id = bytes(v).decode("utf-8")

    def cat(f):
        return np.concatenate(rep[f]) if rep[f] else np.zeros(0, np.float32)
    st.strategy_history_shape, st.strategy_frames_shape = rep[4], rep[5]
    st.strategy_history = np.reshape(cat(1), rep[4]) if rep[4] else cat(1)
    st.strategy_frames = np.reshape(cat(2), rep[5]) if rep[5] else cat(2)
    st.policy_params = cat(3)
    st.obs_stats = cat(8)
    return st


def encode_config(cfg):
    """`Config{params: google.protobuf.Struct}` (proto:13-16; networking/server.py:114-118)."""
    from google.protobuf import struct_pb2                    # protobuf runtime library (not reference code)
    s = struct_pb2.Struct()
    s.update(cfg)
    b = s.SerializeToString(deterministic=True)
    return _tag(1, 2) + _varint(len(b)) + b if cfg is not None else b""


def decode_config(buf):
    """Bytes -> dict, as `json_format.MessageToDict(cfg_msg)["params"]` (networking/client.py:49-50)."""
    from google.protobuf import struct_pb2, json_format
    s = struct_pb2.Struct()
    for f, t, v in _fields(bytes(buf)):
        if f == 1 and t == 2:
            s.MergeFromString(bytes(v))
    return json_format.MessageToDict(s)
