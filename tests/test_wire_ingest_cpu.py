"""CPU-only: ingestion of returns from the RPC loop against bytes and outcomes produced by
the reference's own classes (tests/golden/make_golden.py: gen_wire).  The decoder is host code behind the C ABI
(dfd_wire_decode_returns); no GPU is needed."""
import os
import threading

import numpy as np
import pytest

import __graft_entry__ as G

G.build()

from dfd_starter_b200 import wire                       # noqa: E402
from dfd_starter_b200.fd_return import FDReturn          # noqa: E402
from dfd_starter_b200.fd_state import FDState            # noqa: E402
from dfd_starter_b200.grpc_worker import GRPCWorker, ServerInterface, RPCClient   # noqa: E402


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "wire.npz"))


def _check_batch(b, g, sel=None):
    sel = np.arange(int(g["n"])) if sel is None else np.asarray(sel)
    assert len(b) == len(sel)
    assert np.array_equal(b.epoch, g["epoch"][sel])
    assert np.array_equal(b.reward, g["reward"][sel])              # fp32 on the wire, widened exactly
    assert np.array_equal(b.novelty, g["novelty"][sel])
    assert np.array_equal(b.entropy, g["entropy"][sel])
    assert np.array_equal(b.timesteps, g["timesteps"][sel])
    assert np.array_equal(b.is_eval, g["is_eval"][sel])
    for j, src in enumerate(sel):
        key = str(g["key"][src])
        assert b.key(j) == key
        if "," in key:
            assert b.idx[j] == -1 and b.sign[j] == 0
        else:
            assert b.idx[j] == int(key.lstrip("+-")) and b.sign[j] == (-1 if key[0] == "-" else 1)
        rec = b[j]
        assert rec.encoded_noise == key and rec.epoch == int(g["epoch"][src]) and rec.reward == float(g["reward"][src])
        assert rec.timesteps == int(g["timesteps"][src]) and rec.is_eval == bool(g["is_eval"][src])
        assert np.array_equal(np.asarray(rec.obs_stats_update, dtype=np.float32), g["ret%d_stats" % src])
        if rec.is_eval:
            want = g["ret%d_states" % src]
            got = np.asarray(rec.eval_states, dtype=np.float32)
            assert got.shape == want.shape and np.array_equal(got, want)


def test_decode_return_array_matches_reference_records(g):
    b = wire.decode_returns(g["array_bytes"].tobytes())
    _check_batch(b, g)
    assert b.soa is None                                 # RNG-style keys present: the learner decodes them on the host
    only_table = b.select(b.idx >= 0)
    assert only_table.keys is None or all("," not in k for k in only_table.keys)


def test_decode_single_returns_and_reencode_byte_exact(g):
    for j in range(int(g["n"])):
        raw = g["ret%d_bytes" % j].tobytes()
        b = wire.decode_returns(raw, is_array=False)
        _check_batch(b, g, [j])
        assert wire.encode_return(b[0]) == raw             # the encoder writes what the reference's pb2 class writes
    b = wire.decode_returns(g["array_bytes"].tobytes())
    assert wire.encode_return_array(list(b)) == g["array_bytes"].tobytes()


def test_general_decoder_on_unpacked_repeated_fields(g):
    """Unpacked repeated scalars are legal proto3 input: the C decoder declines (DFD_WIRE_UNSUPPORTED) and the general
    decoder gives the same records."""
    r = FDReturn()
    r.epoch, r.encoded_noise, r.reward, r.timesteps, r.is_eval = 4, "77", 1.5, 9, True
    r.eval_states = np.arange(6, dtype=np.float32).reshape(3, 2)
    packed = wire.decode_returns(wire.encode_return(r), is_array=False)
    import struct
    unpacked = b"".join([wire._int_field(1, 4), wire._len_field(2, b"77"), wire._f32_field(3, 1.5), wire._int_field(6, 9),
                         wire._int_field(7, 1)] +
                        [wire._tag(8, 5) + struct.pack("<f", v) for v in range(6)] +
                        [wire._tag(9, 0) + wire._varint(3), wire._tag(9, 0) + wire._varint(2)])
    lib = wire._lib.load()
    assert wire.decode_returns(unpacked, is_array=False).eval_states[0].shape == (3, 2)
    u = wire.decode_returns(unpacked, is_array=False)
    assert np.array_equal(u.eval_states[0], packed.eval_states[0]) and u.epoch[0] == 4 and u.idx[0] == 77
    assert lib.dfd_wire_count_returns(b"\x0a\x05abc", 5) == -1          # truncated
    with pytest.raises(wire._lib.DfdError):
        wire.decode_returns(b"\x0a\x05abc")
    assert len(wire.decode_returns(b"")) == 0


def test_server_state_and_config_bytes(g):
    st = FDState()
    st.strategy_frames, st.strategy_history = g["state_frames"], g["state_history"]
    st.policy_params = g["state_params"].tolist()
    st.epoch, st.experiment_id = 17, "exp-a1"
    st.obs_stats = g["state_obs_stats"].tolist()
    st.cfg = {"env_id": "Walker2d-v2", "noise_std": 0.02, "normalize_obs": True, "random_seed": 124, "eval_prob": 0.05}
    si = ServerInterface(st)
    assert si.state_bytes() == g["state_bytes"].tobytes()
    back = wire.decode_server_state(si.state_bytes())
    assert back.epoch == 17 and back.experiment_id == "exp-a1"
    assert np.array_equal(back.strategy_frames, g["state_frames"]) and np.array_equal(back.strategy_history, g["state_history"])
    assert np.array_equal(back.policy_params, g["state_params"]) and np.array_equal(back.obs_stats, g["state_obs_stats"])
    cfg = wire.decode_config(si.config_bytes())                   # server.py:145: the seed moves on with every GetConfig
    assert cfg == wire.decode_config(g["config_bytes"].tobytes())
    assert cfg["random_seed"] == 125 and cfg["env_id"] == "Walker2d-v2" and cfg["normalize_obs"] is True
    assert wire.decode_config(si.config_bytes())["random_seed"] == 126


def _state():
    st = FDState()
    st.policy_params, st.epoch, st.experiment_id, st.obs_stats, st.cfg = [0.5, 1.5], 3, "e", [], {"random_seed": 1}
    st.strategy_frames, st.strategy_history = np.zeros((2, 3), np.float32), np.ones((1, 2, 2), np.float32)   # run_server.py:99-100
    return st


def test_get_returns_batch_follows_the_reference_lifo_and_staleness_rules(g):
    """server.py:64-95 replayed: same arrivals, same three pulls (newest first, stale returns dropped and counted,
    eval returns handed over without counting, timesteps of everything popped)."""
    si = ServerInterface(_state())
    raw = [g["ret%d_bytes" % j].tobytes() for j in range(int(g["n"]))]
    # arrivals as a mix of single returns and arrays, as the two RPCs deliver them
    si.submit_batch(wire.decode_returns(raw[0], is_array=False))
    si.submit_batch(wire.decode_returns(wire.encode_return_array(list(wire.decode_returns(g["array_bytes"].tobytes()))[1:23])))
    for j in range(23, 26):
        si.submit_batch(wire.decode_returns(raw[j], is_array=False))
    si.submit_batch(wire.decode_returns(wire.encode_return_array(list(wire.decode_returns(g["array_bytes"].tobytes()))[26:])))
    order = {str(k) + "|%d" % t: None for k, t in zip(g["key"], g["timesteps"])}
    assert len(order) == int(g["n"])                              # (key, timesteps) identifies a return in this fixture
    ident = {str(k) + "|%d" % t: j for j, (k, t) in enumerate(zip(g["key"], g["timesteps"]))}
    for q, (bs, cur, mdr) in enumerate(g["pulls"]):
        rets, ts, n_del, n_disc = si.get_returns_batch(batch_size=int(bs), current_epoch=None if cur == -99 else int(cur),
                                                       max_delayed_return=None if mdr == -99 else int(mdr))
        ids = [ident["%s|%d" % (rets.key(j), rets.timesteps[j])] for j in range(len(rets))]
        assert ids == g["pull%d_ids" % q].tolist(), q
        assert [ts, n_del, n_disc] == g["pull%d_stats" % q].tolist(), q
    assert si.n_waiting() == int(g["left"])
    # nothing new arrives: a bounded wait gives back what there is (the reference would wait forever)
    rets, ts, _, _ = si.get_returns_batch(batch_size=50, timeout=0.05)
    assert len(rets) == int(g["left"]) and si.n_waiting() == 0


def test_grpc_loopback_submit_and_poll():
    """The service end to end on localhost: raw-bytes client -> gzip gRPC -> C decoder -> ReturnBatch; state polling flags
    (networking/client.py:67-88)."""
    w = GRPCWorker(_state())
    w.grpc_server.start(address="127.0.0.1", port=0)
    port = w.grpc_server.bound_port
    try:
        c = RPCClient()
        c.connect("127.0.0.1", port)
        assert c.get_server_state() == RPCClient.NEW_EXPERIMENT_FLAG
        assert c.current_state.epoch == 3 and c.current_state.cfg["random_seed"] == 2
        assert np.array_equal(c.current_state.policy_params, np.float32([0.5, 1.5]))
        assert c.get_server_state() == RPCClient.OPERATION_SUCCESSFUL_FLAG
        rets = []
        for j in range(10):
            r = FDReturn()
            r.epoch, r.encoded_noise, r.reward, r.timesteps, r.is_eval = 3, "%d" % (100 + j), float(j), 5, j == 4
            rets.append(r)

        def push():
            c.submit_returns(rets[:6])
            c.submit_return(rets[6])
            c.submit_returns(rets[7:])
        t = threading.Thread(target=push)
        t.start()
        got, ts, n_del, n_disc = w.collect_returns(batch_size=9, current_epoch=3, max_delayed_return=2, timeout=20)
        t.join()
        assert sorted(got.idx.tolist()) == list(range(100, 110)) and ts == 50 and n_del == 0 and n_disc == 0
        ne = got.non_eval()
        assert len(ne) == 9 and ne.soa is not None and 104 not in ne.idx.tolist()
        st = _state()
        st.epoch = 4
        w.update(st)
        assert c.get_server_state() == RPCClient.NEW_STATE_FLAG and c.current_state.epoch == 4
        c.disconnect()
    finally:
        w.grpc_server.stop(grace=0)


@pytest.mark.skipif(not os.path.isdir("/root/reference/networking"), reason="reference tree not present")
def test_interop_with_the_reference_client(g):
    """The UNMODIFIED reference `RPCClient` (networking/client.py, generated pb2 stubs) talking to this server."""
    import sys
    shim = os.path.join(os.path.dirname(__file__), "golden", "_shim")
    added = [p for p in (shim, "/root/reference") if p not in sys.path]
    sys.path[:0] = added
    try:
        from networking.client import RPCClient as RefClient
        from learner import FDReturn as RefReturn
        w = GRPCWorker(_state())
        w.grpc_server.start(address="127.0.0.1", port=0)
        try:
            c = RefClient()
            c.connect("127.0.0.1", w.grpc_server.bound_port)
            assert c.get_server_state() == RefClient.NEW_EXPERIMENT_FLAG
            assert c.current_state.epoch == 3 and c.current_state.cfg["random_seed"] == 2
            assert np.array_equal(np.float32(c.current_state.policy_params), np.float32([0.5, 1.5]))
            rets = []
            for j in range(5):
                r = RefReturn()
                r.epoch, r.encoded_noise, r.reward, r.timesteps = 3, "%d" % (7 + j), 0.25 * j, 11
                rets.append(r)
            c.submit_returns(rets)
            c.submit_return(rets[0])
            got, ts, _, _ = w.collect_returns(batch_size=6, current_epoch=3, max_delayed_return=2, timeout=20)
            assert got.idx.tolist() == [7, 11, 10, 9, 8, 7] and ts == 66
            assert got.reward.tolist() == [0.0, 1.0, 0.75, 0.5, 0.25, 0.0]
            c.disconnect()
        finally:
            w.grpc_server.stop(grace=0)
    finally:
        for p in added:
            sys.path.remove(p)
        for m in [m for m in sys.modules if m.split(".")[0] in ("networking", "learner", "utils", "gym", "dsgd", "policies", "worker")]:
            sys.modules.pop(m, None)
