"""Learner-side ingestion of returns from the RPC loop, array-first.

Mirrors `GRPCWorker` (worker/grpc_worker.py:6-21) over `RPCServer` / `ServerInterface` / `ServerServicer`
(networking/server.py:9-162): same method names, arguments, LIFO / staleness rules and return tuple, so
`run_server.py:110-201` drives it unchanged.  What differs is the data path: `SubmitReturns` payloads are decoded
by the C ABI straight into structure-of-arrays chunks (`wire.decode_returns`), `collect_returns` hands back one
`ReturnBatch` (a sequence of `FDReturn` for the driver's loop, arrays for the device learner) and the `ServerState`
reply is serialized once per `update`, not once per client poll.  The gRPC service is registered from raw method
handlers (no generated stubs); the transport itself is grpc's and is not part of the measured path.

    worker = GRPCWorker(state); worker.start("localhost", 1025)
    rets, timesteps, n_delayed, n_discarded = worker.collect_returns(batch_size, learner.epoch, max_delayed_return)
    learner.step(rets.non_eval(), policy_reward, policy_novelty, policy_entropy)
"""
import threading
import time

import numpy as np

from . import wire
from .fd_return import ReturnBatch
from .fd_state import FDState

MAX_MESSAGE_LENGTH = 1 * (1024 ** 3)            # networking/server.py:10
SERVICE = "CSInterface"                         # proto:5 (no package => "/CSInterface/<Method>")


class ServerInterface(object):
    """networking/server.py:41-126."""

    def __init__(self, initial_state):
        self.cfg = None
        self.epoch = -1
        self.server_state = FDState()
        self._state_bytes = b""
        self._lock = threading.Lock()
        self.waiting_returns = []                # chunks (ReturnBatch), arrival order; consumed from the END (LIFO, :80)
        self.update(initial_state)

    # ---- producers (servicer threads) ----------------------------------------------------------
    def submit_return(self, ret):
        """One `FDReturn`-like record (kept for in-process callers, server.py:52-62)."""
        self.submit_batch(wire.decode_returns(wire.encode_return(ret), is_array=False))

    def submit_batch(self, batch):
        if len(batch):
            with self._lock:
                self.waiting_returns.append(batch)

    def n_waiting(self):
        with self._lock:
            return sum(len(c) for c in self.waiting_returns)

    # ---- consumer (the learner's thread) ---------------------------------------------------------
    def get_returns_batch(self, batch_size=None, current_epoch=None, max_delayed_return=None, timeout=None):
        """server.py:64-95 on arrays.  Pops the newest return first; every popped return's timesteps count; a return
        more than `max_delayed_return` epochs old is dropped and counted; eval returns are handed over but do not count
        towards `batch_size`; blocks (10 ms naps) until `batch_size` non-eval returns were taken.
        `timeout` (seconds, extension) bounds the wait; None waits forever like the reference."""
        timesteps = n_delayed = n_discarded = n_collected = 0
        taken = []
        if batch_size is None:                   # "every waiting return" (:71-73)
            batch_size = max(self.n_waiting(), 1)
        t_end = None if timeout is None else time.monotonic() + timeout
        while n_collected < batch_size:
            with self._lock:
                chunk = self.waiting_returns.pop() if self.waiting_returns else None
            if chunk is None:
                if t_end is not None and time.monotonic() > t_end:
                    break
                time.sleep(0.01)
                continue
            n = len(chunk)
            order = np.arange(n - 1, -1, -1)                     # pop(-1): newest first
            epoch = chunk.epoch[order]
            dropped = np.zeros(n, dtype=bool)
            delayed = np.zeros(n, dtype=bool)
            if current_epoch is not None:
                diff = int(current_epoch) - epoch
                if max_delayed_return is not None:
                    dropped = (diff > 0) & (diff > int(max_delayed_return))
                delayed = (diff > 0) & ~dropped
            counts = ~dropped & ~chunk.is_eval[order]
            cum = np.cumsum(counts)
            need = batch_size - n_collected
            k = n if cum[-1] < need else int(np.searchsorted(cum, need)) + 1     # returns popped from this chunk
            if k < n:                                            # the older part of the chunk keeps waiting
                with self._lock:
                    self.waiting_returns.append(chunk.select(np.arange(0, n - k)))
            timesteps += int(chunk.timesteps[order[:k]].sum())
            n_discarded += int(dropped[:k].sum())
            n_delayed += int(delayed[:k].sum())
            n_collected += int(cum[k - 1])
            taken.append(chunk.select(order[:k][~dropped[:k]]))
        return ReturnBatch.concat(taken), timesteps, n_delayed, n_discarded

    # ---- state going down to the workers -----------------------------------------------------------
    def update(self, server_state):
        """server.py:97-112: snapshot the learner's `FDState`; the config is re-read only for a new experiment."""
        st = self.server_state
        st.epoch = server_state.epoch
        st.policy_params = server_state.policy_params
        st.strategy_frames = np.ravel(server_state.strategy_frames).tolist() if server_state.strategy_frames is not None else []
        st.strategy_frames_shape = np.shape(server_state.strategy_frames) if server_state.strategy_frames is not None else ()
        st.strategy_history = np.ravel(server_state.strategy_history).tolist() if server_state.strategy_history is not None else []
        st.strategy_history_shape = np.shape(server_state.strategy_history) if server_state.strategy_history is not None else ()
        st.obs_stats = server_state.obs_stats
        if server_state.experiment_id != st.experiment_id:
            self.cfg = dict(server_state.cfg) if server_state.cfg is not None else {}
        st.experiment_id = server_state.experiment_id
        self.epoch = st.epoch
        self._state_bytes = wire.encode_server_state(st)         # serialized once per update, served to every poll

    def state_bytes(self):
        return self._state_bytes

    def config_bytes(self):
        """server.py:144-149: every `GetConfig` hands out the next `random_seed`."""
        with self._lock:
            if "random_seed" in self.cfg:
                self.cfg["random_seed"] += 1
            return wire.encode_config(self.cfg)

    def cleanup(self):
        self.server_state.cleanup()
        self.waiting_returns = []


class RPCServer(object):
    """networking/server.py:9-38, registered from raw-bytes method handlers."""
    MAX_MESSAGE_LENGTH = MAX_MESSAGE_LENGTH

    def __init__(self, initial_state):
        self.server_interface = ServerInterface(initial_state)
        self.grpc_server = None

    def update(self, server_state):
        self.server_interface.update(server_state)

    def get_returns_batch(self, batch_size=None, current_epoch=None, max_delayed_return=None, timeout=None):
        return self.server_interface.get_returns_batch(batch_size=batch_size, current_epoch=current_epoch,
                                                       max_delayed_return=max_delayed_return, timeout=timeout)

    def start(self, max_workers=10, address="localhost", port=50051):
        import grpc
        from concurrent import futures
        si = self.server_interface

        def submit_returns(request, context):
            si.submit_batch(wire.decode_returns(request, is_array=True))
            return b""

        def submit_return(request, context):
            si.submit_batch(wire.decode_returns(request, is_array=False))
            return b""
        unary = grpc.unary_unary_rpc_method_handler
        handlers = grpc.method_handlers_generic_handler(SERVICE, {
            "GetConfig": unary(lambda request, context: si.config_bytes()),
            "GetServerState": unary(lambda request, context: si.state_bytes()),
            "SubmitReturn": unary(submit_return),
            "SubmitReturns": unary(submit_returns),
        })
        server = grpc.server(futures.ThreadPoolExecutor(max_workers=max_workers),
                             options=[("grpc.max_send_message_length", MAX_MESSAGE_LENGTH),
                                      ("grpc.max_receive_message_length", MAX_MESSAGE_LENGTH)],
                             compression=grpc.Compression.Gzip)
        server.add_generic_rpc_handlers((handlers,))
        self.bound_port = server.add_insecure_port("{}:{}".format(address, port))
        server.start()
        self.grpc_server = server

    def stop(self, grace=10):
        if self.grpc_server is not None:
            self.grpc_server.stop(grace=grace)
        self.server_interface.cleanup()


class GRPCWorker(object):
    """worker/grpc_worker.py:6-21."""

    def __init__(self, state):
        self.grpc_server = RPCServer(state)

    def collect_returns(self, batch_size=None, current_epoch=None, max_delayed_return=None, timeout=None):
        return self.grpc_server.get_returns_batch(batch_size=batch_size, current_epoch=current_epoch,
                                                  max_delayed_return=max_delayed_return, timeout=timeout)

    def update(self, state):
        self.grpc_server.update(state)

    def start(self, address, port):
        self.grpc_server.start(address=address, port=port)

    def stop(self):
        self.grpc_server.stop()


class RPCClient(object):
    """Worker-side counterpart (networking/client.py:11-92) for GPU workers that ship `ReturnBatch`es: same flags
    and method names, raw-bytes calls."""
    OPERATION_SUCCESSFUL_FLAG = 0
    NEW_STATE_FLAG = 1
    NEW_EXPERIMENT_FLAG = 2
    RPC_FAILED_FLAG = 3

    def __init__(self):
        self.channel = None
        self.current_state = FDState()

    def connect(self, address="localhost", port=50051):
        import grpc
        self.channel = grpc.insecure_channel("{}:{}".format(address, port),
                                             options=[("grpc.max_send_message_length", MAX_MESSAGE_LENGTH),
                                                      ("grpc.max_receive_message_length", MAX_MESSAGE_LENGTH)],
                                             compression=grpc.Compression.Gzip)
        call = self.channel.unary_unary
        self._get_config = call("/%s/GetConfig" % SERVICE)
        self._get_state = call("/%s/GetServerState" % SERVICE)
        self._submit_return = call("/%s/SubmitReturn" % SERVICE)
        self._submit_returns = call("/%s/SubmitReturns" % SERVICE)

    def submit_return(self, ret):
        self._submit_return(wire.encode_return(ret))

    def submit_returns(self, returns):
        self._submit_returns(wire.encode_return_array(returns))

    def get_server_state(self):
        try:
            state = wire.decode_server_state(self._get_state(b""))
        except Exception:                                        # noqa: BLE001 (client.py:70-74: any failure is a flag)
            return self.RPC_FAILED_FLAG
        cur = self.current_state
        new_experiment = state.experiment_id != cur.experiment_id
        if new_experiment or state.epoch != cur.epoch:
            if new_experiment:
                cur.cfg = wire.decode_config(self._get_config(b""))
            for k in ("strategy_frames", "strategy_history", "policy_params", "obs_stats", "epoch", "experiment_id"):
                setattr(cur, k, getattr(state, k))
            return self.NEW_EXPERIMENT_FLAG if new_experiment else self.NEW_STATE_FLAG
        return self.OPERATION_SUCCESSFUL_FLAG

    def disconnect(self):
        self.channel.close()
