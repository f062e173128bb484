"""`FiniteDifferences` with the reference's constructor and `step` contract
(learner/finite_differences.py:6-114), evaluated on the device.

    learner = FiniteDifferences(policy, gradient_optimizer, omega, noise_source,
                                noise_std=0.1, batch_size=100, ent_coef=0.0, max_delayed_return=10)
    update_size = learner.step(batch, policy_reward, policy_novelty, policy_entropy)

`batch` is a list of FDReturn-like objects (fields epoch, encoded_noise, reward).
Host logic kept here: the epoch acceptance test (:80-85), key parsing, the ring
bookkeeping of theta-history rows.  Device work (three C-ABI calls, no host
synchronisation between them): dfd_fd_prepare -> dfd_fd_reduce -> [one NCCL
allreduce of P floats when the population is sharded] -> dfd_dsgd_step.

`policy` may be one of this package's device policies or any object with the
reference's `get_trainable_flat/set_trainable_flat/num_params` (e.g. a reference
`Policy`); in the second case theta is mirrored back to it after every step
because the drivers serialise it next (run_server.py:192-197).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .device import get_context, ptr, aligned_ptr
from .dsgd import is_dsgd
from .noise_sources import parse_key


class _Staging(object):
    """Pinned host + device staging for one batch: rewards f64 | idx i64 | hist_row i32 | sign i8."""

    def __init__(self, device, cap, ctx=None):
        self.cap = cap
        self.ctx = ctx
        nbytes = cap * (8 + 8 + 4 + 1)
        self.host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
        self.dev = torch.empty(nbytes, dtype=torch.uint8, device=device)
        o = 0
        self.views = {}
        host_np = self.host.numpy()                 # numpy views of the pinned buffer: filling them costs ~1 us per array
        for name, dt, ndt, sz in (("reward", torch.float64, np.float64, 8), ("idx", torch.int64, np.int64, 8),
                                  ("hist_row", torch.int32, np.int32, 4), ("sign", torch.int8, np.int8, 1)):
            self.views[name] = (host_np[o:o + cap * sz].view(ndt), self.dev[o:o + cap * sz].view(dt))
            o += cap * sz

    def upload(self, n, **arrays):
        for k, a in arrays.items():
            self.views[k][0][:n] = a
        if self.ctx is not None:
            # read by the SMs through the pinned buffer's device alias: does not queue on the host-to-device copy engine
            # behind a large upload (the next step's observations) - include/dfd_b200.h, dfd_host_stage
            _lib.check(self.ctx.lib.dfd_host_stage(self.ctx.handle, self.host.data_ptr(), self.dev.data_ptr(),
                                                   self.host.numel(), self.ctx.stream), "dfd_host_stage")
        else:
            self.dev.copy_(self.host, non_blocking=True)
        return {k: v[1] for k, v in self.views.items()}


class FiniteDifferences(object):
    def __init__(self, policy, gradient_optimizer, omega, noise_source, noise_std=0.1, batch_size=100, ent_coef=0.0,
                 max_delayed_return=10, paired=False, process_group=None, device=None, sync_policy=True,
                 peer_exchange=None, fused_step=True, exchange_mode="fd_return"):
        self.max_delayed_return = max_delayed_return
        self.ent_coef = ent_coef
        self.noise_std = noise_std
        self.policy = policy
        self.gradient_optimizer = gradient_optimizer
        self.noise_source = noise_source
        self.omega = omega
        self.paired = paired                    # extension: batch = [R plus-members | R minus-members]
        self.process_group = process_group      # extension: population sharded over ranks, one allreduce per step
        # dist.PeerExchange: the exchange runs as one peer-memory kernel instead of NCCL calls whenever the batch
        # is antithetic pairs of the current epoch (otherwise: rewards all-gather + NCCL all_reduce)
        self.peer_exchange = peer_exchange
        # every rank must take the same exchange path every step, so the mode is fixed per learner:
        #   "fd_return": antithetic pairs of the current epoch only - standardisation deferred, ONE peer-memory kernel;
        #   "general":   any accepted batch (delayed epochs = fd_state mode, one-sided members): rewards gathered over
        #                peer memory, coefficients with the statistics of all ranks' returns, partial gradients summed
        #                over peer memory (two peer kernels per step, no NCCL on the device-resident step)
        if exchange_mode not in ("fd_return", "general"):
            raise _lib.DfdError("exchange_mode must be 'fd_return' or 'general'")
        self.exchange_mode = exchange_mode
        self._gathered = None
        # short parameter vectors: prepare + reduce [+ exchange] + DSGD run as ONE kernel (csrc/fd_tail.cu)
        self.fused_step = fused_step
        self._fused_scratch = {}
        self.sync_policy = sync_policy
        self.using_dsgd = is_dsgd(gradient_optimizer)

        self.ctx = get_context(device if device is not None else getattr(getattr(policy, "ctx", None), "device", None))
        self.lib = self.ctx.lib
        dev = self.ctx.device
        # a shared table lives on the device for the whole run; other noise sources (RNGNoiseSource, SimpleNoiseSource:
        # utils/noise_sources.py:4-33) are staged per batch as a RowTable - drawn on the device (RNGNoiseSource) or on the host
        self.table = noise_source.device_table if hasattr(noise_source, "device_table") else None
        P = int(policy.num_params)
        self.P = P
        self.Ps = (P + 3) // 4 * 4              # row stride of history / dist rows (16-byte aligned rows)
        H = max(int(max_delayed_return), 1)
        self.H = H

        # theta lives on the device; share the policy's tensor when it has one
        if hasattr(policy, "theta") and torch.is_tensor(policy.theta) and policy.theta.is_cuda:
            self.theta = policy.theta
            self._host_policy = False
        else:
            self.theta = torch.from_numpy(np.asarray(policy.get_trainable_flat(), dtype=np.float32).copy()).to(dev)
            self._host_policy = True
        self.grad = torch.zeros(P, dtype=torch.float32, device=dev)
        self._grad_partial = torch.zeros((P + 3) // 4 * 4, dtype=torch.float32, device=dev)
        self._stats5 = torch.zeros(8, dtype=torch.float64, device=dev)
        self.hist = torch.zeros(H, self.Ps, dtype=torch.float32, device=dev)
        self.dist = torch.zeros(H, self.Ps, dtype=torch.float32, device=dev)
        self.hist[0, :P].copy_(self.theta)                       # policy_history = [(theta0, 0)]   (:16)
        self._hist_epoch = [0]                                   # epoch held by each ring row
        self._dist_epoch = {}                                    # epoch -> dist row (delayed epochs only)
        self.epoch = 0
        self.discarded_returns = 0
        self._update_size = torch.zeros(4, dtype=torch.float32, device=dev)   # [0] is written by the kernels (16-byte unit)
        self._update_host16 = torch.zeros(4, dtype=torch.float32).pin_memory()
        self._theta_host = torch.zeros(P, dtype=torch.float32).pin_memory()
        self._stage = None
        self._rows_cap = 0
        self._dsgd_scratch = self.ctx.zeros_bytes(self.lib.dfd_dsgd_scratch_bytes(P))
        self._ensure_capacity(max(int(batch_size), 1))

    # ------------------------------------------------------------------ buffers
    def _ensure_capacity(self, n):
        if self._stage is not None and n <= self._stage.cap:
            return
        cap = max(n, 16)
        dev = self.ctx.device
        self._stage = _Staging(dev, cap, self.ctx)
        self._rows_cap = cap + self.H
        self._row_ptr = torch.zeros(self._rows_cap, dtype=torch.int64, device=dev)
        self._row_coef = torch.zeros(self._rows_cap, dtype=torch.float32, device=dev)
        self._rows = _lib.DfdFdRows(self._row_ptr.data_ptr(), self._row_coef.data_ptr(), self._rows_cap)
        self._prep_scratch = self.ctx.zeros_bytes(self.lib.dfd_fd_prepare_scratch_bytes(cap, self.H))
        self._red_scratch = self.ctx.zeros_bytes(self.lib.dfd_fd_reduce_scratch_bytes(self.ctx.handle, self.P, self._rows_cap))
        self._red_scratch_rows = self._rows_cap

    # ------------------------------------------------------------------ reference attributes
    @property
    def gradient_memory(self):
        """fp64 copy of the last gradient (finite_differences.py:20,49)."""
        return self.grad.double().cpu().numpy()

    @property
    def dist_map(self):
        """Accepted epochs -> ring row (the reference maps epoch -> theta_e - theta_now)."""
        m = {e: r for e, r in self._dist_epoch.items()}
        m[self.epoch] = -1
        return m

    @property
    def policy_history(self):
        return [(self.hist[r, :self.P].cpu().numpy(), e) for r, e in
                sorted(enumerate(self._hist_epoch), key=lambda t: t[1])]

    # ------------------------------------------------------------------ the step
    def step(self, batch, policy_reward, policy_novelty=None, policy_entropy=None):
        soa = getattr(batch, "soa", None)
        if soa is not None:                      # untouched ReturnBatch from the batched Worker: arrays as they are
            return self.step_arrays(soa[0], soa[1], soa[2], soa[3], policy_reward)
        if self.table is None:
            keys = getattr(batch, "keys", None)
            if keys is not None and getattr(batch, "_records", 0) is None:
                # untouched ReturnBatch of a keyed noise source (the batched Worker's): arrays and keys as they are
                return self._step_keyed_arrays(np.asarray(batch.epoch, dtype=np.int64), np.asarray(batch.reward, dtype=np.float64),
                                               keys, policy_reward)
            return self._step_host_noise(batch, policy_reward)
        epochs = np.fromiter((int(r.epoch) for r in batch), dtype=np.int64, count=len(batch))
        rewards = np.fromiter((float(r.reward) for r in batch), dtype=np.float64, count=len(batch))
        keys = [parse_key(r.encoded_noise) for r in batch]
        idx = np.fromiter((k[0] for k in keys), dtype=np.int64, count=len(batch))
        sign = np.fromiter((k[1] for k in keys), dtype=np.int8, count=len(batch))
        return self.step_arrays(epochs, idx, sign, rewards, policy_reward)

    def _step_host_noise(self, batch, policy_reward):
        """Noise sources without a device table, per-record form (finite_differences.py:80-112 over a list of FDReturn)."""
        batch = list(batch)
        epochs = np.fromiter((int(r.epoch) for r in batch), dtype=np.int64, count=len(batch))
        rewards = np.fromiter((float(r.reward) for r in batch), dtype=np.float64, count=len(batch))
        return self._step_keyed_arrays(epochs, rewards, [r.encoded_noise for r in batch], policy_reward)

    def _step_keyed_arrays(self, epochs, rewards, keys, policy_reward):
        """Noise sources without a device table: the vector of every ACCEPTED return is recovered from its key, in batch
        order like finite_differences.py:87 (too-old returns are rejected before they are decoded, :82-85) - redrawn on
        the device for RNGNoiseSource (bit-identical to decode(), csrc/rng_normal.cu; the source's generator ends where the
        last decode() would have left it), decoded on the host otherwise -, staged as a RowTable, and the ordinary device
        step runs over it."""
        from .noise_sources import RowTable
        n = len(keys)
        ok = np.array([(e == self.epoch) or (e in self._dist_epoch) for e in epochs], dtype=bool)
        if not ok.any():
            return self.step_arrays(epochs, np.zeros(n, np.int64), np.ones(n, np.int8), rewards, policy_reward)
        good = keys.take(np.nonzero(ok)[0]) if hasattr(keys, "take") else [keys[j] for j in np.nonzero(ok)[0]]
        if getattr(self.noise_source, "device_rows", False):
            rt = RowTable(self.ctx, shape=(len(good), self.P))
            self.noise_source.decode_rows(self.ctx, good, rt.raw, rt.Ps)
            rt.build()
        else:
            rt = RowTable(self.ctx, np.stack([np.asarray(self.noise_source.decode(k), dtype=np.float32) for k in good]))
        idx = np.zeros(n, dtype=np.int64)
        idx[ok] = rt.idx
        self.table = rt
        try:
            return self.step_arrays(epochs, idx, np.ones(n, dtype=np.int8), rewards, policy_reward)
        finally:
            self.table = None

    def step_arrays(self, epochs, idx, sign, rewards, policy_reward, all_rewards=None):
        """SoA form of `step`: int64 epochs/idx, int8 sign, float64 rewards (host arrays).
        all_rewards: with a process group, the rewards of ALL ranks' accepted returns (the
        standardisation is global); gathered here when omitted."""
        epochs = np.asarray(epochs, dtype=np.int64)
        idx = np.asarray(idx, dtype=np.int64)
        sign = np.asarray(sign, dtype=np.int8)
        rewards = np.asarray(rewards, dtype=np.float64)
        if self.paired and epochs.shape[0]:
            sel = self._pair_order(idx, sign)
            if sel is not None:
                epochs, idx, sign, rewards = epochs[sel], idx[sel], sign[sel], rewards[sel]
        n_in = epochs.shape[0]
        # finite_differences.py:80-85: a return is usable iff its epoch is still in the distance map
        hist_row = np.full(n_in, -2, dtype=np.int32)
        hist_row[epochs == self.epoch] = -1
        for e, r in self._dist_epoch.items():
            hist_row[epochs == e] = r
        keep = hist_row > -2
        n_bad = int(n_in - keep.sum())
        if n_bad:
            for e in epochs[~keep]:
                print("FINITE DIFFERENCE LEARNER RECEIVED RETURN THAT WAS TOO OLD")
                print("RECEIVED EPOCH:", int(e), "ACCEPTABLE EPOCHS:", self.dist_map.keys())
            self.discarded_returns += n_bad
            if self.paired:
                # keep pairs intact: drop both members of a pair if either is unusable
                R = n_in // 2
                pk = keep[:R] & keep[R:2 * R]
                keep = np.concatenate([pk, pk])
        if policy_reward is None:
            policy_reward = 0
        # keys come off the wire (FDReturn.encoded_noise): a slice that does not lie inside the table would be an
        # out-of-bounds read on the device.  The reference fails on the short slice (shape mismatch in :88-89); here the
        # offending returns are dropped and counted like too-old ones
        size = int(getattr(self.table, "size", 0) or 0)
        if size and n_in:
            in_table = (idx >= 0) & (idx + self.P <= size)
            n_oob = int((keep & ~in_table).sum())
            if n_oob:
                print("FINITE DIFFERENCE LEARNER RECEIVED %d RETURN(S) WHOSE NOISE KEY IS OUTSIDE THE TABLE" % n_oob)
                self.discarded_returns += n_oob
                keep = keep & in_table
                if self.paired:
                    R = n_in // 2
                    pk = keep[:R] & keep[R:2 * R]
                    keep = np.concatenate([pk, pk, np.zeros(n_in - 2 * R, dtype=bool)])
        idx = idx[keep]
        sign = sign[keep]
        rewards = rewards[keep]
        hist_row = hist_row[keep]
        n = idx.shape[0]
        pg = self.process_group
        if pg is None and n == 0:
            return 0                                             # :30-31, no update, epoch unchanged
        stats = None
        general_peer = pg is not None and self.peer_exchange is not None and self.exchange_mode == "general"
        if pg is not None and self.peer_exchange is not None and not general_peer:
            # the mode is fixed per learner (every rank must take the same path every step): a learner built
            # with a PeerExchange only accepts what the one-kernel exchange can express
            if not (self.paired and n % 2 == 0 and bool((hist_row == -1).all())):
                raise _lib.DfdError("a learner with peer_exchange takes antithetic pairs of the current epoch only "
                                    "(fd_return mode); build it without peer_exchange for delayed returns")
            if n == 0:                           # this shard is empty this step: still take part in the exchange
                self._grad_partial.zero_()
                self._stats5.zero_()
                self.peer_exchange.allreduce(self._grad_partial, self._stats5, self.grad)
            else:
                self._ensure_capacity(n)
                d = self._stage.upload(n, reward=rewards, idx=idx, hist_row=hist_row, sign=sign)
                scratch = self._fused_scratch_for(n, 1)
                if scratch is not None:
                    return self._apply_update(fused_call=self._fused_step_call(scratch, d["idx"], d["sign"], d["reward"], n, 1,
                                                                               policy_reward))
                self._fused_exchange(d["idx"], d["sign"], d["reward"], d["hist_row"], n, policy_reward)
            return self._apply_update()
        if pg is not None:
            import torch.distributed as dist
            if all_rewards is None:
                from .dist import all_gather_rewards
                all_rewards = all_gather_rewards(rewards, pg)
            if all_rewards.shape[0] == 0:
                return 0
            stats = torch.from_numpy(np.ascontiguousarray(all_rewards, dtype=np.float64)).to(self.ctx.device)

        self._ensure_capacity(n)
        st = self.ctx.stream
        n_hist = len(self._dist_epoch) and (max(self._dist_epoch.values()) + 1)
        if n > 0:
            d = self._stage.upload(n, reward=rewards, idx=idx, hist_row=hist_row, sign=sign)
            paired = 1 if (self.paired and n % 2 == 0) else 0
            if pg is None and bool((hist_row == -1).all()):
                scratch = self._fused_scratch_for(n, paired)
                if scratch is not None:
                    return self._apply_update(fused_call=self._fused_step_call(scratch, d["idx"], d["sign"], d["reward"], n,
                                                                               paired, policy_reward))
            _lib.check(self.lib.dfd_fd_prepare(
                self.ctx.handle, self.table.ref(), self.P, ptr(d["reward"]), ptr(d["idx"]), ptr(d["sign"]),
                ptr(d["hist_row"]), n, paired, float(policy_reward), float(self.noise_std), ptr(self.dist), self.Ps,
                int(n_hist), ptr(stats), 0 if stats is None else int(stats.shape[0]), C.byref(self._rows),
                aligned_ptr(self._prep_scratch), self._prep_scratch.numel() - 256, st), "dfd_fd_prepare")
            n_rows = (n // 2 if paired else n) + int(n_hist)
            _lib.check(self.lib.dfd_fd_reduce(
                self.ctx.handle, C.byref(self._rows), n_rows, self.P, ptr(self.grad),
                aligned_ptr(self._red_scratch), self._red_scratch.numel() - 256, st), "dfd_fd_reduce")
        else:
            self.grad.zero_()
        if general_peer:                        # partial gradients are already standardised: stats with count 0 = identity scale
            self._grad_partial[:self.P].copy_(self.grad)
            self._stats5.zero_()
            self.peer_exchange.allreduce(self._grad_partial, self._stats5, self.grad)
        elif pg is not None:
            import torch.distributed as dist
            dist.all_reduce(self.grad, op=dist.ReduceOp.SUM, group=pg)     # the one parameter-sized exchange
        return self._apply_update()

    @staticmethod
    def _pair_order(idx, sign):
        """paired=True promises the device kernels the layout [R plus-members | R minus-members] of the same R table
        rows (they pair entry r with r + R and read the row pointer from entry r only).  Returns None when the batch
        already has it, else the selection that establishes it: eval members (sign 0) are dropped, as the drivers drop
        them before `step` (run_sequential.py:137-147), and complete +/- pairs in any other order - concatenated RPC
        chunks [+A -A +B -B], LIFO pops (server.py:80) - are regrouped, plus members kept in arrival order.  A batch
        that does not consist of complete pairs is refused rather than mispaired silently."""
        n = idx.shape[0]
        R = n // 2
        if n % 2 == 0 and (sign[:R] == 1).all() and (sign[R:] == -1).all() and np.array_equal(idx[:R], idx[R:]):
            return None
        pos, neg = np.flatnonzero(sign == 1), np.flatnonzero(sign == -1)
        pos = pos[np.argsort(idx[pos], kind="stable")]
        neg = neg[np.argsort(idx[neg], kind="stable")]
        if pos.shape[0] != neg.shape[0] or not np.array_equal(idx[pos], idx[neg]):
            raise _lib.DfdError("paired=True needs complete antithetic pairs (a '+i' and a '-i' return per table row); the "
                                "%d plus / %d minus members received do not pair up - build the learner with paired=False "
                                "for one-sided or mixed batches" % (pos.shape[0], neg.shape[0]))
        by_arrival = np.argsort(pos, kind="stable")
        return np.concatenate([pos[by_arrival], neg[by_arrival]])

    def _fused_scratch_for(self, n, paired):
        """Zero-filled scratch of the one-kernel step for this batch shape, or None when the shape is not served."""
        if not (self.fused_step and self.using_dsgd):
            return None
        key = (int(n), int(paired))
        if key not in self._fused_scratch:
            nbytes = int(self.lib.dfd_fd_step_fused_scratch_bytes(self.ctx.handle, self.P, int(n), int(paired)))
            self._fused_scratch[key] = self.ctx.zeros_bytes(nbytes) if nbytes else None
        return self._fused_scratch[key]

    def _fused_step_call(self, scratch, idx_d, sign_d, reward_d, n, paired, policy_reward):
        """Closure handed to _apply_update: launches dfd_fd_step_fused with the optimizer's current step size."""
        world = 1 if self.process_group is None else self.peer_exchange.world
        rank = 0 if self.process_group is None else self.peer_exchange.rank
        boxes = None if self.process_group is None else ptr(self.peer_exchange.table)

        def call(lr, lr_scale, n_valid, write_row):
            _lib.check(self.lib.dfd_fd_step_fused(
                self.ctx.handle, self.table.ref(), self.P, ptr(reward_d), ptr(idx_d), ptr(sign_d), int(n), int(paired),
                float(policy_reward), float(self.noise_std), ptr(self.theta), ptr(self.grad), lr, lr_scale, ptr(self.hist),
                ptr(self.dist), self.Ps, n_valid, write_row, ptr(self._update_size), boxes, rank, world,
                aligned_ptr(scratch), scratch.numel() - 256, self.ctx.stream), "dfd_fd_step_fused")
        return call

    def _fused_exchange(self, idx_d, sign_d, reward_d, hist_row_d, n, policy_reward):
        """prepare (standardisation deferred) -> reduce -> ONE peer-memory exchange kernel -> self.grad."""
        st = self.ctx.stream
        _lib.check(self.lib.dfd_fd_prepare_partial(
            self.ctx.handle, self.table.ref(), self.P, ptr(reward_d), ptr(idx_d), ptr(sign_d), ptr(hist_row_d), n,
            float(policy_reward), float(self.noise_std), C.byref(self._rows), ptr(self._stats5),
            aligned_ptr(self._prep_scratch), self._prep_scratch.numel() - 256, st), "dfd_fd_prepare_partial")
        _lib.check(self.lib.dfd_fd_reduce(
            self.ctx.handle, C.byref(self._rows), n // 2, self.P, ptr(self._grad_partial), aligned_ptr(self._red_scratch),
            self._red_scratch.numel() - 256, st), "dfd_fd_reduce")
        self.peer_exchange.allreduce(self._grad_partial, self._stats5, self.grad)

    def step_device(self, idx_d, sign_d, reward_d, n, policy_reward=0.0, hist_row_d=None, stats_d=None):
        """Device-resident SoA step (no host synchronisation, CUDA-graph capturable): the n returns
        are already in HBM as int64 idx / int8 sign / float64 reward (and optional int32 hist_row;
        default: all from the current epoch).  DSGD only.  Returns nothing; `last_update_size()`
        reads the update magnitude back when wanted."""
        if not self.using_dsgd:
            raise _lib.DfdError("step_device needs the DSGD optimizer")
        self._ensure_capacity(n)
        st = self.ctx.stream
        if hist_row_d is None:
            if getattr(self, "_minus1", None) is None or self._minus1.shape[0] < n:
                self._minus1 = torch.full((max(n, 16),), -1, dtype=torch.int32, device=self.ctx.device)
            hist_row_d, n_hist = self._minus1, 0
        else:
            n_hist = len(self._dist_epoch) and (max(self._dist_epoch.values()) + 1)
        paired = 1 if (self.paired and n % 2 == 0) else 0
        general_peer = self.process_group is not None and self.peer_exchange is not None and self.exchange_mode == "general"
        if general_peer:
            # rewards of every rank over peer memory (every rank submits the same n), then the ordinary prepare with the
            # statistics of all returns
            world = self.peer_exchange.world
            if self._gathered is None or self._gathered.shape[0] < world * n:
                self._gathered = torch.empty(world * max(n, 16), dtype=torch.float64, device=self.ctx.device)
            self.peer_exchange.gather_f64(reward_d, n, self._gathered)
            stats_d = self._gathered[:world * n]
        if self.process_group is not None and self.peer_exchange is not None and not general_peer:
            if not (paired and n_hist == 0 and stats_d is None):
                raise _lib.DfdError("a learner with peer_exchange takes antithetic pairs of the current epoch only")
            scratch = self._fused_scratch_for(n, 1)
            if scratch is not None:
                self._apply_update(sync=False, fused_call=self._fused_step_call(scratch, idx_d, sign_d, reward_d, n, 1, policy_reward))
                return
            self._fused_exchange(idx_d, sign_d, reward_d, hist_row_d, n, policy_reward)
            self._apply_update(sync=False)
            return
        if self.process_group is None and n_hist == 0:
            scratch = self._fused_scratch_for(n, paired)
            if scratch is not None:
                self._apply_update(sync=False, fused_call=self._fused_step_call(scratch, idx_d, sign_d, reward_d, n, paired,
                                                                               policy_reward))
                return
        _lib.check(self.lib.dfd_fd_prepare(
            self.ctx.handle, self.table.ref(), self.P, ptr(reward_d), ptr(idx_d), ptr(sign_d), ptr(hist_row_d), n,
            paired, float(policy_reward), float(self.noise_std), ptr(self.dist), self.Ps, int(n_hist), ptr(stats_d),
            0 if stats_d is None else int(stats_d.shape[0]), C.byref(self._rows), aligned_ptr(self._prep_scratch),
            self._prep_scratch.numel() - 256, st), "dfd_fd_prepare")
        n_rows = (n // 2 if paired else n) + int(n_hist)
        _lib.check(self.lib.dfd_fd_reduce(
            self.ctx.handle, C.byref(self._rows), n_rows, self.P, ptr(self._grad_partial if general_peer else self.grad),
            aligned_ptr(self._red_scratch), self._red_scratch.numel() - 256, st), "dfd_fd_reduce")
        if general_peer:
            self._stats5.zero_()                # count 0: the partial gradients are already standardised
            self.peer_exchange.allreduce(self._grad_partial, self._stats5, self.grad)
        elif self.process_group is not None:
            import torch.distributed as dist
            dist.all_reduce(self.grad, op=dist.ReduceOp.SUM, group=self.process_group)
        self._apply_update(sync=False)

    def last_update_size(self):
        return float(self._update_size.cpu()[0])

    # ------------------------------------------------------------------ optimizer + history
    def _ring_write_row(self):
        if len(self._hist_epoch) < self.H:
            return len(self._hist_epoch)
        return int(np.argmin(self._hist_epoch))

    def _apply_update(self, sync=True, fused_call=None):
        st = self.ctx.stream
        n_valid = len(self._hist_epoch)
        write_row = self._ring_write_row()
        if self.using_dsgd:
            self.gradient_optimizer.adjust_lr(self.omega)                    # :51-52
            lr, lr_scale = float(self.gradient_optimizer.lr), float(self.gradient_optimizer.lr_scale)
            if fused_call is not None:
                fused_call(lr, lr_scale, n_valid, write_row)
            else:
                _lib.check(self.lib.dfd_dsgd_step(
                    self.ctx.handle, ptr(self.theta), ptr(self.grad), self.P, lr, lr_scale, ptr(self.hist), ptr(self.dist),
                    self.Ps, n_valid, write_row, ptr(self._update_size), aligned_ptr(self._dsgd_scratch),
                    self._dsgd_scratch.numel() - 256, st), "dfd_dsgd_step")
            if hasattr(self.gradient_optimizer, "steps"):
                self.gradient_optimizer.steps += 1
            if not sync:
                return self._advance_epoch(write_row, None)
            # small read-backs go through dfd_host_stage (SM writes into pinned memory), not the DMA engine, so they do
            # not wait behind a large upload in flight; long parameter vectors keep the copy engine
            self.ctx.host_stage(self._update_size, self._update_host16)
            if self._host_policy and self.sync_policy:
                if self.P * 4 <= (1 << 20) and self.theta.data_ptr() % 16 == 0:
                    self.ctx.host_stage(self.theta, self._theta_host)
                else:
                    self._theta_host.copy_(self.theta, non_blocking=True)
            torch.cuda.current_stream(self.ctx.device).synchronize()
            update_size = float(self._update_host16[0])
            if self._host_policy and self.sync_policy:
                self.policy.set_trainable_flat(self._theta_host.numpy())
        else:
            # any other torch optimizer: it owns the update rule (finite_differences.py:54-57); the device keeps the
            # history / distance rows
            self.gradient_optimizer.zero_grad()
            if not self._host_policy:
                # device policy: its parameters() is ONE flat nn.Parameter aliasing theta, so the optimizer updates the
                # vector the kernels read, on the device, with no host round trip
                before = self.theta.clone()
                self.policy.set_grad_from_flat(-self.grad)                       # policy.py:63-70
                self.gradient_optimizer.step()
                update_size = float(torch.linalg.vector_norm(before - self.theta.detach()))
            else:
                flat = np.asarray(self.policy.get_trainable_flat(), dtype=np.float32).copy()
                with torch.enable_grad():
                    self.policy.set_grad_from_flat(-self.gradient_memory)
                self.gradient_optimizer.step()
                new_flat = np.asarray(self.policy.get_trainable_flat(), dtype=np.float32)
                update_size = float(np.linalg.norm(flat - new_flat))
                self.theta.copy_(torch.from_numpy(new_flat.copy()))
            zero = torch.zeros_like(self.grad)
            _lib.check(self.lib.dfd_dsgd_step(
                self.ctx.handle, ptr(self.theta), ptr(zero), self.P, 0.0, 0.0, ptr(self.hist), ptr(self.dist),
                self.Ps, n_valid, write_row, ptr(self._update_size), aligned_ptr(self._dsgd_scratch),
                self._dsgd_scratch.numel() - 256, st), "dfd_dsgd_step")
        return self._advance_epoch(write_row, update_size)

    def _advance_epoch(self, write_row, update_size):
        self.epoch += 1                                                      # :60
        # :66-78 — the distance map is rebuilt from the history BEFORE the new theta is appended, so
        # H+1 epochs stay acceptable while the ring holds H rows
        self._dist_epoch = {e: r for r, e in enumerate(self._hist_epoch)}
        if write_row < len(self._hist_epoch):
            self._hist_epoch[write_row] = self.epoch
        else:
            self._hist_epoch.append(self.epoch)
        return update_size
