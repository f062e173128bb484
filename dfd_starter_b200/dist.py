"""Population sharding helpers (one process per GPU).  Members are independent given
(theta, table, idx), so the path shards with no data-path collective in the forward; the
estimator needs the rewards of ALL ranks for its mean / std (N doubles) and one parameter-sized
all_reduce(SUM) of the partial gradients."""
import numpy as np
import torch
import torch.distributed as dist


def shard_pairs(n_pairs, rank, world):
    """Contiguous, balanced [lo, hi) slice of the antithetic pairs for `rank`; a +/- pair always
    stays on one rank so its table row is read once."""
    base, rem = divmod(int(n_pairs), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_gather_rewards(local_rewards, group=None):
    """Rewards of every rank's accepted returns, concatenated in rank order (float64 numpy).
    Shards may be ragged, so sizes are exchanged first."""
    world = dist.get_world_size(group)
    local = torch.as_tensor(np.ascontiguousarray(local_rewards, dtype=np.float64))
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    n = torch.tensor([local.numel()], dtype=torch.int64, device=dev)
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    mx = max(sizes) if sizes else 0
    buf = torch.zeros(max(mx, 1), dtype=torch.float64, device=dev)
    buf[:local.numel()] = local.to(dev)
    out = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    return np.concatenate([o[:s].cpu().numpy() for o, s in zip(out, sizes)]) if mx else np.zeros(0)


class PeerExchange(object):
    """The sharded learner's one exchange step over NVLink peer memory (csrc/xchg_allreduce.cu).

    Every rank creates a mailbox in its own HBM, the 64-byte CUDA IPC handles travel through the process
    group once (all_gather_object), every rank maps its peers' mailboxes, and from then on
    `allreduce(grad_partial, stats5, grad_out)` is ONE kernel launch per step and rank: push to all peers,
    publish, wait, sum in rank order, scale by 1/std of all ranks' rewards.  No NCCL call on the step."""

    def __init__(self, ctx, n_params, group=None):
        import ctypes as C
        from . import _lib
        self.ctx, self.lib, self.group = ctx, ctx.lib, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.P = int(n_params)
        nbytes = int(self.lib.dfd_xchg_mailbox_bytes(self.P, self.world))
        if nbytes == 0:
            raise _lib.DfdError("PeerExchange: world size %d not supported (1..16)" % self.world)
        mine = C.c_void_p()
        handle = C.create_string_buffer(64)
        _lib.check(self.lib.dfd_xchg_mailbox_create(ctx.handle, nbytes, C.byref(mine), handle), "dfd_xchg_mailbox_create")
        self._mine = mine.value
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle.raw), group=group)
        self._peers = []
        ptrs = []
        for r, h in enumerate(handles):
            if r == self.rank:
                ptrs.append(self._mine)
                continue
            p = C.c_void_p()
            _lib.check(self.lib.dfd_xchg_mailbox_open(ctx.handle, C.create_string_buffer(h, 64), C.byref(p)),
                       "dfd_xchg_mailbox_open (rank %d)" % r)
            self._peers.append(p.value)
            ptrs.append(p.value)
        self.table = torch.tensor(ptrs, dtype=torch.int64).to(ctx.device)     # device array of world pointers
        torch.cuda.synchronize(ctx.device)
        dist.barrier(group=group)               # every mailbox is zero-filled and mapped before the first step

    def allreduce(self, grad_partial, stats5, grad_out):
        from . import _lib
        from .device import ptr
        _lib.check(self.lib.dfd_xchg_allreduce(self.ctx.handle, ptr(self.table), self.rank, self.world, self.P,
                                               ptr(grad_partial), ptr(stats5), ptr(grad_out), self.ctx.stream),
                   "dfd_xchg_allreduce")

    def gather_f64(self, src, n, dst):
        """dst[world, n] = every rank's src[:n] (float64 device tensors), over the same mailboxes: the rewards all-gather of
        fd_state batches without NCCL.  Every rank must call it with the same n."""
        from . import _lib
        from .device import ptr
        _lib.check(self.lib.dfd_xchg_gather_f64(self.ctx.handle, ptr(self.table), self.rank, self.world, self.P, ptr(src),
                                                int(n), ptr(dst), self.ctx.stream), "dfd_xchg_gather_f64")

    def close(self):
        """Unmap the peers' mailboxes, then - once EVERY rank has unmapped (barrier) - free this rank's own."""
        for p in self._peers:
            self.lib.dfd_xchg_mailbox_close(self.ctx.handle, p)
        self._peers = []
        if self._mine:
            dist.barrier(group=self.group)
            self.lib.dfd_xchg_mailbox_destroy(self.ctx.handle, self._mine)
            self._mine = None
