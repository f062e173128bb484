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
