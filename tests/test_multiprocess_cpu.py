"""World-size-2 `gloo` test of the sharded estimator's host logic (CPU): pair sharding, the ragged
reward all-gather, and the algebra the GPU path relies on — every rank standardises its shard with the
GLOBAL reward statistics, reduces only its own rows, and one all_reduce(SUM) of P values reproduces the
single-process estimator."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import dfd_oracle as O
from dfd_starter_b200.dist import shard_pairs, all_gather_rewards


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, R, P, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    noise = O.NoiseTableOracle(200_000, P, 124)              # replicated table, same seed on every rank
    rng = np.random.RandomState(0)
    idx_all = rng.randint(0, 200_000 - P, size=R).astype(np.int64)
    rew_plus, rew_minus = rng.randn(R) * 2 + 3, rng.randn(R) * 2 + 3
    lo, hi = shard_pairs(R, rank, world)
    idx = np.concatenate([idx_all[lo:hi], idx_all[lo:hi]])
    sign = np.concatenate([np.ones(hi - lo), -np.ones(hi - lo)])
    rewards = np.concatenate([rew_plus[lo:hi], rew_minus[lo:hi]])
    allr = all_gather_rewards(rewards)
    g = O.fd_partial_gradient(noise.table, idx, sign, rewards, allr, 0.02, P, baseline=0.1)
    t = torch.from_numpy(g)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)                 # the one parameter-sized exchange
    ref = O.fd_gradient_closed_form(noise.table, np.concatenate([idx_all, idx_all]),
                                    np.concatenate([np.ones(R), -np.ones(R)]),
                                    np.concatenate([rew_plus, rew_minus]), 0.02, P, baseline=0.1)
    ok = np.max(np.abs(t.numpy() - ref)) <= 1e-12 * np.max(np.abs(ref))
    same_set = np.allclose(np.sort(allr), np.sort(np.concatenate([rew_plus, rew_minus])))
    out[rank] = bool(ok and same_set and allr.shape[0] == 2 * R)
    dist.destroy_process_group()


def test_sharded_estimator_world2():
    world, R, P = 2, 37, 513                                  # odd pair count: ragged shards
    assert [shard_pairs(R, r, world) for r in range(world)] == [(0, 19), (19, 37)]
    assert [shard_pairs(8, r, 3) for r in range(3)] == [(0, 3), (3, 6), (6, 8)]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), R, P, out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}
