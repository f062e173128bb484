"""Sharded learner over >= 2 GPUs of one box: the one-kernel NVLink peer-memory exchange (dist.PeerExchange,
csrc/xchg_allreduce.cu) against the NCCL path and the CPU oracle of the unsharded batch.  The check itself is
scripts/xchg_check.py (a torchrun script: one process per GPU); skipped on single-GPU boxes."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.timeout(300)
@pytest.mark.parametrize("shape", ["", "odd", "wide", "state"])
def test_peer_exchange_matches_nccl_and_oracle(shape):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import __graft_entry__ as G
    G.build()
    n = min(torch.cuda.device_count(), 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "scripts", "xchg_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=280, cwd=ROOT, env=dict(os.environ, XCHG_SHAPE=shape))
    assert out.returncode == 0 and "xchg ok" in out.stdout, (out.stdout[-2000:], out.stderr[-2000:])
