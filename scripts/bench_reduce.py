#!/usr/bin/env python
"""Reduction-only microbenchmark: dfd_fd_reduce at the BASELINE config sizes, back-to-back launches over
10 different row sets (so a launch never finds its rows in L2 from the previous one), CUDA-event timed
inside a CUDA graph.  Prints achieved GB/s of ALGORITHMIC bytes (rows*P*4 + P*4) vs the measured HBM peak.
Set DFD_REDUCE_MODE=ldg|tma to pick the implementation."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dfd_starter_b200 as D                                # noqa: E402
from dfd_starter_b200 import _lib                           # noqa: E402
from dfd_starter_b200.device import get_context, ptr, aligned_ptr  # noqa: E402

peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0) if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
ctx = get_context(0)
lib = ctx.lib
dev = ctx.device
CONFIGS = [("C2", 6092, 1024), ("C3/64x64", 30498, 1024), ("C3", 171042, 1024), ("C4", 678294, 512), ("C5 shard", 1158709, 256)]
if len(sys.argv) > 1:
    CONFIGS = [c for c in CONFIGS if c[0].split()[0] in sys.argv[1:]]
table = D.SharedNoiseTable(25_000_000, 1158709, 124, device=0)
results = {}
for name, P, R in CONFIGS:
    NSET = 10
    rng = np.random.RandomState(1)
    sets = []
    grad = torch.empty(P, device=dev)
    scratch = ctx.zeros_bytes(lib.dfd_fd_reduce_scratch_bytes(ctx.handle, P, R))
    for c in range(NSET):
        idx = rng.randint(0, 25_000_000 - P, size=R).astype(np.int64)
        coef = torch.from_numpy(rng.randn(R).astype(np.float32)).to(dev)
        rp = torch.from_numpy(np.array([table.device_table.replicas.data_ptr() + 4 * ((i & 3) * table.device_table.stride + (i - (i & 3)))
                                        for i in idx], dtype=np.int64)).to(dev)
        sets.append((rp, coef, _lib.DfdFdRows(rp.data_ptr(), coef.data_ptr(), R), idx))

    def launch(c):
        _lib.check(lib.dfd_fd_reduce(ctx.handle, C.byref(sets[c][2]), R, P, ptr(grad), aligned_ptr(scratch),
                                     scratch.numel() - 256, ctx.stream))
    # correctness of this implementation against a torch fp64 gather-sum on one set
    launch(0)
    torch.cuda.synchronize()
    t = torch.from_numpy(table._table)
    idx0 = sets[0][3]
    ref = torch.zeros(P, dtype=torch.float64)
    cf = sets[0][1].cpu().double()
    for r in range(R):
        ref += cf[r] * t[idx0[r]:idx0[r] + P].double()
    err = float((grad.cpu().double() - ref).abs().max() / ref.abs().max())
    for c in range(NSET):
        launch(c)
    torch.cuda.synchronize()
    reps = 4 * NSET
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for r in range(reps):
            launch(r % NSET)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.replay()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / reps * 1e3)
    nbytes = R * P * 4 + P * 4
    results[name] = dict(P=P, rows=R, us=best, GBps=nbytes / best * 1e-3, frac=nbytes / best * 1e-3 / peak, rel_err=err)
    print("%-10s P=%8d rows=%5d  %8.1f us  %7.0f GB/s  frac %.3f of measured %.0f  (max rel err %.1e)"
          % (name, P, R, best, nbytes / best * 1e-3, nbytes / best * 1e-3 / peak, peak, err), flush=True)
print(json.dumps({"mode": os.environ.get("DFD_REDUCE_MODE", "auto"), "results": results}))
