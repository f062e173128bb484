#!/usr/bin/env python
"""dfd_fd_reduce on rows that cannot be L2-resident (disjoint rows of a multi-GB buffer, NSET row sets cycled), one
staging plan per process (DFD_TMA_PLAN="vec,rb,stages,ctas_per_sm,min_rows" or the built-in plan).  Prints one JSON line
per shape: achieved GB/s of algorithmic bytes, fraction of the measured HBM peak, max relative error vs an fp64 matvec.
usage: python scripts/reduce_cold.py [C3 C5 C4 ...]"""
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dfd_starter_b200 import _lib                           # noqa: E402
from dfd_starter_b200.device import get_context, ptr, aligned_ptr  # noqa: E402

SHAPES = {"C3": (171042, 1024, 6), "C5": (1158709, 266, 4), "C4": (678294, 512, 4), "C3x64": (30498, 1024, 8)}
peak = 6545.9
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
ctx = get_context(0)
lib, dev = ctx.lib, ctx.device
for name in (sys.argv[1:] or ["C3", "C5"]):
    Pb, Rb, NSET = SHAPES[name]
    Pb4 = (Pb + 3) // 4 * 4
    cold = torch.randn(NSET * Rb * Pb4, device=dev)
    gb = torch.empty(Pb, device=dev)
    sb = ctx.zeros_bytes(lib.dfd_fd_reduce_scratch_bytes(ctx.handle, Pb, Rb))
    sets = []
    for c in range(NSET):
        rp = (cold.data_ptr() + 4 * Pb4 * (c * Rb + torch.arange(Rb, dtype=torch.int64))).to(dev)
        rc = torch.randn(Rb, device=dev)
        sets.append((rp, rc, _lib.DfdFdRows(rp.data_ptr(), rc.data_ptr(), Rb)))

    def k(r):
        _lib.check(lib.dfd_fd_reduce(ctx.handle, C.byref(sets[r % NSET][2]), Rb, Pb, ptr(gb), aligned_ptr(sb),
                                     sb.numel() - 256, ctx.stream))
    k(0)
    want = sets[0][1].double() @ cold[:Rb * Pb4].view(Rb, Pb4)[:, :Pb].double()
    err = float((gb.double() - want).abs().max() / want.abs().max())
    reps = 3 * NSET
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for r in range(reps):
            k(r)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(4):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.replay()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / reps * 1e3)
    nbytes = Rb * Pb * 4 + Pb * 4
    print(json.dumps({"shape": name, "plan": os.environ.get("DFD_TMA_PLAN", "built-in"), "us": round(best, 2),
                      "GBps": round(nbytes / best * 1e-3, 1), "frac": round(nbytes / best * 1e-3 / peak, 4), "rel_err": err}),
          flush=True)
    del cold, sets, g
    torch.cuda.empty_cache()
