"""Forward-only timing of the per-member MLP forward (CUDA events, graph replay of `reps` launches that
cycle through different index / observation sets).  usage: fwd_bench.py [C2|C3] [E ...]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dfd_starter_b200 as D

def run(shape, E, precision, M=2048, reps=20, nset=5):
    n_in, h, n_act = (17, 64, 6) if shape == "C2" else (376, 256, 17)
    P = n_in * h + h + h * h + h + h * 2 * n_act + 2 * n_act
    table = run.tables.get(P) or D.SharedNoiseTable(25_000_000, P, 124, device=0)
    run.tables[P] = table
    pol = D.MujocoPolicy(n_in, n_act, seed=3, h1=h, h2=h, device=0, precision=precision).bind_table(table)
    R = M // 2
    sets = []
    for c in range(nset):
        i = table.sample_indices(R)
        sets.append((torch.from_numpy(np.concatenate([i, i])).cuda(), torch.randn(M, E, n_in, device="cuda")))
    sign = torch.from_numpy(np.concatenate([np.ones(R), -np.ones(R)]).astype(np.int8)).cuda()
    out = torch.empty(M, E, 2 * n_act, device="cuda")
    for c in range(nset):
        pol.forward_members(sets[c][0], sign, sets[c][1], 0.02, out=out)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for r in range(reps):
            pol.forward_members(sets[r % nset][0], sign, sets[r % nset][1], 0.02, out=out)
    g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g.replay(); b.record(); torch.cuda.synchronize()
    us = a.elapsed_time(b) / reps * 1e3
    flops = M * E * 2 * (n_in * h + h * h + h * 2 * n_act)
    byts = R * P * 4 + M * E * (n_in + 2 * n_act) * 4
    print("%s E=%d precision=%d: %.1f us  %.1f TFLOP/s  %.0f GB/s" % (shape, E, precision, us, flops / us * 1e-6, byts / us * 1e-3), flush=True)
run.tables = {}

if __name__ == "__main__":
    shape = sys.argv[1] if len(sys.argv) > 1 else "C2"
    Es = [int(x) for x in sys.argv[2:]] or [128]
    for E in Es:
        for prec in (1, 2):
            run(shape, E, prec)
