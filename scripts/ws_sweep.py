"""Forward time vs. members per SM (fixed cost vs. per-item cost of mlp_forward_ws_kernel), CUDA-graph replay."""
import numpy as np, torch, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dfd_starter_b200 as D
E, prec = 128, 2
table = D.SharedNoiseTable(25_000_000, 6092, 124, device=0)
pol = D.MujocoPolicy(17, 6, seed=3, device=0, precision=prec).bind_table(table)
for per_sm in (1, 2, 3, 4, 6, 9, 12, 14):
    M = 148 * per_sm
    sets = []
    for c in range(8):
        ix = table.sample_indices(M // 2)
        sets.append((torch.from_numpy(np.concatenate([ix, ix])).cuda(), torch.randn(M, E, 17, device='cuda')))
    sign = torch.cat([torch.ones(M // 2, dtype=torch.int8), -torch.ones(M // 2, dtype=torch.int8)]).cuda()
    out = torch.empty(M, E, 12, device='cuda')
    for c in range(8):
        pol.forward_members(sets[c][0], sign, sets[c][1], 0.02, out=out)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for r in range(32):
            pol.forward_members(sets[r % 8][0], sign, sets[r % 8][1], 0.02, out=out)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / 32 * 1e3)
    print("items per SM %2d  members %5d  %6.2f us per forward" % (per_sm, M, best), flush=True)
