"""Forward timeline (DFD_WS_PROF=1) with cold table rows: fresh indices every call and an L2 flush in between."""
import numpy as np, torch, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dfd_starter_b200 as D
E = int(sys.argv[1]) if len(sys.argv) > 1 else 128
prec = int(sys.argv[2]) if len(sys.argv) > 2 else 2
table = D.SharedNoiseTable(25_000_000, 6092, 124, device=0)
pol = D.MujocoPolicy(17, 6, seed=3, device=0, precision=prec).bind_table(table)
M = 2048
sign = torch.cat([torch.ones(M // 2, dtype=torch.int8), -torch.ones(M // 2, dtype=torch.int8)]).cuda()
flush = torch.empty(64 << 20, device='cuda')
for i in range(3):
    ix = table.sample_indices(M // 2)
    idx = torch.from_numpy(np.concatenate([ix, ix])).cuda()
    obs = torch.randn(M, E, 17, device='cuda')
    flush.fill_(float(i))
    torch.cuda.synchronize()
    print("---- call", i, file=sys.stderr, flush=True)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    out = pol.forward_members(idx, sign, obs, 0.02)
    b.record()
    torch.cuda.synchronize()
    print("call %d: %.1f us (includes the timeline read-back when profiling)" % (i, a.elapsed_time(b) * 1e3), file=sys.stderr)
