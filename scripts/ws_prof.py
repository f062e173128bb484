import numpy as np, torch, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dfd_starter_b200 as D
E = int(sys.argv[1]) if len(sys.argv) > 1 else 128
prec = int(sys.argv[2]) if len(sys.argv) > 2 else 1
table = D.SharedNoiseTable(25_000_000, 6092, 124, device=0)
pol = D.MujocoPolicy(17, 6, seed=3, device=0, precision=prec).bind_table(table)
M = 2048
idx = torch.from_numpy(table.sample_indices(M)).cuda(); sign = torch.ones(M, dtype=torch.int8, device='cuda')
obs = torch.randn(M, E, 17, device='cuda')
for i in range(3):
    out = pol.forward_members(idx, sign, obs, 0.02)
torch.cuda.synchronize()
