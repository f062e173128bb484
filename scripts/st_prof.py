import numpy as np, torch, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dfd_starter_b200 as D
E = int(sys.argv[1]) if len(sys.argv) > 1 else 128
prec = int(sys.argv[2]) if len(sys.argv) > 2 else 2
P = 376 * 256 + 256 + 256 * 256 + 256 + 256 * 34 + 34
table = D.SharedNoiseTable(25_000_000, P, 124, device=0)
pol = D.MujocoPolicy(376, 17, seed=3, h1=256, h2=256, device=0, precision=prec).bind_table(table)
M = 2048
i = table.sample_indices(M // 2)
idx = torch.from_numpy(np.concatenate([i, i])).cuda()
sign = torch.from_numpy(np.concatenate([np.ones(M // 2), -np.ones(M // 2)]).astype(np.int8)).cuda()
obs = torch.randn(M, E, 376, device='cuda')
for k in range(3):
    out = pol.forward_members(idx, sign, obs, 0.02)
torch.cuda.synchronize()
