"""Per-CTA cycle timeline of the direct-from-table forward (DFD_DR_PROF=1), C3 shape."""
import os, sys
os.environ["DFD_DR_PROF"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dfd_starter_b200 as D
n_in, h, n_act, M, E = 376, 256, 17, 2048, 128
P = n_in * h + h + h * h + h + h * 2 * n_act + 2 * n_act
table = D.SharedNoiseTable(25_000_000, P, 124, device=0)
pol = D.MujocoPolicy(n_in, n_act, seed=3, h1=h, h2=h, device=0, precision=2).bind_table(table)
i = table.sample_indices(M // 2)
idx = torch.from_numpy(np.concatenate([i, i])).cuda()
sign = torch.from_numpy(np.concatenate([np.ones(M // 2), -np.ones(M // 2)]).astype(np.int8)).cuda()
obs = torch.randn(M, E, n_in, device="cuda")
for _ in range(2):
    pol.forward_members(idx, sign, obs, 0.02)
torch.cuda.synchronize()
