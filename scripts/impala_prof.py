"""C5 forward alone: kernel time (CUDA events, distinct noise rows per launch) and, with DFD_IMPALA_PROF=1, the phase
timeline of one CTA.  python scripts/impala_prof.py [members]"""
import os
import sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dfd_starter_b200 as D

M = int(sys.argv[1]) if len(sys.argv) > 1 else 512
prec = int(sys.argv[2]) if len(sys.argv) > 2 else 0
P = 1158709
table = D.SharedNoiseTable(25_000_000, P, 124, device=0)
pol = D.ImpalaPolicy((3, 64, 64), 15, seed=124, device=0, precision=prec).bind_table(table)
sign = torch.from_numpy(np.concatenate([np.ones(M // 2), -np.ones(M // 2)]).astype(np.int8)).cuda()
frames = torch.randint(0, 256, (M, 1, 3, 64, 64), device="cuda").float()
rew = torch.zeros(M, 1, device="cuda")
done = torch.zeros(M, 1, dtype=torch.bool, device="cuda")
h = torch.zeros(M, 1, 256, device="cuda")
c = torch.zeros(M, 1, 256, device="cuda")
idxs = []
for k in range(6):
    i = table.sample_indices(M // 2)
    idxs.append(torch.from_numpy(np.concatenate([i, i])).cuda())
prof = os.environ.pop("DFD_IMPALA_PROF", None)
for k in range(2):
    pol.forward_members_impala(idxs[k], sign, frames, rew, done, h, c, 0.02)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
ev[0].record()
for k in range(4):
    pol.forward_members_impala(idxs[2 + k], sign, frames, rew, done, h, c, 0.02)
    ev[k + 1].record()
torch.cuda.synchronize()
print("forward us:", [round(ev[k].elapsed_time(ev[k + 1]) * 1000, 1) for k in range(4)])
if prof:
    os.environ["DFD_IMPALA_PROF"] = "1"
    pol.forward_members_impala(idxs[0], sign, frames, rew, done, h, c, 0.02)
    torch.cuda.synchronize()
