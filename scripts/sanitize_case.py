"""Tiny invocations of the hand-indexed kernels for compute-sanitizer (one tool per gpurun call):
smoke() (small exact MLP forward + one-kernel learner step) and the Atari forward (TMA-fed first Linear)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as G
G.smoke()
import dfd_starter_b200 as D
from oracle import dfd_oracle as O
L = O.atari_layout(6)
table = D.SharedNoiseTable(1_000_000, L.num_params, 123, device=0)
pol = D.AtariPolicy((84, 84), 6, seed=3, device=0).bind_table(table)
idx = torch.from_numpy(table.sample_indices(2)).cuda()
sign = torch.tensor([1, -1], dtype=torch.int8).cuda()
obs = torch.rand(2, 1, 4, 84, 84).cuda()
out = pol.forward_members(idx, sign, obs, 0.02)
torch.cuda.synchronize()
print("atari ok", out.sum().item())
