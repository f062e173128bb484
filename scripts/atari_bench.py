"""Forward-only timing of the Atari forward at BASELINE config 4 size (512 pairs, E = 1): tensor path vs exact fp32."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dfd_starter_b200 as D
M, E, P = 1024, int(sys.argv[1]) if len(sys.argv) > 1 else 1, 678294
table = D.SharedNoiseTable(25_000_000, P, 124, device=0)
for prec in (1, 0):
    pol = D.AtariPolicy((84, 84), 6, seed=124, device=0, precision=prec).bind_table(table)
    sets = []
    for c in range(3):
        i = table.sample_indices(M // 2)
        sets.append((torch.from_numpy(np.concatenate([i, i])).cuda(), torch.rand(M, E, 4, 84, 84, device="cuda")))
    sign = torch.from_numpy(np.concatenate([np.ones(M // 2), -np.ones(M // 2)]).astype(np.int8)).cuda()
    out = torch.empty(M, E, 6, device="cuda")
    for c in range(3):
        pol.forward_members(sets[c][0], sign, sets[c][1], 0.02, out=out)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 12
    a.record()
    for r in range(reps):
        pol.forward_members(sets[r % 3][0], sign, sets[r % 3][1], 0.02, out=out)
    b.record(); torch.cuda.synchronize()
    us = a.elapsed_time(b) / reps * 1e3
    print("C4 E=%d precision=%d: %.1f us (incl. theta->fp16 kernel), %.0f GB/s of pairs*P*4" % (E, prec, us, (M // 2) * P * 4 / us * 1e-3), flush=True)
