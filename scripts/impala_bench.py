"""Forward-only timing of the IMPALA forward at one GPU's share of BASELINE config 5 (256 pairs, E = 1) per precision level."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dfd_starter_b200 as D
M, E, P = 512, 1, 1158709
table = D.SharedNoiseTable(25_000_000, P, 124, device=0)
levels = [int(x) for x in sys.argv[1:]] or [2, 1]
for prec in levels:
    pol = D.ImpalaPolicy((3, 64, 64), 15, seed=124, device=0, precision=prec).bind_table(table)
    sets = []
    for c in range(3):
        i = table.sample_indices(M // 2)
        sets.append((torch.from_numpy(np.concatenate([i, i])).cuda(), torch.randint(0, 256, (M, E, 3, 64, 64), device="cuda").float()))
    sign = torch.from_numpy(np.concatenate([np.ones(M // 2), -np.ones(M // 2)]).astype(np.int8)).cuda()
    rew = torch.zeros(M, E, device="cuda"); done = torch.zeros(M, E, dtype=torch.bool, device="cuda")
    h = torch.zeros(M, E, 256, device="cuda"); c0 = torch.zeros(M, E, 256, device="cuda")
    for c in range(3):
        pol.forward_members_impala(sets[c][0], sign, sets[c][1], rew, done, h, c0, 0.02)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 9
    a.record()
    for r in range(reps):
        pol.forward_members_impala(sets[r % 3][0], sign, sets[r % 3][1], rew, done, h, c0, 0.02)
    b.record(); torch.cuda.synchronize()
    us = a.elapsed_time(b) / reps * 1e3
    print("C5 share E=%d precision=%d: %.1f us, %.0f GB/s of pairs*P*4" % (E, prec, us, (M // 2) * P * 4 / us * 1e-3), flush=True)
