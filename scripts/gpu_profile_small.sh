#!/bin/bash
# ncu full captures of the small per-step kernels (coefficient kernel, DSGD update, synthetic return)
mkdir -p gpurun_out
A="--steps 3 --warmup 3 --graph off --profile-mode"
python bench.py --workload C2 $A > gpurun_out/plain_small.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"fd_coef_kernel|dsgd_update_kernel|synthetic_reward_kernel|sumsq_partial_kernel" -s 40 -c 4 -o gpurun_out/prof_small_c2 -f python bench.py --workload C2 $A > gpurun_out/ncu_small.log 2>&1
tail -n 2 gpurun_out/ncu_small.log
