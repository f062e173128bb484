#!/bin/bash
mkdir -p gpurun_out
A="--steps 3 --warmup 3 --graph off --profile-mode"
python bench.py --workload C5 $A > gpurun_out/plain_c5d.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fd_dots -s 6 -c 1 -o gpurun_out/prof_dots_c5 -f python bench.py --workload C5 $A > gpurun_out/ncu_dots.log 2>&1
tail -2 gpurun_out/ncu_dots.log
