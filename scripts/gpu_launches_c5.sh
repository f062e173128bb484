#!/bin/bash
mkdir -p gpurun_out
A="--steps 3 --warmup 3 --graph off --profile-mode"
python bench.py --workload C5 $A > gpurun_out/plain_c5.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c5.csv python bench.py --workload C5 $A > gpurun_out/ncu_c5.log 2>&1
echo rc=$?
