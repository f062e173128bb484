#!/bin/bash
# C5 ncu evidence after the tcgen05 trunk became the default: launch list of the plain-launch step + full captures of the
# forward and the dots pass (each after its plain run exited 0)
mkdir -p gpurun_out
A="--steps 3 --warmup 3 --graph off --profile-mode"
timeout 300 python bench.py --workload C5 $A > gpurun_out/plain_c5.log 2>&1 &&
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_c5.csv python bench.py --workload C5 $A > gpurun_out/ncu_c5.log 2>&1
echo "launch list rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:impala_direct -s 4 -c 1 -o gpurun_out/prof_fwd_c5 -f python bench.py --workload C5 $A > gpurun_out/ncu_full_fwd_c5.log 2>&1
echo "fwd capture rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:fd_dots_epoch -s 4 -c 1 -o gpurun_out/prof_dots_c5 -f python bench.py --workload C5 $A > gpurun_out/ncu_full_dots_c5.log 2>&1
echo "dots capture rc=$?"
ls -la gpurun_out/*.ncu-rep; tail -n 2 gpurun_out/ncu_full_fwd_c5.log
