#!/bin/bash
# ncu evidence for profiles/: launch list of the bench step + full captures of the top kernels.
# Each profiled command first runs plain (same args) and must exit 0.
mkdir -p gpurun_out
A="--steps 3 --warmup 3 --graph off --profile-mode"
python bench.py --workload C2 $A > gpurun_out/plain_c2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_c2.csv python bench.py --workload C2 $A > gpurun_out/ncu_c2.log 2>&1
python bench.py --workload C3 --obs-per-member 128 $A > gpurun_out/plain_c3.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_c3.csv python bench.py --workload C3 --obs-per-member 128 $A > gpurun_out/ncu_c3.log 2>&1
# full captures: reduction at C3 size (HBM-bound), forward at C2 (resident-weight tcgen05 kernel) and C3 (streaming)
ncu --set full --clock-control none --import-source on -k regex:fd_reduce -s 12 -c 1 -o gpurun_out/prof_reduce_c3 -f python bench.py --workload C3 --obs-per-member 128 $A > gpurun_out/ncu_full_reduce.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mlp_forward_ws -s 12 -c 1 -o gpurun_out/prof_fwd_c2 -f python bench.py --workload C2 $A > gpurun_out/ncu_full_fwd.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mlp_forward_stream -s 12 -c 1 -o gpurun_out/prof_fwd_c3 -f python bench.py --workload C3 --obs-per-member 128 $A > gpurun_out/ncu_full_fwd_c3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fd_tail -s 6 -c 1 -o gpurun_out/prof_tail_c2 -f python bench.py --workload C2 $A > gpurun_out/ncu_full_tail_c2.log 2>&1
ls -la gpurun_out/*.ncu-rep
tail -n 2 gpurun_out/ncu_c2.log gpurun_out/ncu_full_reduce.log gpurun_out/ncu_full_fwd.log gpurun_out/ncu_full_fwd_c3.log
