#!/bin/bash
# round-2 ncu evidence for profiles/r02: launch lists of the plain-launch bench step of every workload + full captures of
# the new dominant kernels.  Every profiled command first runs plain with the same arguments and must exit 0.
mkdir -p gpurun_out
A="--steps 3 --warmup 3 --graph off --profile-mode"
for W in C2 C3 C4 C5; do
  w=$(echo $W | tr 'A-Z' 'a-z')
  python bench.py --workload $W $A > gpurun_out/plain_$w.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$w.csv python bench.py --workload $W $A > gpurun_out/ncu_$w.log 2>&1
  echo "$W launch list rc=$?"
done
python bench.py --workload C3 $A > gpurun_out/plain_c3b.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:mlp_forward_direct -s 6 -c 1 -o gpurun_out/prof_fwd_c3 -f python bench.py --workload C3 $A > gpurun_out/ncu_full_fwd_c3.log 2>&1
python bench.py --workload C4 $A > gpurun_out/plain_c4b.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:atari_forward_tc -s 6 -c 1 -o gpurun_out/prof_fwd_c4 -f python bench.py --workload C4 $A > gpurun_out/ncu_full_fwd_c4.log 2>&1
python bench.py --workload C5 $A > gpurun_out/plain_c5b.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:impala_forward_kernel -s 6 -c 1 -o gpurun_out/prof_fwd_c5 -f python bench.py --workload C5 $A > gpurun_out/ncu_full_fwd_c5.log 2>&1
python bench.py --workload C5 $A > gpurun_out/plain_c5c.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fd_reduce -s 6 -c 1 -o gpurun_out/prof_reduce_c5 -f python bench.py --workload C5 $A > gpurun_out/ncu_full_reduce_c5.log 2>&1
ls -la gpurun_out/*.ncu-rep
tail -n 2 gpurun_out/ncu_full_fwd_c3.log gpurun_out/ncu_full_fwd_c4.log gpurun_out/ncu_full_fwd_c5.log gpurun_out/ncu_full_reduce_c5.log
