#!/bin/bash
# round-end check: full GPU suite, smoke, default bench, then the ncu evidence for the default workload
# (launch list of the plain-launch step + one full capture of the dominant kernel), each after its plain run exited 0
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as G; G.build(); G.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 600 gpurun_out/bench_default.err
A="--steps 3 --warmup 3 --graph off --profile-mode"
timeout 300 python bench.py --workload C2 $A > gpurun_out/plain_c2.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_c2.csv python bench.py --workload C2 $A > gpurun_out/ncu_c2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_forward_ws -s 12 -c 1 -o gpurun_out/prof_fwd_c2 -f python bench.py --workload C2 $A > gpurun_out/ncu_full_fwd.log 2>&1
tail -n 2 gpurun_out/ncu_c2.log gpurun_out/ncu_full_fwd.log
