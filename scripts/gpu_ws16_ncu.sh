#!/bin/bash
mkdir -p gpurun_out
export DFD_WS16=1
python scripts/fwd_bench.py C2 128 > gpurun_out/plain_ws16.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:mlp_forward_ws16 -s 4 -c 1 -o gpurun_out/prof_fwd_c2_ws16 -f python scripts/fwd_bench.py C2 128 > gpurun_out/ncu_ws16.log 2>&1
tail -3 gpurun_out/ncu_ws16.log
