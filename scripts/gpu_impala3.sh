#!/bin/bash
# tcgen05 IMPALA trunk (level 3): parity tests, then timing of levels 3 and 2, then the per-phase timeline
mkdir -p gpurun_out
T=${1:-i3}
timeout 600 python -m pytest tests/test_gpu_direct.py -q -x -k "impala" --timeout 300 > gpurun_out/${T}_test.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_test.log
tail -25 gpurun_out/${T}_test.log
timeout 300 python scripts/impala_bench.py 3 2 2>&1 | tail -4 | tee gpurun_out/${T}_bench.log
DFD_IMPALA_PROF=1 timeout 300 python scripts/impala_bench.py 3 2>&1 | grep timeline | tail -6 | tee gpurun_out/${T}_prof.log
