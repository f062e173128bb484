#!/bin/bash
# N-GPU sanity after kernel changes: the multi-GPU parity tests, then the bench line at N (C2 + workloads map)
N=${1:-2}; T=${2:-mq}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -q -x --timeout 500 2>&1 | tail -3 | tee gpurun_out/${T}_multi_test.log
bash scripts/gpu_scale.sh $N $T
