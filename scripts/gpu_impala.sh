#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_direct.py -q -x --timeout 300 -k impala -s > gpurun_out/impala_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/impala_pytest.log
tail -30 gpurun_out/impala_pytest.log
DFD_IMPALA_PROF=1 timeout 300 python scripts/impala_bench.py 3 > gpurun_out/impala_fwd.log 2>&1; echo "rc=$?" >> gpurun_out/impala_fwd.log
timeout 300 python scripts/impala_bench.py 2 3 1 >> gpurun_out/impala_fwd.log 2>&1
grep -v "^$" gpurun_out/impala_fwd.log | tail -12
