#!/bin/bash
# the direct-from-table forward (C3): parity tests, then forward-only timing against the streaming kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py -q -x -k "humanoid" --timeout 300 > gpurun_out/direct_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/direct_pytest.log
tail -15 gpurun_out/direct_pytest.log
timeout 300 python scripts/fwd_bench.py C3 128 > gpurun_out/direct_fwd.log 2>&1; echo "rc=$?" >> gpurun_out/direct_fwd.log
DFD_TC_NO_DIRECT=1 timeout 300 python scripts/fwd_bench.py C3 128 >> gpurun_out/direct_fwd.log 2>&1
cat gpurun_out/direct_fwd.log
