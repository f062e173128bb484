#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tensor_core.py -q -x --timeout 200 > gpurun_out/ws16_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/ws16_pytest.log
tail -15 gpurun_out/ws16_pytest.log
timeout 200 python scripts/fwd_bench.py C2 128 2>&1 | tail -3
DFD_NO_WS16=1 timeout 200 python scripts/fwd_bench.py C2 128 2>&1 | tail -3
