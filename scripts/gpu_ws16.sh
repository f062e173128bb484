#!/bin/bash
mkdir -p gpurun_out
DFD_WS16=1 timeout 300 python -m pytest tests/test_gpu_tensor_core.py tests/test_gpu_direct.py -q -x --timeout 200 -k "halfcheetah or small_net or resident" > gpurun_out/ws16_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/ws16_pytest.log
tail -5 gpurun_out/ws16_pytest.log
DFD_WS16=1 DFD_W6_PROF=1 timeout 200 python scripts/fwd_bench.py C2 128 2>&1 | grep "timeline" | tail -1
DFD_WS16=1 timeout 200 python scripts/fwd_bench.py C2 128 2>&1 | tail -2
timeout 200 python scripts/fwd_bench.py C2 128 2>&1 | tail -2
