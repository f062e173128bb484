#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_direct.py -q -x --timeout 300 > gpurun_out/atari_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/atari_pytest.log
tail -25 gpurun_out/atari_pytest.log
timeout 300 python scripts/atari_bench.py 1 > gpurun_out/atari_fwd.log 2>&1; echo "rc=$?" >> gpurun_out/atari_fwd.log
cat gpurun_out/atari_fwd.log
