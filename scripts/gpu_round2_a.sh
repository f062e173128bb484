#!/bin/bash
# first GPU pass of round 2: all GPU tests, smoke, default bench (with workloads map), reference arm
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
timeout 300 python -c "import __graft_entry__ as G; G.smoke()" > gpurun_out/r2a_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2a_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?" >> gpurun_out/r2a_bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2a_ref.json 2> gpurun_out/r2a_ref.err; echo "ref rc=$?" >> gpurun_out/r2a_ref.err
tail -3 gpurun_out/r2a_pytest.log; tail -2 gpurun_out/r2a_smoke.log; tail -2 gpurun_out/r2a_bench.err; head -c 600 gpurun_out/r2a_bench.json
