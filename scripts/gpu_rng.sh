#!/bin/bash
# device RNGNoiseSource: its own tests + every existing test that drives the learner / worker with that noise source
mkdir -p gpurun_out
T=${1:-rng}
timeout 900 python -m pytest tests/test_gpu_rng.py -q -x -s --timeout 600 > gpurun_out/${T}_rng.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_rng.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -q -k "noise_sources or server_train" --timeout 600 > gpurun_out/${T}_users.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_users.log
tail -15 gpurun_out/${T}_rng.log; tail -15 gpurun_out/${T}_users.log
timeout 600 python scripts/rng_bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "rng_bench rc=$?"; tail -5 gpurun_out/${T}_bench.err; cat gpurun_out/${T}_bench.json
if [ -n "$2" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:rng_ -c 60 --csv --log-file gpurun_out/${T}_launches.csv python scripts/rng_bench.py --rows-only > gpurun_out/${T}_ncu.log 2>&1; tail -2 gpurun_out/${T}_ncu.log
fi
