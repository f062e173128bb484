#!/bin/bash
# fd_state / dots-pass parity tests + the C5 step (kernel times) + IMPALA parity
mkdir -p gpurun_out
T=${1:-c5q}
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_direct.py tests/test_gpu_round2.py -q -x -k "fd_state or impala or dots or estimator_steps_golden" --timeout 600 2>&1 | tail -4 | tee gpurun_out/${T}_test.log
timeout 600 python bench.py --workload C5 --steps 30 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python - gpurun_out/${T}_bench.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("C5 ms/step %.4f" % d["ms_per_step"], {k: round(v["us"], 1) for k, v in d["kernels"].items()}, "parity", d["parity"])
PY
