#!/bin/bash
# direct-from-table MLP forward (C3): parity tests + the C3 step
mkdir -p gpurun_out
T=${1:-c3q}
timeout 900 python -m pytest tests/test_gpu_direct.py tests/test_gpu_round2.py tests/test_gpu_tensor_core.py -q -x -k "not impala and not atari" --timeout 600 2>&1 | tail -4 | tee gpurun_out/${T}_test.log
timeout 600 python bench.py --workload C3 --steps 30 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python - gpurun_out/${T}_bench.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("C3 ms/step %.4f" % d["ms_per_step"], {k: round(v["us"], 1) for k, v in d["kernels"].items()}, "parity ok", d["parity"]["ok"], d["parity"]["grad_rel_max"])
PY
DFD_DR_PROF=1 timeout 300 python scripts/dr_prof.py 2>&1 | grep "direct timeline" | head -2 | cut -c1-400
