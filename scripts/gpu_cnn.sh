#!/bin/bash
# CNN forwards: parity tests + C4 / C5 bench lines (kernel times)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "atari or impala" 2>&1 | tail -3
for wl in C4 C5; do
  timeout 200 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_${wl}_q.json 2> gpurun_out/bench_${wl}_q.err
  python - $wl <<'PY'
import json, sys
f = "gpurun_out/bench_%s_q.json" % sys.argv[1]
try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(sys.argv[1], "ms/step %.4f" % d["ms_per_step"], {k: round(v["us"], 1) for k, v in d["kernels"].items()})
except Exception as e:
    print(f, "unreadable", e); print(open(f[:-4] + "err").read()[-1500:])
PY
done
