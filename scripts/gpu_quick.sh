#!/bin/bash
# quick loop: tensor-core tests + named benches (no full suite)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tensor_core.py -q 2>&1 | tail -5
for spec in "$@"; do
  set -- $spec
  tag=$1_E$2_$4
  timeout 900 python bench.py --workload $1 --obs-per-member $2 --steps $3 --warmup 5 --precision $4 --no-cpu-baseline --no-e2e > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
  python - gpurun_out/bench_$tag.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], 'ms/step %.4f'%d['ms_per_step'], 'value %.3e'%d['value'], {k:round(v['us'],1) for k,v in d['kernels'].items()}, 'reduce frac %.3f'%(d['roofline']['fd_reduce']['frac']))
except Exception as e:
    print(sys.argv[1], 'unreadable', e); print(open(sys.argv[1][:-4]+'err').read()[-1500:])
PY
done
