#!/bin/bash
# One gpurun call: GPU parity tests, smoke, short benches. Everything is logged under gpurun_out/.
# usage: gpu_check.sh ["C2 128 200 auto" "C3 16 20 fp32" ...]   (workload, obs per member, steps, precision)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/pytest_gpu.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as G; G.build(); G.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit: $?" >> gpurun_out/smoke.log
if [ $# -eq 0 ]; then set -- "C2 128 200 auto" "C2 1 200 auto" "C3 128 20 auto"; fi
for spec in "$@"; do
  set -- $spec
  tag=$1_E$2_$4
  timeout 900 python bench.py --workload $1 --obs-per-member $2 --steps $3 --warmup 5 --precision $4 --no-cpu-baseline > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
  echo "bench $tag exit: $?" >> gpurun_out/bench_$tag.err
done
tail -25 gpurun_out/pytest_gpu.log; tail -n 3 gpurun_out/smoke.log
for f in gpurun_out/bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], 'ms/step %.4f'%d['ms_per_step'], 'value %.3e'%d['value'], {k:round(v['us'],1) for k,v in d['kernels'].items()}, 'reduce GB/s %.0f frac %.3f'%(d['roofline']['fd_reduce']['achieved'], d['roofline']['fd_reduce']['frac']), 'e2e', d['e2e'] and round(d['e2e']['ms_per_step'],3))
except Exception as e:
    print(sys.argv[1], 'unreadable', e)
PY
done
for f in gpurun_out/*.err; do tail -n 2 $f; done
