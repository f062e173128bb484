#!/bin/bash
# full GPU pass: all GPU tests, smoke, default bench (with workloads map)
mkdir -p gpurun_out
T=${1:-r2}
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
timeout 300 python -c "import __graft_entry__ as G; G.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${T}_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?" >> gpurun_out/${T}_bench.err
tail -3 gpurun_out/${T}_pytest.log; tail -2 gpurun_out/${T}_smoke.log; tail -3 gpurun_out/${T}_bench.err
python - gpurun_out/${T}_bench.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("C2 ms/step %.4f value %.3e e2e %.3e frac %.3f" % (d["ms_per_step"], d["value"], d["e2e"]["value"], d["roofline"]["frac"]))
for k, w in d.get("workloads", {}).items():
    print(k, "ms/step %.4f fwd %.1f us reduce %.1f us frac %.3f parity %s e2e ms %.3f" % (w["ms_per_step"], w["forward_us"], w["reduce_us"], w["roofline"]["frac"], w["parity"]["ok"], w["e2e"]["ms_per_step"]))
PY
