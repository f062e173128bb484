#!/bin/bash
# round-end check: full GPU suite, smoke, default bench, C5 bench, then the ncu evidence for C5 (launch list of the
# plain-launch step + one full capture of the forward), each after its plain run exited 0
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as G; G.build(); G.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 300 gpurun_out/bench_default.err
timeout 300 python bench.py --workload C5 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_C5.json 2> gpurun_out/bench_C5.err
A="--steps 3 --warmup 3 --graph off --profile-mode"
timeout 200 python bench.py --workload C5 $A > gpurun_out/plain_c5.log 2>&1 &&
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_c5.csv python bench.py --workload C5 $A > gpurun_out/ncu_c5.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:impala_forward -s 4 -c 1 -o gpurun_out/prof_fwd_c5 -f python bench.py --workload C5 $A > gpurun_out/ncu_full_fwd_c5.log 2>&1
python - <<'PY'
import json
for f in ("gpurun_out/bench_default.json", "gpurun_out/bench_C5.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms/step %.4f" % d["ms_per_step"], "value %.3e" % d["value"], {k: round(v["us"], 1) for k, v in d["kernels"].items()},
              "e2e", d.get("e2e") and round(d["e2e"]["ms_per_step"], 3), "frac %.3f" % d["roofline"]["frac"], d["clocks"])
    except Exception as e:
        print(f, "unreadable", e); print(open(f[:-4] + "err").read()[-1500:])
PY
tail -n 2 gpurun_out/ncu_c5.log gpurun_out/ncu_full_fwd_c5.log
