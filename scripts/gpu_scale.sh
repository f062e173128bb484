#!/bin/bash
# bench at N GPUs of one box (torchrun, one rank per GPU); prints the per-workload summary
N=${1:-2}; T=${2:-scale}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${T}_n$N.json 2> gpurun_out/${T}_n$N.err; echo "rc=$?" >> gpurun_out/${T}_n$N.err
tail -3 gpurun_out/${T}_n$N.err
python - gpurun_out/${T}_n$N.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("N=%d C2 ms/step %.4f value %.3e e2e ms %.3f parity %s" % (d["n_gpus"], d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d.get("parity")))
print("host", d.get("host"))
for k, w in d.get("workloads", {}).items():
    print(k, "ms/step %.4f fwd %.1f us reduce %.1f us e2e ms %.3f parity %s" % (w["ms_per_step"], w["forward_us"], w["reduce_us"], w["e2e"]["ms_per_step"], w["parity"]))
PY
