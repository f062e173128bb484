#!/bin/bash
# A/B of the in-kernel exchange of the one-kernel learner step at N GPUs: low-latency packets (default) vs the flag protocol
N=${1:-8}
mkdir -p gpurun_out
for mode in ll flags ll flags; do
  if [ $mode = flags ]; then export DFD_TAIL_FLAGS=1; else unset DFD_TAIL_FLAGS; fi
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 40 --warmup 5 --workload C2 --no-e2e --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$mode', 'N', d['n_gpus'], 'ms/step %.4f' % d['ms_per_step'], 'parity', d['parity']['ok'])"
done
