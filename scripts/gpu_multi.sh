#!/bin/bash
# multi-GPU checks on N GPUs: exchange checks (all shapes) and the bench line at N
N=${1:-2}; T=${2:-multi}
mkdir -p gpurun_out
for shape in "" odd wide state; do
  XCHG_SHAPE=$shape timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 scripts/xchg_check.py > gpurun_out/${T}_xchg_${shape:-plain}_n$N.log 2>&1
  echo "shape=${shape:-plain} rc=$? $(grep 'xchg ok' gpurun_out/${T}_xchg_${shape:-plain}_n$N.log | tail -1)"
  grep -i "error\|assert" gpurun_out/${T}_xchg_${shape:-plain}_n$N.log | tail -5
done
bash scripts/gpu_scale.sh $N $T
