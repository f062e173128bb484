#!/bin/bash
# sweep staging plans of the TMA reduction on cold rows; one process per plan (the override is read once per process)
out=gpurun_out/reduce_sweep.jsonl; : > $out
python scripts/reduce_cold.py C3 C5 C4 >> $out 2>gpurun_out/reduce_sweep.err
for plan in "4,1,10,4,32" "4,1,10,4,64" "4,1,12,4,128" "4,2,5,4,64" "4,2,6,4,128" "4,4,3,4,64" "4,4,3,4,128" "4,4,4,3,128" \
            "8,1,6,4,64" "8,1,6,4,128" "8,2,3,4,64" "8,2,3,4,128" "8,2,4,3,128" "8,4,2,3,128" "8,1,8,3,128" \
            "16,1,3,4,128" "16,1,4,3,128" "16,2,2,3,128" "16,1,6,2,128" "8,2,6,2,128" "8,4,3,2,128" "4,4,6,2,128"; do
  DFD_TMA_PLAN=$plan timeout 120 python scripts/reduce_cold.py C3 C5 C4 >> $out 2>>gpurun_out/reduce_sweep.err
done
cat $out
