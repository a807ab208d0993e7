#!/bin/bash
# Small-size probe (BASELINE config 2, 2^16): proofs in flight (ZKB_LANES) and the kernel timeline of one proof.
tag=${1:-small}
mkdir -p gpurun_out
L=gpurun_out/${tag}.log
: > $L
for lanes in 1 2 3 4; do
  for lg in 16 18; do
    echo "== quick_prove $lg ZKB_LANES=$lanes" >> $L
    ZKB_LANES=$lanes timeout 90 python tools/quick_prove.py $lg 40 >> $L 2>&1
  done
done
echo "== trace 2^16" >> $L
timeout 90 python tools/trace_prove.py 16 gpurun_out/${tag}_trace16.csv >> $L 2>&1
cat $L
