#!/bin/bash
# chunk-size sweep of the bucket accumulations on the workload of one rank of an 8-GPU proof at 2^20 (and of a 2-GPU one)
tag=${1:-r02e}
mkdir -p gpurun_out
L=gpurun_out/${tag}.log; : > $L
for w in 8 2; do
for s1 in 16 24 32 48 64 96 128; do
  for fr in 0.75 0.875; do
    echo "== world $w ZKB_ACC_S1=$s1 ZKB_ACC_S1_G2=$s1 ZKB_ACC_FRAC=$fr" >> $L
    ZKB_ACC_S1=$s1 ZKB_ACC_S1_G2=$s1 ZKB_ACC_FRAC=$fr timeout 120 python tools/probe_shard_rank.py 20 $w 2>&1 | tail -1 >> $L
  done
done
done
cat $L
