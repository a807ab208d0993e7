#!/bin/bash
# Prepared at the end of round 1 (not yet run): the first GPU visit of the next round.
#  1. L2 fetch granularity for the 64-byte table gathers (ZKB_L2_FETCH): DRAM bytes of the
#     G1 accumulation per launch and ms per proof at 128 (default) / 64 / 32 bytes.
#  2. G2 window size around the cost model's choice (ZKB_MSM_C2 = 16 .. 19 at 2^20).
#  3. The 2^22 bench line (BASELINE config 5, one GPU) with the measured peaks and the CPU legs.
tag=${1:-r02a}
mkdir -p gpurun_out
L=gpurun_out/${tag}.log
: > $L
for g in 0 64 32; do
  echo "== ZKB_L2_FETCH=$g" >> $L
  env $( [ $g != 0 ] && echo ZKB_L2_FETCH=$g ) timeout 120 python tools/quick_prove.py 20 20 >> $L 2>&1
  env $( [ $g != 0 ] && echo ZKB_L2_FETCH=$g ) timeout 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
      --clock-control none -k regex:k_accumulate_chunks -s 4 -c 2 --csv python tools/quick_prove.py 20 2 2>/dev/null | grep -E "k_accumulate|dram__|gpu__time" | cut -c1-220 >> $L
done
for c2 in 16 18 19; do
  echo "== ZKB_MSM_C2=$c2" >> $L
  ZKB_MSM_C2=$c2 timeout 120 python tools/quick_prove.py 20 20 >> $L 2>&1
done
timeout 600 python bench.py --log-n 22 --steps 5 --warmup 3 > gpurun_out/${tag}_bench_2pow22.json 2> gpurun_out/${tag}_bench_2pow22.err; echo "bench 2^22 exit $?" >> $L
cat $L
