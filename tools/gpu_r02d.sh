#!/bin/bash
tag=${1:-r02d}
mkdir -p gpurun_out
L=gpurun_out/${tag}.log; : > $L
echo "== new chunk plan (wave-aware)" >> $L
timeout 300 python tools/probe_shard_rank.py 20 1 2 8 >> $L 2>&1
timeout 200 python tools/quick_prove.py 20 20 >> $L 2>&1
timeout 200 python tools/quick_prove.py 16 40 >> $L 2>&1
echo "== old fixed chunks for comparison: ZKB_ACC_S=64" >> $L
ZKB_ACC_S=64 timeout 300 python tools/probe_shard_rank.py 20 8 >> $L 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "msm or prove or shard" >> $L 2>&1
for v in hint nohint; do
  lib=""; [ $v = nohint ] && lib="ZKB200_LIB=zksnark-rs_b200/_var/libzkb200_nohint.so"
  env $lib timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
      --clock-control none -k regex:k_accumulate_chunks -s 4 -c 2 --csv --log-file gpurun_out/${tag}_dram_${v}.csv python tools/quick_prove.py 20 2 > /dev/null 2>&1
  echo "ncu $v exit $?" >> $L
done
cat $L
