#!/bin/bash
# quad-lane bucket hierarchy: parity (MSM + prove tests), then A/B against the previous plan (ZKB_TAIL=1)
tag=${1:-r02g}
mkdir -p gpurun_out
L=gpurun_out/${tag}.log; : > $L
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -x -q -k "msm or prove or golden or setup" > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
for mode in 1 0; do
  for lg in 16 18 20; do
    echo "== ZKB_TAIL=$mode 2^$lg" >> $L
    ZKB_TAIL=$mode timeout 300 python tools/quick_prove.py $lg 10 2>&1 | tail -2 >> $L
  done
  echo "== ZKB_TAIL=$mode shard ranks" >> $L
  ZKB_TAIL=$mode timeout 300 python tools/probe_shard_rank.py 20 1 8 2>&1 | tail -2 >> $L
  ZKB_TAIL=$mode timeout 120 python tools/trace_prove.py 16 gpurun_out/${tag}_trace16_tail$mode.csv > /dev/null 2>&1
  ZKB_TAIL=$mode timeout 120 python tools/trace_prove.py 20 gpurun_out/${tag}_trace20_tail$mode.csv > /dev/null 2>&1
done
for l0 in 65536 262144 2097152; do
  echo "== ZKB_TAIL_L0=$l0" >> $L
  ZKB_TAIL_L0=$l0 timeout 300 python tools/quick_prove.py 20 10 2>&1 | tail -2 >> $L
  ZKB_TAIL_L0=$l0 timeout 300 python tools/probe_shard_rank.py 20 8 2>&1 | tail -1 >> $L
done
cat $L
