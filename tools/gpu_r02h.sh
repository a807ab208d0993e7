#!/bin/bash
# (1) G2 accumulation with the accumulator in shared memory (168 registers, 3 blocks/SM) vs registers (230, 2 blocks/SM)
# (2) bucket hierarchy: quad plan with thread levels above QMAX, automatic policy (batch of big proofs -> v1)
tag=${1:-r02h}
mkdir -p gpurun_out
L=gpurun_out/${tag}.log; : > $L
ZKB_ACC_SM=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -x -q -k "msm or prove or golden" > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
tail -4 gpurun_out/${tag}_pytest.log
for sm in 0 1; do
  echo "== ZKB_ACC_SM=$sm" >> $L
  ZKB_ACC_SM=$sm timeout 300 python tools/quick_prove.py 20 10 2>&1 | tail -2 >> $L
  ZKB_ACC_SM=$sm PROBE_BATCH=12 timeout 300 python tools/probe_shard_rank.py 20 1 8 2>&1 | tail -4 >> $L
  ZKB_ACC_SM=$sm timeout 300 python tools/quick_prove.py 16 10 2>&1 | tail -2 >> $L
done
for q in 2048 8192 32768; do
  for lg in 16 18; do
    echo "== ZKB_TAIL_QMAX=$q 2^$lg" >> $L
    ZKB_TAIL_QMAX=$q timeout 300 python tools/quick_prove.py $lg 10 2>&1 | tail -2 >> $L
  done
  echo "== ZKB_TAIL_QMAX=$q rank of 8" >> $L
  ZKB_TAIL_QMAX=$q PROBE_BATCH=12 timeout 300 python tools/probe_shard_rank.py 20 8 2>&1 | tail -2 >> $L
done
echo "== ZKB_TAIL=1 rank of 8 batch" >> $L
ZKB_TAIL=1 PROBE_BATCH=12 timeout 300 python tools/probe_shard_rank.py 20 8 2>&1 | tail -2 >> $L
timeout 120 python tools/trace_prove.py 16 gpurun_out/${tag}_trace16.csv > /dev/null 2>&1
ZKB_ACC_SM=1 timeout 120 python tools/trace_prove.py 20 gpurun_out/${tag}_trace20_sm1.csv > /dev/null 2>&1
cat $L
