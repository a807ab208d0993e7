#!/bin/bash
# full GPU suite on the round-2 code, NTT warp-shuffle A/B, ncu of the G2 accumulation (registers vs shared-memory accumulator)
tag=${1:-r02i}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
L=gpurun_out/${tag}_ntt.log; : > $L
ZKB_NTT_SHFL=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_large.py -x -q -k "ntt" >> $L 2>&1
for round in 1 2; do
  for v in "0 0" "0 1" "3 0"; do
    set -- $v
    echo "== round $round ZKB_NTT_TMA=$1 ZKB_NTT_SHFL=$2" >> $L
    ZKB_NTT_TMA=$1 ZKB_NTT_SHFL=$2 timeout 200 python tools/ntt_ab.py >> $L 2>&1
  done
done
cat $L
for sm in 0 1; do
  ZKB_ACC_SM=$sm timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_accumulate_chunks -s 4 -c 2 -o gpurun_out/${tag}_acc_sm$sm python bench.py --steps 2 --warmup 1 --skip-cpu > gpurun_out/${tag}_ncu_sm$sm.log 2>&1; echo "ncu sm$sm exit $?"
done
ls -la gpurun_out/${tag}_acc_sm*.ncu-rep
