#!/bin/bash
# G2 mixed addition with the two-reduction a b - c d (fq2_msm_lazy): parity, then timing against the previous build
tag=${1:-r02l}
mkdir -p gpurun_out
L=gpurun_out/${tag}.log; : > $L
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -x -q -k "fq2 or msm or prove or golden" > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
tail -4 gpurun_out/${tag}_pytest.log
for rep in 1 2; do
  timeout 300 python tools/quick_prove.py 20 10 2>&1 | tail -2 >> $L
  PROBE_BATCH=12 timeout 300 python tools/probe_shard_rank.py 20 1 2>&1 | tail -2 >> $L
done
timeout 300 python tools/quick_prove.py 16 10 2>&1 | tail -2 >> $L
cat $L
