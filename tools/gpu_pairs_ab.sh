#!/bin/bash
# A/B of the pair-sum accumulation (ZKB_ACC_PAIRS bit 0 = G1, bit 1 = G2): parity tests with it on, then timings.
tag=${1:-pairs}
mkdir -p gpurun_out
ZKB_ACC_PAIRS=3 timeout 700 python -m pytest tests -m gpu -x -q -k "msm or prove or sharded or single_gate or parser or cpp" > gpurun_out/${tag}_pytest.log 2>&1; tail -3 gpurun_out/${tag}_pytest.log
for v in 0 1 2 3; do
  echo "== ZKB_ACC_PAIRS=$v" | tee -a gpurun_out/${tag}_ab.log
  ZKB_ACC_PAIRS=$v timeout 200 python tools/msm_bench.py 20 2>&1 | tee -a gpurun_out/${tag}_ab.log
  ZKB_ACC_PAIRS=$v timeout 200 python tools/msm_bench.py --g1 22 2>&1 | tee -a gpurun_out/${tag}_ab.log
  ZKB_ACC_PAIRS=$v timeout 300 python bench.py --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('bench', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'g1 acc ms', round(d['roofline']['avg_launch_ms'],3), 'g2 acc ms', round(d['msm_g2']['total_ms']/d['msm_g2']['launches'],3))" | tee -a gpurun_out/${tag}_ab.log
done
