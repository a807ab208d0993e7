#!/bin/bash
# Lanes-by-size default + G1 tail on its own stream: parity suite, then batch / single timings at three sizes.
tag=${1:-lanes}
mkdir -p gpurun_out
L=gpurun_out/${tag}.log
: > $L
timeout 420 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> $L
tail -3 gpurun_out/${tag}_pytest.log >> $L
for lg in 16 18 20; do
  echo "== quick_prove $lg (default lanes)" >> $L
  timeout 120 python tools/quick_prove.py $lg 40 >> $L 2>&1
done
echo "== quick_prove 20 ZKB_LANES=3" >> $L
ZKB_LANES=3 timeout 120 python tools/quick_prove.py 20 40 >> $L 2>&1
echo "== trace 2^16" >> $L
timeout 90 python tools/trace_prove.py 16 gpurun_out/${tag}_trace16.csv | tail -22 >> $L 2>&1
timeout 300 python bench.py --log-n 16 --steps 40 --warmup 5 > gpurun_out/${tag}_bench16.json 2> gpurun_out/${tag}_bench16.err; echo "bench16 exit $?" >> $L
cat $L
cat gpurun_out/${tag}_bench16.json
