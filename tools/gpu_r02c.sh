#!/bin/bash
# Single-GPU visit: full GPU suite, the contract bench line (new: int_pipe roofline, sustained pass, measured CPU ratio),
# reference arm, DRAM bytes of the G1 accumulation at L2 fetch granularity 128 / 64 / 32.
tag=${1:-r02c}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
tail -3 gpurun_out/${tag}_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench exit $?"
tail -3 gpurun_out/${tag}_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2>> gpurun_out/${tag}_bench.err; echo "ref exit $?"
for g in 0 64 32; do
  env $( [ $g != 0 ] && echo ZKB_L2_FETCH=$g ) timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
      --clock-control none -k regex:k_accumulate_chunks -s 4 -c 2 --csv --log-file gpurun_out/${tag}_l2fetch_${g}.csv python tools/quick_prove.py 20 2 > /dev/null 2>&1
  echo "ncu l2fetch $g exit $?"
done
ls -la gpurun_out | tail -12
