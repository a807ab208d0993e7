#!/bin/bash
# One GPU visit: parity tests, the contract bench, an ncu launch list of one bench run.
# Usage (under gpurun): bash tools/gpu_round.sh [tag]
tag=${1:-r1}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/${tag}_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench exit $?"
cat gpurun_out/${tag}_bench.json
timeout 300 python tools/quick_bench.py 16 20 > gpurun_out/${tag}_quick.log 2>&1
cat gpurun_out/${tag}_quick.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 1 --warmup 1 > gpurun_out/${tag}_ncu_bench.log 2>&1; echo "ncu exit $?"
