#!/bin/bash
# Short GPU visit: parity tests + quick timings + contract bench.  Usage: bash tools/gpu_quick.sh tag
tag=${1:-q}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
timeout 300 python tools/quick_bench.py 16 20 > gpurun_out/${tag}_quick.log 2>&1
cat gpurun_out/${tag}_quick.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench exit $?"
cat gpurun_out/${tag}_bench.json
