#!/bin/bash
# witness generation v2 (cluster kernel + programmatic dependent launch) vs v1 (ZKB_WIT_CLUSTER=0: block runs up to 512 gates)
tag=${1:-r02j}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_witness.py -x -q -s > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
tail -6 gpurun_out/${tag}_pytest.log
timeout 600 python tools/witness_bench.py > gpurun_out/${tag}_witness.log 2>&1
cat gpurun_out/${tag}_witness.log
timeout 300 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_witness.py -x -q -k "layered and (64-64 or 33-20)" > gpurun_out/${tag}_racecheck.log 2>&1; tail -4 gpurun_out/${tag}_racecheck.log
