#!/bin/bash
# witness generation: parity tests + timing of the layered 2^20 circuit and the Horner chain
tag=${1:-r02f}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_witness.py -x -q -s > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
tail -15 gpurun_out/${tag}_pytest.log
timeout 600 python tools/witness_bench.py > gpurun_out/${tag}_witness.log 2>&1
cat gpurun_out/${tag}_witness.log
