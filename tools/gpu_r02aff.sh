#!/bin/bash
# Batched-affine pair tree (ZKB_AFF_*): parity on the device, memcheck on the sanitizer workload, then the sweep against the chain.
tag=${1:-r02aff}
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_affine_tree.py -q --timeout 120 > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
timeout 300 python tools/aff_bench.py 20 > gpurun_out/${tag}_bench.jsonl 2> gpurun_out/${tag}_bench.err; echo "aff_bench exit $?"
cat gpurun_out/${tag}_bench.jsonl; tail -3 gpurun_out/${tag}_bench.err
ZKB_AFF_G1=3 ZKB_AFF_G2=3 timeout 240 compute-sanitizer --tool memcheck python tests/sanitize_case.py 2>&1 | grep -v "^=========     \|^$" | tail -8 > gpurun_out/${tag}_memcheck.txt
cat gpurun_out/${tag}_memcheck.txt
