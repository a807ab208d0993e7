#!/bin/bash
# Pair tree, third visit: the staged level body (cp.async + shared memory, ZKB_AFF_VAR=2): parity, memcheck, sweep, ncu of the G2 level-0 kernel
tag=${1:-r02aff3}
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_affine_tree.py -q -x --timeout 120 > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
tail -3 gpurun_out/${tag}_pytest.log
timeout 300 python tools/aff_bench.py 20 > gpurun_out/${tag}_bench.jsonl 2> gpurun_out/${tag}_bench.err; echo "aff_bench exit $?"
cat gpurun_out/${tag}_bench.jsonl; tail -3 gpurun_out/${tag}_bench.err
ZKB_AFF_VAR=2 timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_affine_level -c 1 -f -o gpurun_out/${tag}_g2 python tools/aff_probe.py 2 5 > gpurun_out/${tag}_ncu_g2.log 2>&1; echo "ncu g2 exit $?"
ZKB_AFF_VAR=2 ZKB_AFF_G1=3 ZKB_AFF_G2=3 timeout 240 compute-sanitizer --tool memcheck python tests/sanitize_case.py 2>&1 | grep -v "^=========     \|^$" | tail -8 > gpurun_out/${tag}_memcheck.txt
cat gpurun_out/${tag}_memcheck.txt
