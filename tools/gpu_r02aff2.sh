#!/bin/bash
# Pair tree, second visit: K = 1 grids, short chunks behind the tree, the 168-register G2 build; per-level trace; ncu of the level-0 kernels
tag=${1:-r02aff2}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_affine_tree.py -q -x --timeout 120 > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
tail -3 gpurun_out/${tag}_pytest.log
timeout 300 python tools/aff_bench.py 20 > gpurun_out/${tag}_bench.jsonl 2> gpurun_out/${tag}_bench.err; echo "aff_bench exit $?"
cat gpurun_out/${tag}_bench.jsonl; tail -3 gpurun_out/${tag}_bench.err
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_affine_level -c 2 -f -o gpurun_out/${tag}_g2 python tools/aff_probe.py 2 5 > gpurun_out/${tag}_ncu_g2.log 2>&1; echo "ncu g2 exit $?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_affine_level -c 1 -f -o gpurun_out/${tag}_g1 python tools/aff_probe.py 1 4 > gpurun_out/${tag}_ncu_g1.log 2>&1; echo "ncu g1 exit $?"
