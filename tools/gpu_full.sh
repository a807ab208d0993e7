#!/bin/bash
# ncu --set full capture of selected kernels during one bench step.  Usage: bash tools/gpu_full.sh tag regex [skip] [count]
tag=${1:-f}; rx=${2:-k_accumulate_chunks}; skip=${3:-2}; cnt=${4:-2}
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c $cnt -f -o gpurun_out/${tag}_full python bench.py --steps 1 --warmup 1 > gpurun_out/${tag}_full.log 2>&1; echo "ncu full exit $?"
ls -la gpurun_out/${tag}_full.ncu-rep
