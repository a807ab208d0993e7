#!/bin/bash
# ncu launch list of one bench run (shares, not absolutes).  Usage: bash tools/gpu_ncu.sh tag [extra bench args]
tag=${1:-n}; shift
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 1 --warmup 1 "$@" > gpurun_out/${tag}_ncu_bench.log 2>&1; echo "ncu exit $?"
