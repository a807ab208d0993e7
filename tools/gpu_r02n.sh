#!/bin/bash
# ncu --set full of the round's new kernels: quad bucket level (G2), Horner, witness level (PDL) and witness chain
tag=${1:-r02n}
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_level_quad|k_horner_final" -c 6 -f -o gpurun_out/${tag}_quad python tools/quick_prove.py 16 1 > gpurun_out/${tag}_quad.log 2>&1; echo "ncu quad exit $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_wit_level|k_wit_run" -s 4 -c 4 -f -o gpurun_out/${tag}_wit python tools/witness_bench.py > gpurun_out/${tag}_wit.log 2>&1; echo "ncu wit exit $?"
timeout 200 python bench.py --log-n 16 --steps 40 --warmup 5 --skip-cpu > gpurun_out/${tag}_bench_2pow16.json 2> gpurun_out/${tag}_bench_2pow16.err; echo "bench 2^16 exit $?"
tail -c 600 gpurun_out/${tag}_bench_2pow16.json
ls -la gpurun_out/${tag}_*.ncu-rep
