#!/bin/bash
# Last visit: parity suite + smoke on the final build, then full captures of the SM-filling kernels besides the accumulations
# (record sort, head folding) for the next round.
tag=${1:-last}
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?"
tail -2 gpurun_out/${tag}_pytest.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/${tag}_smoke.log
timeout 60 python tools/trace_prove.py 16 gpurun_out/${tag}_trace16.csv | grep ntt | head -3
timeout 200 ncu --set full --clock-control none --import-source on -k "regex:k_digits_scatter|k_digits_count|k_fix_heads" -s 10 -c 5 -f -o gpurun_out/${tag}_sort python bench.py --skip-cpu --steps 2 --warmup 1 > gpurun_out/${tag}_full_sort.log 2>&1; echo "ncu full sort exit $?"
ls -la gpurun_out | grep ${tag}
