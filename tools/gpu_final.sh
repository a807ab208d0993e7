#!/bin/bash
# Final visit of a round, most important first (the call may be cut by the GPU budget): parity suite, both bench arms,
# ncu launch list, full captures (G1 accumulation, TMA-staged NTT pass), sanitizers, remaining full captures.
tag=${1:-r01}
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
tail -3 gpurun_out/${tag}_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench exit $?"
timeout 120 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2>> gpurun_out/${tag}_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --skip-cpu --steps 2 --warmup 1 > gpurun_out/${tag}_ncu_bench.log 2>&1; echo "ncu list exit $?"
full() {  # name regex extra-bench-args
  timeout 240 ncu --set full --clock-control none --import-source on -k regex:$2 -s 8 -c 1 -f -o gpurun_out/${tag}_$1 python bench.py --skip-cpu $3 --steps 2 --warmup 1 > gpurun_out/${tag}_full_$1.log 2>&1; echo "ncu full $1 exit $?"
}
full acc_g1 "k_accumulate_chunks.*FqParams" ""
full ntt_tma "k_ntt_pass" "--log-n 16"
: > gpurun_out/${tag}_sanitizer.txt
for tool in racecheck memcheck; do
  echo "== compute-sanitizer --tool $tool python tests/sanitize_case.py" >> gpurun_out/${tag}_sanitizer.txt
  timeout 240 compute-sanitizer --tool $tool python tests/sanitize_case.py 2>&1 | grep -v "^=========     \|^$" | tail -6 >> gpurun_out/${tag}_sanitizer.txt
done
tail -8 gpurun_out/${tag}_sanitizer.txt
full acc_g2 "k_accumulate_chunks.*Fq2" ""
full ntt "k_ntt_pass" ""
ls -la gpurun_out/ | tail -24
