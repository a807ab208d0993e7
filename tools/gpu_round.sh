#!/bin/bash
# One full GPU visit: parity tests, the contract bench (both arms), an ncu launch list of one bench
# run, `ncu --set full` captures of the three kernels the roofline talks about, sanitizers.
# Usage (under gpurun): bash tools/gpu_round.sh [tag]
tag=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/${tag}_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
tail -3 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench exit $?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2>> gpurun_out/${tag}_bench.err
timeout 300 python tools/verify_bench.py 10 1 1024 16384 > gpurun_out/${tag}_verify.log 2>&1
timeout 300 python tools/msm_bench.py --g1 18 20 22 > gpurun_out/${tag}_msm_sweep.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --skip-cpu --steps 2 --warmup 1 > gpurun_out/${tag}_ncu_bench.log 2>&1; echo "ncu list exit $?"
for k in "acc_g1:k_accumulate_chunks.*FqParams" "acc_g2:k_accumulate_chunks.*Fq2" "ntt:k_ntt_pass"; do
  name=${k%%:*}; rx=${k#*:}
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$rx -s 8 -c 1 -f -o gpurun_out/${tag}_${name} python bench.py --skip-cpu --steps 2 --warmup 1 > gpurun_out/${tag}_full_${name}.log 2>&1; echo "ncu full $name exit $?"
done
# the TMA-staged twiddle path (default up to 2^18): one full capture at 2^16
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ntt_pass -s 8 -c 1 -f -o gpurun_out/${tag}_ntt_tma python bench.py --skip-cpu --log-n 16 --steps 2 --warmup 1 > gpurun_out/${tag}_full_ntt_tma.log 2>&1; echo "ncu full ntt_tma exit $?"
: > gpurun_out/${tag}_sanitizer.txt
for tool in memcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $tool python tests/sanitize_case.py" >> gpurun_out/${tag}_sanitizer.txt
  timeout 900 compute-sanitizer --tool $tool python tests/sanitize_case.py 2>&1 | grep -v "^=========     \|^$" | tail -6 >> gpurun_out/${tag}_sanitizer.txt
done
tail -12 gpurun_out/${tag}_sanitizer.txt
ls -la gpurun_out/ | tail -20
