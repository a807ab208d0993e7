#!/bin/bash
# Final one-GPU visit of round 2 (after the pair-tree work): parity suite, smoke, both bench arms, ncu launch list.
tag=${1:-r02z}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
tail -3 gpurun_out/${tag}_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/${tag}_smoke.log
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench exit $?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2>> gpurun_out/${tag}_bench.err; echo "reference arm exit $?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --skip-cpu --steps 2 --warmup 1 --sustain 0 > gpurun_out/${tag}_ncu_bench.log 2>&1; echo "ncu list exit $?"
python - <<PY
import json
d = json.loads(open("gpurun_out/${tag}_bench.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "ms_per_step", "gpu_launches", "clocks")}, d["e2e"], d["roofline"]["frac"], d["config"].get("single_proof_latency_ms"), (d.get("sustained") or {}).get("value"))
PY
# occupancy pad for the G1 accumulation (4 instead of 5 blocks per SM): does the pipeline of a batch gain what the kernel loses?
for pad in 0 47000; do ZKB_ACC_PAD_G1=$pad timeout 200 python tools/quick_prove.py 20 12 2>&1 | tail -2; done > gpurun_out/${tag}_pad.log; cat gpurun_out/${tag}_pad.log
