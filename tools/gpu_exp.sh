#!/bin/bash
# Experiment visit: parity tests, then the contract bench under a few developer switches.  Usage: bash tools/gpu_exp.sh tag
tag=${1:-x}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
tail -4 gpurun_out/${tag}_pytest.log
for l in 2 3; do
  ZKB_LANES=$l timeout 300 python bench.py --steps 12 --warmup 3 > gpurun_out/${tag}_bench_lanes$l.json 2> gpurun_out/${tag}_bench_lanes$l.err; echo "lanes $l exit $?"
  python - <<PY
import json
d=json.load(open("gpurun_out/${tag}_bench_lanes$l.json"))
print("lanes $l:", round(d["value"],2), "proofs/s, e2e", round(d["e2e"]["value"],2), "acc frac", round(d["roofline"]["int_pipe"]["frac"],3), "cpu", d.get("cpu_baseline",{}).get("value"))
PY
done
timeout 300 python bench.py --log-n 16 --steps 20 --warmup 3 > gpurun_out/${tag}_bench_2pow16.json 2>> gpurun_out/${tag}_bench.err; echo "2^16 exit $?"
timeout 600 python bench.py --log-n 22 --steps 4 --warmup 3 > gpurun_out/${tag}_bench_2pow22.json 2>> gpurun_out/${tag}_bench.err; echo "2^22 exit $?"
python - <<PY
import json
for t in ("2pow16","2pow22"):
    try:
        d=json.load(open("gpurun_out/${tag}_bench_%s.json"%t)); print(t, round(d["value"],2), "proofs/s", round(d["ms_per_step"],2), "ms/proof; e2e", round(d["e2e"]["value"],2))
    except Exception as e: print(t, "failed", e)
PY
