#!/bin/bash
# Multi-GPU visit (N ranks, one process per GPU): one proof over N ranks -- correctness vs one GPU, latency, throughput,
# timeline, sharded transforms (tools/shard_multi_gpu.py) -- then the contract bench line at 2^20 and at 2^22 (BASELINE config 5).
# Usage (under gpurun --gpus N): bash tools/gpu_multi.sh <tag> <N> [sizes, default "20 22"]
tag=${1:-r02_n8}
N=${2:-8}
sizes=${3:-"20 22"}
mkdir -p gpurun_out
run() { timeout "$1" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port "$2" "${@:3}"; }
[ -n "$SKIP_TOOL" ] || run 600 29655 tools/shard_multi_gpu.py --log-n 20 --steps 20 --ntt 20 22 24 --trace gpurun_out/${tag}_trace_shard_2pow20.csv \
  > gpurun_out/${tag}_shard.json 2> gpurun_out/${tag}_shard.err
echo "shard tool exit $?"; tail -c 1800 gpurun_out/${tag}_shard.json; grep -v "^\*\|OMP_NUM\|^$\|NCCL version" gpurun_out/${tag}_shard.err | tail -5
port=29660
for lg in $sizes; do
  port=$((port + 1))
  run 1200 $port bench.py --gpus $N --log-n $lg --steps 20 --warmup 3 > gpurun_out/${tag}_bench_2pow${lg}.json 2> gpurun_out/${tag}_bench_2pow${lg}.err
  echo "bench 2^$lg exit $?"; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_bench_2pow${lg}.json").read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ("value", "ms_per_step", "scaling", "n_gpus", "other_mode", "clocks")}, d["e2e"], d["config"]["single_proof_latency_ms"], (d.get("sustained") or {}).get("value"))
except Exception as e:
    print("no line:", e)
PY
  grep -v "^\*\|OMP_NUM\|^$\|NCCL version" gpurun_out/${tag}_bench_2pow${lg}.err | tail -5
done
