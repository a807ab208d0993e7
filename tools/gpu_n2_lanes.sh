#!/bin/bash
# 2 GPUs at 2^18: per-rank work of an 8-GPU proof at 2^20 (n/G = 2^17).  Throughput vs proofs in flight; batch timeline.
tag=${1:-r02_n2l}
mkdir -p gpurun_out
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port "$1" "${@:2}"; }
p=29670
for l in 1 2 3 4; do
  p=$((p+1))
  ZKB_LANES=$l run $p tools/shard_multi_gpu.py --log-n 18 --steps 40 --skip-single > gpurun_out/${tag}_lanes$l.json 2> gpurun_out/${tag}_lanes$l.err
  echo "lanes $l: $(python -c "import json;d=json.loads(open('gpurun_out/${tag}_lanes$l.json').read().strip().splitlines()[-1]);print(d['latency_ms'], d['batch_ms_per_proof'], d['batch_equal'])")"
done
run 29680 tools/shard_multi_gpu.py --log-n 18 --steps 40 --trace-batch gpurun_out/${tag}_trace_batch.csv > gpurun_out/${tag}_tb.json 2> gpurun_out/${tag}_tb.err
timeout 200 python tools/quick_prove.py 18 40
