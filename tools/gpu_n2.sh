#!/bin/bash
# Two-GPU visit: one proof over 2 ranks (correctness vs one GPU, latency, throughput, timeline), sharded transforms, bench line.
tag=${1:-r02_n2}
N=${2:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/${tag}_topo.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655 \
  tools/shard_multi_gpu.py --log-n 20 --steps 20 --ntt 20 22 24 --trace gpurun_out/${tag}_trace_shard_2pow20.csv > gpurun_out/${tag}_shard.json 2> gpurun_out/${tag}_shard.err
echo "shard tool exit $?"; tail -c 1500 gpurun_out/${tag}_shard.json; tail -5 gpurun_out/${tag}_shard.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29656 \
  bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench exit $?"; tail -c 3000 gpurun_out/${tag}_bench.json; tail -5 gpurun_out/${tag}_bench.err
