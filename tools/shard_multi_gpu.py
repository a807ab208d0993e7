"""One proof over N GPUs, one process per GPU (torchrun): correctness against the one-GPU proof and timings of the
sharded proof and of the sharded transform.  The exchange is inside libzkb200 (peer windows over NVLink); torch.distributed
only carries the window handles and the comparison of results.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29655 \
      tools/shard_multi_gpu.py [--log-n 20] [--steps 10] [--ntt 20 22 24]
"""
import argparse
import importlib
import json
import os
import random
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")  # before CUDA starts: one hardware queue per stream
FR = 21888242871839275222246405745257275088548364400416034343698204186575808495617


def main():
    import torch
    import torch.distributed as dist
    ap = argparse.ArgumentParser()
    ap.add_argument("--log-n", type=int, default=20)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--ntt", type=int, nargs="*", default=[])
    ap.add_argument("--trace", default=None, help="rank 0: per-kernel timeline of one sharded proof (zkb_trace_dump) to this path")
    ap.add_argument("--trace-batch", default=None, help="rank 0: timeline of a batch of 12 sharded proofs in flight")
    ap.add_argument("--skip-single", action="store_true", help="do not build the full CRS on rank 0 for the one-GPU comparison")
    args = ap.parse_args()
    zk = importlib.import_module("zksnark-rs_b200")
    zg = importlib.import_module("zksnark-rs_b200.groth16")
    zd = importlib.import_module("zksnark-rs_b200.dist")
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)  # NCCL's banner goes to stderr
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = torch.device("cuda", local)
    ctx = zk.Context(local)
    max_log = max([args.log_n] + args.ntt)
    comm = zd.connect(ctx, max_log, device=dev)
    out = {"world": world, "log_n": args.log_n}
    n = 1 << args.log_n
    rng = random.Random(3)
    toxic = tuple(rng.randrange(1, FR) for _ in range(5))
    r, s = rng.randrange(1, FR), rng.randrange(1, FR)
    qap = zk.QAP.horner(ctx, n)
    t0 = time.perf_counter()
    crs = zk.setup_shard(ctx, comm, qap, toxic)
    out["setup_shard_s"] = time.perf_counter() - t0
    wrng = random.Random(2)
    w = zg.fr_limbs(zg.horner_witness(n, wrng.randrange(1, FR), [wrng.getrandbits(253) for _ in range(n)]))
    d_w = ctx.dev_alloc(w.nbytes)
    ctx.h2d(d_w, w)
    dist.barrier()
    proof = zk.prove_shard(ctx, comm, qap, crs, d_w, r, s, on_device=True)
    flat = np.concatenate([zg.g1_pack([proof.a]).reshape(-1), zg.g2_pack([proof.b]).reshape(-1), zg.g1_pack([proof.c]).reshape(-1)])
    mine = torch.from_numpy(flat.view(np.int64).copy()).to(dev)
    allp = torch.empty(world * mine.numel(), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(allp, mine)
    allp = allp.cpu().numpy().reshape(world, -1)
    out["all_ranks_equal"] = bool((allp == allp[0]).all())
    if rank == 0 and not args.skip_single:
        crs1 = zk.setup(ctx, qap, toxic)
        single = zg.prove_dev(ctx, qap, crs1, d_w, r, s)
        out["equals_single_gpu"] = (single.a, single.b, single.c) == (proof.a, proof.b, proof.c)
        t0 = time.perf_counter()
        for _ in range(3):
            zg.prove_dev(ctx, qap, crs1, d_w, r, s)
        out["single_gpu_ms"] = (time.perf_counter() - t0) / 3 * 1e3
        crs1.free()
    dist.barrier()
    # latency: one proof at a time
    for _ in range(2):
        zk.prove_shard(ctx, comm, qap, crs, d_w, r, s, on_device=True)
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        zk.prove_shard(ctx, comm, qap, crs, d_w, r, s, on_device=True)
    out["latency_ms"] = (time.perf_counter() - t0) / args.steps * 1e3
    # throughput: several sharded proofs in flight
    zk.prove_shard_batch(ctx, comm, qap, crs, [d_w] * 4, [r] * 4, [s] * 4, on_device=True)
    dist.barrier()
    t0 = time.perf_counter()
    ps = zk.prove_shard_batch(ctx, comm, qap, crs, [d_w] * args.steps, [r] * args.steps, [s] * args.steps, on_device=True)
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out["batch_ms_per_proof"] = float(t.item()) / args.steps * 1e3
    out["batch_equal"] = all((p.a, p.b, p.c) == (proof.a, proof.b, proof.c) for p in ps)
    out["comm_status"] = comm.status()
    if args.trace:
        dist.barrier()
        if rank == 0:
            ctx.profile(2)
        zk.prove_shard(ctx, comm, qap, crs, d_w, r, s, on_device=True)
        if rank == 0:
            ctx.trace_dump(args.trace)
            ctx.profile(0)
    if args.trace_batch:
        dist.barrier()
        if rank == 0:
            ctx.profile(2)
        zk.prove_shard_batch(ctx, comm, qap, crs, [d_w] * 12, [r] * 12, [s] * 12, on_device=True)
        if rank == 0:
            ctx.trace_dump(args.trace_batch)
            ctx.profile(0)
    # the transform alone
    ntt = []
    for lg in args.ntt:
        nn = 1 << lg
        m = nn // world
        g = np.random.default_rng(lg)
        x = g.integers(0, 1 << 62, size=(m, 4), dtype=np.uint64)
        x[:, 3] &= np.uint64((1 << 60) - 1)
        d = ctx.dev_alloc(x.nbytes)
        ctx.h2d(d, x)
        for _ in range(3):
            zk.ntt_shard(ctx, comm, d, lg)
        dist.barrier()
        reps = 10
        t0 = time.perf_counter()
        for _ in range(reps):
            zk.ntt_shard(ctx, comm, d, lg, wait=False)
        ctx.sync()
        dt = (time.perf_counter() - t0) / reps
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item()) * 1e3
        ntt.append({"log_n": lg, "ms": ms, "GBps_algorithmic": 64.0 * nn / (ms * 1e-3) / 1e9,
                    "bytes_on_wire_total": nn * 32 * (world - 1) // world,
                    "note": "includes canonical<->Montgomery conversion of the n/G local elements on both sides"})
        ctx.dev_free(d)
    if ntt:
        out["ntt_shard"] = ntt
    if rank == 0:
        os.dup2(saved, 1)
        print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
