"""One Fr transform with its outer dimension sharded over N GPUs (one process per GPU, torchrun).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29521 \
      tools/ntt_multi_gpu.py --sizes 20 22 24

Rank g transforms x[g::N] (size n/N, zkb_ntt_fr), the partial transforms are all-gathered over NCCL (n*32 bytes
arrive at every rank) and every rank evaluates its slice of the final coefficient reduction (zkb_ntt_combine).
Checked: the slices of all ranks, gathered, equal the single-GPU transform of the same vector (rank 0 computes it).
Timing: device-resident input, CUDA events on the torch stream around local transform + all-gather + combine,
max over ranks; the single-GPU zkb_ntt_fr on the whole vector is timed beside it.  One JSON line per size on rank 0.
"""
import argparse
import importlib
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", type=int, nargs="*", default=[20, 22, 24])
    ap.add_argument("--reps", type=int, default=10)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    zk = importlib.import_module("zksnark-rs_b200")
    zg = importlib.import_module("zksnark-rs_b200.groth16")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)
    ctx = zk.Context(local)
    log_g = world.bit_length() - 1
    assert 1 << log_g == world
    lines = []
    for lg in args.sizes:
        n = 1 << lg
        sub = n >> log_g
        rng = np.random.default_rng(2000 + lg)  # the same vector on every rank
        x = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64)
        x[:, 3] &= np.uint64((1 << 60) - 1)
        mine0 = torch.from_numpy(np.ascontiguousarray(x[rank::world]).view(np.int64)).to(dev)
        mine = torch.empty_like(mine0)
        parts = torch.empty((world, sub, 4), dtype=torch.int64, device=dev)
        out = torch.empty((sub, 4), dtype=torch.int64, device=dev)

        def step():
            mine.copy_(mine0)
            torch.cuda.current_stream().synchronize()
            zg.ntt_dev(ctx, mine.data_ptr(), lg - log_g, False)
            if world > 1:
                dist.all_gather_into_tensor(parts.view(-1), mine.view(-1))
                torch.cuda.current_stream().synchronize()
            else:
                parts[0].copy_(mine)
                torch.cuda.current_stream().synchronize()
            zg.ntt_combine(ctx, parts.data_ptr(), lg, log_g, False, rank * sub, sub, out.data_ptr())

        for _ in range(3):
            step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.reps
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        # check against the single-GPU transform (rank 0), and time it
        if world > 1:
            full = torch.empty((world, sub, 4), dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(full.view(-1), out.view(-1))
        else:
            full = out.view(1, sub, 4)
        single_ms = None
        if rank == 0:
            whole0 = torch.from_numpy(x.view(np.int64)).to(dev)
            whole = whole0.clone()
            zg.ntt_dev(ctx, whole.data_ptr(), lg, False)
            assert torch.equal(whole.view(-1), full.view(-1)), "sharded transform != single-GPU transform"
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record()
            for _ in range(args.reps):
                zg.ntt_dev(ctx, whole.data_ptr(), lg, False)
            f1.record()
            torch.cuda.synchronize()
            single_ms = f0.elapsed_time(f1) / args.reps
            lines.append({"what": "Fr NTT, outer dimension sharded", "log_n": lg, "n_gpus": world, "sharded_ms": ms,
                          "single_gpu_zkb_ntt_fr_ms": single_ms, "allgather_bytes_per_rank": n * 32,
                          "note": "both timings include the canonical <-> Montgomery conversions and bit-reversal of zkb_ntt_fr; "
                                  "sharded = copy + local transform + NCCL all-gather + combine, host-synchronised between stages"})
    if world > 1:
        sys.stdout.flush()
        os.dup2(saved, 1)
    if rank == 0:
        for l in lines:
            print(json.dumps(l), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
