"""BASELINE.json config 4: stand-alone G1 MSM sweep, sharded over N GPUs (one process per GPU, torchrun).

  python tools/msm_multi_gpu.py [--sizes 18 20 22 24] [--reps 5]                    # N = 1
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
      tools/msm_multi_gpu.py --sizes 20 22 24

Two shardings of the same MSM:
  window : every rank holds all points and all scalars and takes the table rows (windows) j = rank (mod N):
           1/N of the point additions, full sort input; zkb_msm_windows.
  point  : rank g holds points and scalars [gn/N, (g+1)n/N): full Pippenger over its slice; zkb_msm.
Either way each rank ends with ONE partial point (64 B affine); the only exchange is an NCCL all-gather of those
N x 64 B, folded on the device by every rank (zkb_points_sum) -- EC addition is not an NCCL reduction operator.
The fold of both shardings is checked against each other and, at N = 1, is the plain zkb_msm result.
Timing: CUDA events on the library stream around `reps` back-to-back MSM calls, max over ranks, + the measured
all-gather + fold time.  Prints one JSON line per size on rank 0.
"""
import argparse
import importlib
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def rand_fr(rng, n):
    a = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64((1 << 60) - 1)
    return a


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", type=int, nargs="*", default=[18, 20, 22, 24])
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    zk = importlib.import_module("zksnark-rs_b200")
    zg = importlib.import_module("zksnark-rs_b200.groth16")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)  # NCCL's banner goes to stderr
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = zk.Context(local)
    rate, _ = ctx.bench_modmul(1, 4000)
    stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local))

    def gather_fold(part):
        """all-gather the ranks' partial points (64 B each) and fold them on the device."""
        if world == 1:
            return part
        t = torch.from_numpy(zg.g1_pack([part]).view(np.int64).reshape(-1)).cuda()
        out = torch.empty(world * t.numel(), dtype=torch.int64, device="cuda")
        dist.all_gather_into_tensor(out, t)
        pts = zg.g1_unpack(out.cpu().numpy().view(np.uint64).reshape(world, 8))
        return zg.points_sum(ctx, 1, pts)

    def timed(fn):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); ctx.sync()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0.record(stream)
        for _ in range(args.reps):
            r = fn()
        e1.record(stream)
        ctx.sync(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.reps
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, r

    lines = []
    for lg in args.sizes:
        n = 1 << lg
        rng = np.random.default_rng(1000 + lg)  # same stream on every rank
        k = rand_fr(rng, n)
        s = rand_fr(rng, n)
        lo, hi = n * rank // world, n * (rank + 1) // world
        # window sharding: all points, all scalars
        b_all = zk.Bases.generate(ctx, 1, k)
        d_s = ctx.dev_alloc(s.nbytes)
        ctx.h2d(d_s, s)
        zk.msm(ctx, b_all, d_s, on_device=True, n=n, windows=(rank, world))
        ms_win, part_w = timed(lambda: zk.msm(ctx, b_all, d_s, on_device=True, n=n, windows=(rank, world)))
        t0 = time.perf_counter()
        res_w = gather_fold(part_w)
        fold_ms = (time.perf_counter() - t0) * 1e3
        b_all.free()
        # point sharding: this rank's slice only
        b_loc = zk.Bases.generate(ctx, 1, k[lo:hi])
        ctx.h2d(d_s, np.ascontiguousarray(s[lo:hi]))
        zk.msm(ctx, b_loc, d_s, on_device=True, n=hi - lo)
        ms_pt, part_p = timed(lambda: zk.msm(ctx, b_loc, d_s, on_device=True, n=hi - lo))
        res_p = gather_fold(part_p)
        b_loc.free()
        ctx.dev_free(d_s)
        assert res_w == res_p, "window-sharded and point-sharded MSM disagree"
        if rank == 0:
            c = 0
            line = {"what": "G1 MSM stand-alone (config 4)", "log_n": lg, "n_gpus": world,
                    "window_sharded_ms": ms_win, "point_sharded_ms": ms_pt, "allgather_fold_ms_host_clock": fold_ms,
                    "mpoints_per_s_window": n / ms_win / 1e3, "mpoints_per_s_point": n / ms_pt / 1e3,
                    "modmul_peak_g": rate / 1e9, "result_x_low64": int(res_w[0] & 0xFFFFFFFFFFFFFFFF) if res_w else 0}
            lines.append(line)
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
