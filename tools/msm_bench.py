"""Developer probe: standalone MSM timing (whole call and the accumulate kernel alone) at given sizes.
Usage: python tools/msm_bench.py [--g1] [--g2] [log_n ...]"""
import importlib
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
zk = importlib.import_module("zksnark-rs_b200")


def rand_fr(rng, n):
    a = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64((1 << 60) - 1)
    return a


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    groups = [g for g, f in ((1, "--g1"), (2, "--g2")) if f in sys.argv] or [1, 2]
    sizes = [int(a) for a in args] or [20]
    ctx = zk.Context(0)
    rate, _ = ctx.bench_modmul(1, 4000)
    rng = np.random.default_rng(1)
    for lg in sizes:
        n = 1 << lg
        for group in groups:
            b = zk.Bases.generate(ctx, group, rand_fr(rng, n))
            s = rand_fr(rng, n)
            ds = ctx.dev_alloc(s.nbytes)
            ctx.h2d(ds, s)
            for _ in range(2):
                zk.msm(ctx, b, ds, on_device=True, n=n)
            ctx.profile(True)
            t0 = time.perf_counter()
            reps = 5
            for _ in range(reps):
                zk.msm(ctx, b, ds, on_device=True, n=n)
            wall = (time.perf_counter() - t0) / reps
            ms, cnt, units = ctx.profile_read(2 if group == 1 else 3)
            ctx.profile(False)
            per = ms / cnt
            mm = (10 if group == 1 else 28) * units / cnt / (per * 1e-3)
            print(f"2^{lg} G{group}: call {wall * 1e3:.3f} ms; accumulate {per:.3f} ms, {units / cnt / 1e6:.1f} M records, "
                  f"{mm / 1e9:.1f} Gmodmul/s = {mm / rate:.3f} of measured peak {rate / 1e9:.1f}", flush=True)
            ctx.dev_free(ds)
            b.free()


if __name__ == "__main__":
    main()
