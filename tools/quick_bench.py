"""Developer timing probe (not the contract bench): modmul peak, NTT, MSM, prove at a few sizes.
Usage: python tools/quick_bench.py [log_n ...]"""
import ctypes as C
import importlib
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
zk = importlib.import_module("zksnark-rs_b200")
zg = importlib.import_module("zksnark-rs_b200.groth16")


def rand_fr(rng, n):
    a = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64((1 << 60) - 1)
    return a


def timeit(f, reps=3):
    f()
    best = 1e9
    for _ in range(reps):
        t = time.perf_counter()
        f()
        best = min(best, time.perf_counter() - t)
    return best


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [16, 20]
    ctx = zk.Context(0)
    for field in (0, 1):
        rate, ms = ctx.bench_modmul(field, 4000)
        print(f"modmul peak field={field}: {rate / 1e9:.1f} Gmodmul/s ({ms:.2f} ms)", flush=True)
    rng = np.random.default_rng(1)
    for lg in sizes:
        n = 1 << lg
        a = rand_fr(rng, n)
        d = ctx.dev_alloc(a.nbytes)
        ctx.h2d(d, a)
        t = timeit(lambda: ctx.check(ctx.lib.zkb_ntt_fr_raw(ctx.h, C.c_void_p(d), lg, 0), "ntt"))
        print(f"2^{lg} NTT raw: {t * 1e3:.3f} ms  {64 * n / t / 1e9:.1f} GB/s", flush=True)
        ctx.dev_free(d)
        for group in (1, 2):
            t0 = time.perf_counter()
            b = zk.Bases.generate(ctx, group, rand_fr(rng, n))
            tg = time.perf_counter() - t0
            s = rand_fr(rng, n)
            ds = ctx.dev_alloc(s.nbytes)
            ctx.h2d(ds, s)
            t = timeit(lambda: zk.msm(ctx, b, ds, on_device=True, n=n))
            print(f"2^{lg} MSM G{group}: {t * 1e3:.3f} ms  {n / t / 1e6:.2f} Mpoints/s  (bases gen {tg:.2f} s)", flush=True)
            ctx.dev_free(ds)
            b.free()
        t0 = time.perf_counter()
        q = zk.QAP.horner(ctx, n)
        crs = zk.setup(ctx, q, (3, 5, 7, 11, 13))
        print(f"2^{lg} qap+setup: {time.perf_counter() - t0:.2f} s", flush=True)
        w = rand_fr(rng, 2 * n + 2)
        t = timeit(lambda: zk.prove(ctx, q, crs, w, 17, 19))
        print(f"2^{lg} prove (host weights): {t * 1e3:.3f} ms  {1 / t:.2f} proofs/s  launches={ctx.launches}", flush=True)
        crs.free()
        q.free()


if __name__ == "__main__":
    main()
