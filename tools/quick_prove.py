"""Developer probe: device-resident prove() timing at 2^log_n (mean of reps).  Usage: python tools/quick_prove.py [log_n] [reps]"""
import importlib
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
zk = importlib.import_module("zksnark-rs_b200")
zg = importlib.import_module("zksnark-rs_b200.groth16")


def main():
    lg = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    n = 1 << lg
    ctx = zk.Context(0)
    q = zk.QAP.horner(ctx, n)
    t0 = time.perf_counter()
    crs = zk.setup(ctx, q, (3, 5, 7, 11, 13))
    ts = time.perf_counter() - t0
    rng = np.random.default_rng(1)
    w = rng.integers(0, 1 << 63, size=(2 * n + 2, 4), dtype=np.uint64)
    w[:, 3] &= np.uint64((1 << 60) - 1)
    d_w = ctx.dev_alloc(w.nbytes)
    ctx.h2d(d_w, w)
    for _ in range(3):
        p0 = zg.prove_dev(ctx, q, crs, d_w, 17, 19)
    t0 = time.perf_counter()
    for _ in range(reps):
        p = zg.prove_dev(ctx, q, crs, d_w, 17, 19)
    t = (time.perf_counter() - t0) / reps
    assert (p.a, p.b, p.c) == (p0.a, p0.b, p0.c)
    zk.prove_batch(ctx, q, crs, [d_w] * 4, [17] * 4, [19] * 4, on_device=True)  # warm-up: lane 1 scratch
    t0 = time.perf_counter()
    pb = zk.prove_batch(ctx, q, crs, [d_w] * reps, [17] * reps, [19] * reps, on_device=True)
    tb = (time.perf_counter() - t0) / reps
    assert all((x.a, x.b, x.c) == (p0.a, p0.b, p0.c) for x in pb)
    print(f"2^{lg} prove_batch x{reps}: {tb * 1e3:.3f} ms/proof ({1 / tb:.2f} proofs/s)")
    print(f"2^{lg} prove_dev: {t * 1e3:.3f} ms ({1 / t:.2f} proofs/s); setup {ts:.2f} s; env "
          + " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("ZKB_")), flush=True)


if __name__ == "__main__":
    main()
