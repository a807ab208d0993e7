"""Developer probe: kernel timeline (both streams) of one prove() at 2^log_n via zkb_profile(ctx, 2).
Usage: python tools/trace_prove.py [log_n] [out.csv]"""
import importlib
import os
import random
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
zk = importlib.import_module("zksnark-rs_b200")
zg = importlib.import_module("zksnark-rs_b200.groth16")


def main():
    lg = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    out = sys.argv[2] if len(sys.argv) > 2 else f"gpurun_out/trace_{lg}.csv"
    n = 1 << lg
    ctx = zk.Context(0)
    q = zk.QAP.horner(ctx, n)
    crs = zk.setup(ctx, q, (3, 5, 7, 11, 13))
    rng = np.random.default_rng(1)
    w = rng.integers(0, 1 << 63, size=(2 * n + 2, 4), dtype=np.uint64)
    w[:, 3] &= np.uint64((1 << 60) - 1)
    d_w = ctx.dev_alloc(w.nbytes)
    ctx.h2d(d_w, w)
    for _ in range(3):
        zg.prove_dev(ctx, q, crs, d_w, 17, 19)
    ctx.profile(2)
    zg.prove_dev(ctx, q, crs, d_w, 17, 19)
    ctx.trace_dump(out)
    ctx.profile(0)
    import csv
    rows = list(csv.reader(open(out)))[1:]
    for r in rows:
        print(f"s{r[0]} {float(r[2]):8.3f} -> {float(r[3]):8.3f}  {float(r[4]):7.3f} ms  {r[1][:60]}")


if __name__ == "__main__":
    main()
