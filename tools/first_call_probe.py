"""Developer probe: wall time of the first single proofs after a batch warm-up (first use of the quad-hierarchy kernels).
Usage: [CUDA_MODULE_LOADING=EAGER] python tools/first_call_probe.py [log_n]"""
import importlib
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
zk = importlib.import_module("zksnark-rs_b200")
zg = importlib.import_module("zksnark-rs_b200.groth16")

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n = 1 << lg
ctx = zk.Context(0)
q = zk.QAP.horner(ctx, n)
crs = zk.setup(ctx, q, (3, 5, 7, 11, 13))
rng = np.random.default_rng(1)
w = rng.integers(0, 1 << 63, size=(2 * n + 2, 4), dtype=np.uint64)
w[:, 3] &= np.uint64((1 << 60) - 1)
d_w = ctx.dev_alloc(w.nbytes)
ctx.h2d(d_w, w)
zk.prove_batch(ctx, q, crs, [d_w] * 6, [17] * 6, [19] * 6, on_device=True)
ts = []
for _ in range(5):
    t0 = time.perf_counter()
    zg.prove_dev(ctx, q, crs, d_w, 17, 19)
    ts.append(round((time.perf_counter() - t0) * 1e3, 2))
print("prove_dev calls after a batch warm-up, ms:", ts, "CUDA_MODULE_LOADING=" + os.environ.get("CUDA_MODULE_LOADING", "(default)"), flush=True)
