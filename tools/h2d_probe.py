"""Developer probe: host -> device copy rate of one 2^20 witness (64 MiB) from pinned and from pageable memory."""
import importlib
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
zk = importlib.import_module("zksnark-rs_b200")

ctx = zk.Context(0)
m = 2 * (1 << 20) + 2
pin = ctx.pinned((m, 4))
pin[:] = 7
pag = np.full((m, 4), 7, dtype=np.uint64)
d = ctx.dev_alloc(pin.nbytes)
for name, src in (("pinned", pin), ("pageable", pag)):
    ctx.h2d(d, src)
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        ctx.h2d(d, src)
        ts.append(time.perf_counter() - t0)
    t = min(ts)
    print(f"{name}: {pin.nbytes / 2**20:.1f} MiB in {t * 1e3:.2f} ms = {pin.nbytes / t / 1e9:.1f} GB/s", flush=True)
