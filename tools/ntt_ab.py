"""NTT A/B probe: forward DIF transform (zkb_ntt_fr_raw) per size, timed with the library's CUDA-event brackets, plus the
per-pass kernel durations at 2^20 from the launch trace.  Env switches select the variant (ZKB_NTT_SHFL, ZKB_NTT_TMA)."""
import csv
import ctypes as C
import importlib
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
zk = importlib.import_module("zksnark-rs_b200")
ctx = zk.Context(0)
rng = np.random.default_rng(1)
for lg in (10, 12, 16, 18, 20, 22):
    n = 1 << lg
    a = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64((1 << 60) - 1)
    d = ctx.dev_alloc(a.nbytes)
    ctx.h2d(d, a)
    for _ in range(3):
        ctx.check(ctx.lib.zkb_ntt_fr_raw(ctx.h, C.c_void_p(d), lg, 0), "ntt")
    ctx.profile(True)
    reps = 50 if lg <= 20 else 20
    for _ in range(reps):
        ctx.check(ctx.lib.zkb_ntt_fr_raw(ctx.h, C.c_void_p(d), lg, 0), "ntt")
    ms, cnt, units = ctx.profile_read(1)
    ctx.profile(False)
    b = np.empty_like(a)
    ctx.d2h(b, d)
    line = f"2^{lg}: {ms / reps:.4f} ms per transform ({cnt // reps} passes)  checksum {int(b.sum(dtype=np.uint64)):x}"
    if lg == 20:
        ctx.profile(2)
        for _ in range(5):
            ctx.check(ctx.lib.zkb_ntt_fr_raw(ctx.h, C.c_void_p(d), lg, 0), "ntt")
        path = os.path.join(tempfile.gettempdir(), "ntt_trace.csv")
        ctx.trace_dump(path)
        ctx.profile(0)
        rows = [r for r in csv.DictReader(open(path)) if "ntt_pass" in r["kernel"]]
        per = {}
        for i, r in enumerate(rows):
            per.setdefault((i % (len(rows) // 5), r["kernel"]), []).append(float(r["dur_ms"]))
        line += "  passes: " + ", ".join(f"{k[1]} {sum(v) / len(v) * 1e3:.1f} us" for k, v in sorted(per.items()))
    print(line, flush=True)
    ctx.dev_free(d)
