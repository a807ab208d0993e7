#!/bin/bash
# NTT block size (ZKB_NTT_THREADS): 1024-element tiles with 128 / 96 / 64 threads -- blocks per SM vs the 1024-tile grid at 2^20.
tag=${1:-nttt}
mkdir -p gpurun_out
L=gpurun_out/${tag}.log
: > $L
for round in 1 2; do
for thr in 0 96 64; do
  echo "== round $round ZKB_NTT_THREADS=$thr" >> $L
  ZKB_NTT_THREADS=$thr timeout 120 python - >> $L 2>&1 <<'PY'
import ctypes as C, importlib, sys, os
import numpy as np
sys.path.insert(0, os.getcwd())
zk = importlib.import_module("zksnark-rs_b200")
ctx = zk.Context(0)
rng = np.random.default_rng(1)
for lg in (18, 20, 22):
    n = 1 << lg
    a = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64); a[:, 3] &= np.uint64((1 << 60) - 1)
    d = ctx.dev_alloc(a.nbytes); ctx.h2d(d, a)
    for _ in range(3): ctx.check(ctx.lib.zkb_ntt_fr_raw(ctx.h, C.c_void_p(d), lg, 0), "ntt")
    ctx.profile(True)
    reps = 50 if lg <= 20 else 20
    for _ in range(reps): ctx.check(ctx.lib.zkb_ntt_fr_raw(ctx.h, C.c_void_p(d), lg, 0), "ntt")
    ms, cnt, units = ctx.profile_read(1)
    ctx.profile(False)
    b = np.empty_like(a); ctx.d2h(b, d)
    print(f"2^{lg}: {ms / reps:.4f} ms per transform ({cnt // reps} passes)  {64 * n / (ms / reps * 1e-3) / 1e9:.1f} GB/s  checksum {int(b.sum(dtype=np.uint64)):x}", flush=True)
    ctx.dev_free(d)
PY
done
done
cat $L
