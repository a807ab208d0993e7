#!/bin/bash
# NTT twiddle staging modes (ZKB_NTT_TMA = 0 gathers | 1 tiles in every pass | 2 tiles in the low passes), two rounds each.
tag=${1:-nttm}
mkdir -p gpurun_out
L=gpurun_out/${tag}.log
: > $L
for round in 1 2; do
for tma in 0 1 2; do
  echo "== round $round ZKB_NTT_TMA=$tma" >> $L
  ZKB_NTT_TMA=$tma timeout 120 python - >> $L 2>&1 <<'PY'
import ctypes as C, importlib, sys, os
import numpy as np
sys.path.insert(0, os.getcwd())
zk = importlib.import_module("zksnark-rs_b200")
ctx = zk.Context(0)
rng = np.random.default_rng(1)
for lg in (12, 16, 18, 20, 22, 24):
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
ZKB_NTT_TMA=2 timeout 200 python -m pytest tests -m gpu -x -q -k "ntt or prove_matches or golden or closed_form" > gpurun_out/${tag}_pytest_mode2.log 2>&1; echo "pytest(mode 2 subset) exit $?" >> $L
tail -2 gpurun_out/${tag}_pytest_mode2.log >> $L
for tma in 0 2; do
  echo "== quick_prove 16 ZKB_NTT_TMA=$tma" >> $L
  ZKB_NTT_TMA=$tma timeout 90 python tools/quick_prove.py 16 20 >> $L 2>&1
done
cat $L
