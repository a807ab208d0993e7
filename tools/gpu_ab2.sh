#!/bin/bash
# A/B visit: twiddle tiles staged by TMA bulk copies (ZKB_NTT_TMA=1) vs 32-byte gathers, and the prefetch variants of
# the bucket accumulation (tools/build_variant.py pf1 / pf2), all bit-exactness-checked by the GPU suite.
tag=${1:-ab2}
mkdir -p gpurun_out
L=gpurun_out/${tag}.log
: > $L
echo "== full GPU suite with ZKB_NTT_TMA=1" >> $L
ZKB_NTT_TMA=1 timeout 420 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_tma.log 2>&1; echo "pytest(tma) exit $?" >> $L
tail -3 gpurun_out/${tag}_pytest_tma.log >> $L
echo "== default path: NTT + golden subset" >> $L
timeout 150 python -m pytest tests -m gpu -x -q -k "ntt_matches or ntt_coset or golden or prove_matches" > gpurun_out/${tag}_pytest_default.log 2>&1; echo "pytest(default subset) exit $?" >> $L
tail -2 gpurun_out/${tag}_pytest_default.log >> $L
for tma in 0 1; do
  echo "== NTT ZKB_NTT_TMA=$tma" >> $L
  ZKB_NTT_TMA=$tma timeout 120 python - >> $L 2>&1 <<'PY'
import ctypes as C, importlib, sys, os
import numpy as np
sys.path.insert(0, os.getcwd())
zk = importlib.import_module("zksnark-rs_b200")
ctx = zk.Context(0)
rng = np.random.default_rng(1)
for lg in (16, 20, 22, 24):
    n = 1 << lg
    a = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64); a[:, 3] &= np.uint64((1 << 60) - 1)
    d = ctx.dev_alloc(a.nbytes); ctx.h2d(d, a)
    for _ in range(3): ctx.check(ctx.lib.zkb_ntt_fr_raw(ctx.h, C.c_void_p(d), lg, 0), "ntt")
    ctx.profile(True)
    for _ in range(20): ctx.check(ctx.lib.zkb_ntt_fr_raw(ctx.h, C.c_void_p(d), lg, 0), "ntt")
    ms, cnt, units = ctx.profile_read(1)
    ctx.profile(False)
    b = np.empty_like(a); ctx.d2h(b, d)
    print(f"2^{lg}: {ms / 20:.4f} ms per transform ({cnt // 20} passes)  {64 * n / (ms / 20 * 1e-3) / 1e9:.1f} GB/s  checksum {int(b.sum(dtype=np.uint64)):x}", flush=True)
    ctx.dev_free(d)
PY
done
for v in "default:" "tma:ZKB_NTT_TMA=1" "pf1:ZKB200_LIB=zksnark-rs_b200/_var/libzkb200_pf1.so" "pf2:ZKB200_LIB=zksnark-rs_b200/_var/libzkb200_pf2.so"; do
  name=${v%%:*}; envs=${v#*:}
  echo "== quick_prove 20 [$name]" >> $L
  env $envs timeout 90 python tools/quick_prove.py 20 10 >> $L 2>&1
done
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench exit $?" >> $L
echo "== msm 2^20 pf1 vs default" >> $L
timeout 60 python tools/msm_bench.py 20 >> $L 2>&1
ZKB200_LIB=zksnark-rs_b200/_var/libzkb200_pf1.so timeout 60 python tools/msm_bench.py 20 >> $L 2>&1
cat $L
cat gpurun_out/${tag}_bench.json
