#!/bin/bash
# A/B of NTT kernel variants (tools/build_variant.py) + one ncu --set full capture of both accumulation kernels.
tag=${1:-ab}
mkdir -p gpurun_out
for lib in "" zksnark-rs_b200/_var/libzkb200_ntt6.so zksnark-rs_b200/_var/libzkb200_ntt7.so zksnark-rs_b200/_var/libzkb200_nttu3.so zksnark-rs_b200/_var/libzkb200_nttu4.so; do
  echo "== ${lib:-default}" >> gpurun_out/${tag}_ntt_ab.log
  ZKB200_LIB=$lib timeout 120 python - >> gpurun_out/${tag}_ntt_ab.log 2>&1 <<'PY'
import ctypes as C, importlib, sys, time, os
import numpy as np
sys.path.insert(0, os.getcwd())
zk = importlib.import_module("zksnark-rs_b200")
ctx = zk.Context(0)
rng = np.random.default_rng(1)
for lg in (16, 20, 22):
    n = 1 << lg
    a = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64); a[:, 3] &= np.uint64((1 << 60) - 1)
    d = ctx.dev_alloc(a.nbytes); ctx.h2d(d, a)
    for _ in range(3): ctx.check(ctx.lib.zkb_ntt_fr_raw(ctx.h, C.c_void_p(d), lg, 0), "ntt")
    ctx.profile(True)
    for _ in range(20): ctx.check(ctx.lib.zkb_ntt_fr_raw(ctx.h, C.c_void_p(d), lg, 0), "ntt")
    ms, cnt, units = ctx.profile_read(1)
    ctx.profile(False)
    print(f"2^{lg}: {ms / 20:.4f} ms per transform ({cnt // 20} passes)  {64 * n / (ms / 20 * 1e-3) / 1e9:.1f} GB/s", flush=True)
    ctx.dev_free(d)
PY
done
cat gpurun_out/${tag}_ntt_ab.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_accumulate_chunks -s 8 -c 2 -f -o gpurun_out/${tag}_acc python bench.py --steps 2 --warmup 1 > gpurun_out/${tag}_full_acc.log 2>&1; echo "ncu full acc exit $?"
ls -la gpurun_out/${tag}_acc.ncu-rep
