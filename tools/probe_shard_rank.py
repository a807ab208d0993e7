"""Developer probe (one GPU): the MSM work of ONE rank of a G-rank proof, timed kernel by kernel -- the CRS shard of rank 0
of `world` (contiguous-range layout; same point counts and window sizes as the sharded layout) through zkb_prove_partial
with the library's per-launch event trace.  Usage: python tools/probe_shard_rank.py [log_n] [world ...]"""
import csv
import importlib
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
zk = importlib.import_module("zksnark-rs_b200")


def main():
    lg = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    worlds = [int(x) for x in sys.argv[2:]] or [1, 2, 8]
    n = 1 << lg
    ctx = zk.Context(0)
    q = zk.QAP.horner(ctx, n)
    rng = np.random.default_rng(1)
    w = rng.integers(0, 1 << 63, size=(2 * n + 2, 4), dtype=np.uint64)
    w[:, 3] &= np.uint64((1 << 60) - 1)
    d_w = ctx.dev_alloc(w.nbytes)
    ctx.h2d(d_w, w)
    for world in worlds:
        crs = zk.setup(ctx, q, (3, 5, 7, 11, 13), rank=0, world=world)
        for _ in range(3):
            zk.prove_partial(ctx, q, crs, d_w, 17, 19, on_device=True)
        ctx.profile(2)
        zk.prove_partial(ctx, q, crs, d_w, 17, 19, on_device=True)
        path = os.path.join(tempfile.gettempdir(), f"probe_{world}.csv")
        ctx.trace_dump(path)
        ctx.profile(0)
        rows = list(csv.DictReader(open(path)))
        acc = [float(r["dur_ms"]) for r in rows if "accumulate" in r["kernel"]]
        heads = [float(r["dur_ms"]) for r in rows if "fix_heads" in r["kernel"]]
        lv = sum(float(r["dur_ms"]) for r in rows if any(s in r["kernel"] for s in ("bucket_", "level_quad", "horner")))
        end = max(float(r["end_ms"]) for r in rows)
        print(f"2^{lg} rank 0 of {world}: accumulate G2 {acc[0]:.3f} ms, G1 {acc[1]:.3f} ms (x{world}: {acc[0] * world:.2f} / {acc[1] * world:.2f}); "
              f"fix_heads {heads[0]:.3f} / {heads[1]:.3f}; bucket levels total {lv:.3f}; one proof {end:.3f} ms", flush=True)
        if os.environ.get("PROBE_BATCH"):
            import time
            k = int(os.environ["PROBE_BATCH"])
            zk.prove_batch(ctx, q, crs, [d_w] * 4, [17] * 4, [19] * 4, on_device=True)
            t0 = time.perf_counter()
            zk.prove_batch(ctx, q, crs, [d_w] * k, [17] * k, [19] * k, on_device=True)
            print(f"2^{lg} rank 0 of {world}: batch of {k} (proxy: full polynomial stage on this rank): {(time.perf_counter() - t0) / k * 1e3:.3f} ms/proof", flush=True)
        crs.free()


if __name__ == "__main__":
    main()
