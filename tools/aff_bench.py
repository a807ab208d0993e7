"""Developer probe: the batched-affine pair tree (ZKB_AFF_G1 / ZKB_AFF_G2 levels, ZKB_AFF_VAR body, ZKB_AFF_K) against the
XYZZ chain alone -- stand-alone MSMs at 2^log_n (accumulation phase from the library's event brackets) and whole proofs
(one alone, batch).  The switches are read at every call, so one process sweeps them.
Usage: python tools/aff_bench.py [log_n] [--msm-only] [--prove-only]"""
import importlib
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
zk = importlib.import_module("zksnark-rs_b200")
zg = importlib.import_module("zksnark-rs_b200.groth16")


def rand_fr(rng, n):
    a = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64((1 << 60) - 1)
    return a


def mode(g1, g2, b=16, k=1, var=1):
    """levels per group; b is fixed at 16 in the library now (kept in the records of the earlier sweeps); var 0 plain / 1 staged"""
    os.environ["ZKB_AFF_G1"], os.environ["ZKB_AFF_G2"] = str(g1), str(g2)
    os.environ["ZKB_AFF_K"], os.environ["ZKB_AFF_VAR"] = str(k), str(var)


def trace_levels(ctx, fn):
    """One call under the per-launch trace: [(kernel, ms)] of the accumulation kernels in launch order."""
    import csv
    path = "/tmp/aff_trace.csv"
    ctx.profile(2)
    fn()
    ctx.trace_dump(path)
    ctx.profile(0)
    out = []
    for r in list(csv.reader(open(path)))[1:]:
        name = r[1]
        for key, short in (("k_affine_level", "lvl"), ("k_accumulate_points", "chain"), ("k_accumulate_chunks", "chain0"),
                           ("k_fix_heads", "heads"), ("k_lvl_scan", "scan")):
            if key in name:
                out.append((short, round(float(r[4]), 3)))
    return out


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    lg = int(args[0]) if args else 20
    n = 1 << lg
    ctx = zk.Context(0)
    rng = np.random.default_rng(1)
    out = []
    if "--prove-only" not in sys.argv:
        for group, sweeps in ((2, [(0, 16, 1, 0), (5, 16, 1, 0), (5, 16, 1, 1), (6, 16, 1, 1), (4, 16, 1, 1)]),
                              (1, [(0, 16, 1, 0), (4, 16, 1, 0), (4, 16, 1, 1), (3, 16, 1, 1)])):
            b = zk.Bases.generate(ctx, group, rand_fr(rng, n))
            s = rand_fr(rng, n)
            ds = ctx.dev_alloc(s.nbytes)
            ctx.h2d(ds, s)
            ref = None
            for levels, batch, k, occ in sweeps:
                mode(levels if group == 1 else 0, levels if group == 2 else 0, batch, k, occ)
                for _ in range(2):
                    r = zk.msm(ctx, b, ds, on_device=True, n=n)
                ref = ref or r
                assert r == ref, "pair tree result differs from the chain"
                ctx.profile(True)
                reps = 4
                t0 = time.perf_counter()
                for _ in range(reps):
                    zk.msm(ctx, b, ds, on_device=True, n=n)
                wall = (time.perf_counter() - t0) / reps
                ms, cnt, units = ctx.profile_read(2 if group == 1 else 3)
                ctx.profile(False)
                rec = {"what": "msm", "group": group, "log_n": lg, "levels": levels, "batch": batch, "k": k, "var": occ,
                       "call_ms": round(wall * 1e3, 3), "accumulate_ms": round(ms / reps, 3), "launches_per_call": cnt // reps,
                       "records_M": round(units / reps / 1e6, 2),
                       "trace": trace_levels(ctx, lambda: zk.msm(ctx, b, ds, on_device=True, n=n))}
                out.append(rec)
                print(json.dumps(rec), flush=True)
            ctx.dev_free(ds)
            b.free()
    if "--msm-only" not in sys.argv:
        mode(0, 0)
        q = zk.QAP.horner(ctx, n)
        crs = zk.setup(ctx, q, (3, 5, 7, 11, 13))
        w = rand_fr(rng, 2 * n + 2)
        d_w = ctx.dev_alloc(w.nbytes)
        ctx.h2d(d_w, w)
        ref = None
        for g1, g2, batch, k, occ in ((0, 0, 16, 1, 0), (0, 5, 16, 1, 1), (0, 6, 16, 1, 1), (0, 4, 16, 1, 1), (4, 5, 16, 1, 1)):
            mode(g1, g2, batch, k, occ)
            for _ in range(2):
                p = zg.prove_dev(ctx, q, crs, d_w, 17, 19)
            ref = ref or (p.a, p.b, p.c)
            assert (p.a, p.b, p.c) == ref, "proof differs from the chain-only proof"
            reps = 6
            t0 = time.perf_counter()
            for _ in range(reps):
                zg.prove_dev(ctx, q, crs, d_w, 17, 19)
            t1 = (time.perf_counter() - t0) / reps
            zk.prove_batch(ctx, q, crs, [d_w] * 4, [17] * 4, [19] * 4, on_device=True)
            reps = 12
            t0 = time.perf_counter()
            pb = zk.prove_batch(ctx, q, crs, [d_w] * reps, [17] * reps, [19] * reps, on_device=True)
            tb = (time.perf_counter() - t0) / reps
            assert all((x.a, x.b, x.c) == ref for x in pb)
            rec = {"what": "prove", "log_n": lg, "g1_levels": g1, "g2_levels": g2, "batch": batch, "k": k, "var": occ,
                   "one_proof_ms": round(t1 * 1e3, 3), "batch_ms_per_proof": round(tb * 1e3, 3)}
            out.append(rec)
            print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
