"""Timing of witness generation on the device (zkb_witness_generate, values and result resident in HBM) on layered
circuits of 2^20 gates with different aspect ratios and on the depth-n Horner chain.  Prints one JSON line per case."""
import importlib
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
zk = importlib.import_module("zksnark-rs_b200")
zg = importlib.import_module("zksnark-rs_b200.groth16")


def timed(ctx, plan, d_vals, n_vals, d_out, reps):
    zk.witness_generate_dev(ctx, plan, d_vals, n_vals, d_out)
    t0 = time.perf_counter()
    for _ in range(reps):
        zk.witness_generate_dev(ctx, plan, d_vals, n_vals, d_out)  # synchronous on return
    return (time.perf_counter() - t0) / reps * 1e3


def main():
    ctx = zk.Context(0)
    rng = np.random.default_rng(1)
    cases = [("layered", 1 << 18, 4), ("layered", 1 << 16, 16), ("layered", 1 << 14, 64), ("layered", 1 << 12, 256), ("layered", 1 << 10, 1024),
             ("layered", 1 << 8, 4096), ("layered-unit", 1 << 18, 4), ("layered-unit", 1 << 12, 256), ("layered-unit", 1 << 8, 4096),
             ("horner", 1 << 16, None), ("horner", 1 << 20, None)]
    for kind, a, b in cases:
        if kind.startswith("layered"):
            n, m, n_input, rows, free = zg.layered_qap_rows(a, b, seed=3, unit_coeffs=kind.endswith("unit"))
            qap = zk.QAP(ctx, n, m, n_input, rows)
        else:
            n = a
            m, n_input, rows = zg.horner_qap_rows(n)
            qap = zk.QAP(ctx, n, m, n_input, rows)
            free = [1] + [2 * k + 2 for k in range(1, n)] + [2 * n + 1]
        t0 = time.perf_counter()
        plan = zk.WitnessPlan(ctx, qap, free)
        t_plan = (time.perf_counter() - t0) * 1e3
        vals = rng.integers(0, 1 << 62, size=(len(free), 4), dtype=np.uint64)
        vals[:, 3] >>= 4  # < 2^250 < r
        d_vals, d_out = ctx.dev_alloc(vals.nbytes), ctx.dev_alloc(32 * m)
        ctx.h2d(d_vals, vals)
        ms = timed(ctx, plan, d_vals, len(free), d_out, 3 if kind == "horner" else 10)
        info = plan.info()
        print(json.dumps({"circuit": kind, "gates": n, "wires": m, **info, "plan_ms": round(t_plan, 1), "generate_ms": round(ms, 3),
                          "gates_per_s": round(n / ms * 1e3), "us_per_level": round(ms * 1e3 / info["n_levels"], 3)}), flush=True)
        ctx.dev_free(d_vals)
        ctx.dev_free(d_out)
        plan.free()
        qap.free()
    ctx.close()


if __name__ == "__main__":
    main()
