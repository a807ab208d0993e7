"""Workload for compute-sanitizer (memcheck / racecheck / synccheck / initcheck): exercises the
multi-pass NTT (2^13: three shared-memory passes incl. strided tiles), the MSM with graded chunks,
head folding on both paths (skewed scalars -> one huge bucket -> warp-cooperative fold), the
bucket hierarchy (quad plan for single proofs, its small-batch variant in the batches), batch proving on two lanes, the
sharded fold and witness generation, at sizes a sanitizer finishes in a minute.  Results are checked against the oracle's closed forms.
Usage: compute-sanitizer --tool memcheck python tests/sanitize_case.py"""
import importlib
import os
import random
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
zk = importlib.import_module("zksnark-rs_b200")
zg = importlib.import_module("zksnark-rs_b200.groth16")
from oracle import bn254 as bn  # noqa: E402  (checker only)
from oracle.fields import FR  # noqa: E402

P = FR.p


def main():
    rng = random.Random(5)
    ctx = zk.Context(0)
    # NTT round trip at 2^13 (passes: 10 + 3 stages) and 2^11
    for lg in (11, 13):
        x = [rng.randrange(P) for _ in range(1 << lg)]
        assert zk.ntt(ctx, zk.ntt(ctx, x), inverse=True) == x
    # MSM: uniform, skewed (same scalar everywhere: one bucket takes every record), witness-like 0/1
    n = 1 << 12
    ks = [rng.randrange(P) for _ in range(n)]
    for group in (1, 2):
        b = zk.Bases.generate(ctx, group, ks)
        base, mul = (bn.BASE_G1, bn.g1_mul) if group == 1 else (bn.BASE_G2, bn.g2_mul)
        for ss in ([rng.randrange(P) for _ in range(n)], [ks[7]] * n, [i & 1 for i in range(n)]):
            e = sum(s * k for s, k in zip(ss, ks)) % P
            assert zk.msm(ctx, b, ss) == mul(base, e)
        b.free()
    # prove: batch on two lanes == single, sharded fold == single
    n = 1 << 10
    m, n_input, rows = zg.horner_qap_rows(n)
    q = zk.QAP(ctx, n, m, n_input, rows)
    toxic = tuple(rng.randrange(1, P) for _ in range(5))
    crs = zk.setup(ctx, q, toxic)
    wits = [zg.horner_witness(n, rng.randrange(1, P), [rng.randrange(P) for _ in range(n)]) for _ in range(3)]
    rs, ss = [rng.randrange(1, P) for _ in range(3)], [rng.randrange(1, P) for _ in range(3)]
    single = [zk.prove(ctx, q, crs, w, r, s) for w, r, s in zip(wits, rs, ss)]
    batch = zk.prove_batch(ctx, q, crs, wits, rs, ss)
    assert [(p.a, p.b, p.c) for p in batch] == [(p.a, p.b, p.c) for p in single]
    parts = [zk.prove_batch(ctx, q, zk.setup(ctx, q, toxic, rank=k, world=2), wits, rs, ss) for k in range(2)]
    got = zk.prove_combine_batch(ctx, np.stack(parts))
    assert [(p.a, p.b, p.c) for p in got] == [(p.a, p.b, p.c) for p in single]
    # verify (Miller loops + final exponentiation, 13 KB of per-thread stack), batch with a rejected proof; window-sharded MSM
    pubs = [w[1:3] for w in wits]
    assert zg.verify_batch(ctx, crs, pubs + [[pubs[0][0], (pubs[0][1] + 1) % P]], single + [single[0]]) == [True, True, True, False]
    b = zk.Bases.generate(ctx, 1, ks[:256])
    sc = [rng.randrange(P) for _ in range(256)]
    assert zg.points_sum(ctx, 1, [zk.msm(ctx, b, sc, windows=(g, 3)) for g in range(3)]) == zk.msm(ctx, b, sc)
    # witness generation: a chain (one block, barrier per level), wide levels (programmatic dependent launches), mixed
    from oracle import circuit  # noqa: E402  (checker only)
    for width, depth in ((1, 50), (40, 12), (700, 4)):
        n, m, n_input, rows, free = zg.layered_qap_rows(width, depth, seed=width)
        q2 = zk.QAP(ctx, max(2, 1 << (n - 1).bit_length()), m, n_input, rows)
        plan = zk.WitnessPlan(ctx, q2, free)
        vals = [rng.randrange(P) for _ in free]
        assert zk.weights(ctx, plan, vals) == circuit.weights_from_rows(P, n, m, circuit.csr_by_gate(n, m, rows), free, vals)
    print(f"sanitize case ok: {ctx.launches} kernel launches")
    ctx.close()


if __name__ == "__main__":
    main()
