"""One-off full-size parity check: device setup + prove at 2^log_n against the closed-form proof from the
toxic waste (oracle/closed_form.py), single and batch.  Usage: python tests/check_full_size.py [log_n]"""
import importlib
import os
import random
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
zk = importlib.import_module("zksnark-rs_b200")
zg = importlib.import_module("zksnark-rs_b200.groth16")
from oracle import closed_form as cf, synthetic  # noqa: E402  (checker only)
from oracle.fields import FR  # noqa: E402

P = FR.p


def main():
    lg = int(sys.argv[1]) if len(sys.argv) > 1 else 22
    n = 1 << lg
    rng = random.Random(lg)
    t0 = time.perf_counter()
    m, n_input, rows = zg.horner_qap_rows(n)
    x, cs = rng.randrange(1, P), [rng.randrange(P) for _ in range(n)]
    wit = zg.horner_witness(n, x, cs)
    toxic = tuple(rng.randrange(1, P) for _ in range(5))
    r, s = rng.randrange(1, P), rng.randrange(1, P)
    ctx = zk.Context(0)
    q = zk.QAP(ctx, n, m, n_input, rows)
    t1 = time.perf_counter()
    crs = zk.setup(ctx, q, toxic)
    t2 = time.perf_counter()
    w = zg.fr_limbs(wit)
    got = zk.prove(ctx, q, crs, w, r, s)
    t3 = time.perf_counter()
    for _ in range(3):
        zk.prove(ctx, q, crs, w, r, s)
    t4 = time.perf_counter()
    batch = zk.prove_batch(ctx, q, crs, [w] * 4, [r] * 4, [s] * 4)
    t5 = time.perf_counter()
    print(f"2^{lg}: host prep {t1 - t0:.1f} s, setup {t2 - t1:.2f} s, first prove {t3 - t2:.3f} s, "
          f"prove {(t4 - t3) / 3 * 1e3:.2f} ms, batch {(t5 - t4) / 4 * 1e3:.2f} ms/proof (incl. lane warm-up)", flush=True)
    rows_of = []
    for ptr, gate, coeff in rows:
        vals = zg.limbs_to_ints(coeff)
        rows_of.append([[(int(gate[e]), vals[e]) for e in range(int(ptr[i]), int(ptr[i + 1]))] for i in range(m)])
    want = cf.expected_proof(n, synthetic.omega(lg), rows_of[0], rows_of[1], rows_of[2], n_input, wit, toxic, r, s)
    assert (got.a, got.b, got.c) == want, "proof != closed form"
    assert all((p.a, p.b, p.c) == want for p in batch), "batch proof != closed form"
    print(f"2^{lg}: proof bit-exact vs closed form (single and batch); checker took {time.perf_counter() - t5:.1f} s")


if __name__ == "__main__":
    main()
