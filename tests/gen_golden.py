"""Generate tests/golden/proofs.json: proofs, CRS samples and h(x) for fixed secrets, computed by Oracle A (the literal
CPU restatement of the reference's setup()/prove(), oracle/groth16.py) -- the reference itself cannot run in this
image (Rust, no cargo).  The fixtures pin BOTH sides: `-m "not gpu"` tests re-derive them with Oracle A and Oracle B
(so the oracle cannot drift silently), `-m gpu` tests compare the device output with the committed numbers.
Cases: BASELINE config #1 (test_programs/simple.zk text, parser row order, roots 1..=n), the single-gate QAP of
fr.rs:248-271, a Horner circuit on the roots of unity (valid and invalid witness), the quadratic share of
mod.rs:635-690.
Usage: python tests/gen_golden.py   (rewrites tests/golden/proofs.json; ~1 min)"""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import circuit, groth16 as og, synthetic  # noqa: E402
from oracle.fields import FR  # noqa: E402

P = FR.p
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_oracle_kats import SIMPLE  # noqa: E402  (the text of test_programs/simple.zk, as the KAT tests hold it)


def rep_json(rep):
    return {"u": rep.u, "v": rep.v, "w": rep.w, "roots": rep.roots, "input": rep.input}


def case(name, rep, wit, seed, cite):
    rng = random.Random(seed)
    toxic = [rng.randrange(1, P) for _ in range(5)]
    r, s = rng.randrange(1, P), rng.randrange(1, P)
    B = og.BN254Backend()
    dense = og.qap_from_root_rep(FR, rep)
    s1, s2 = og.setup(B, dense, tuple(toxic))
    pr = og.prove(B, dense, (s1, s2), wit, r, s)
    u, v, w = og.weighted_sums(FR, dense, wit)
    h = og.quotient_h(FR, dense, u, v, w)
    pub = wit[1:rep.input + 1]
    return {"name": name, "reference": cite, "rep": rep_json(rep), "weights": wit, "toxic": toxic, "r": r, "s": s,
            "h": h, "u_sum": u, "v_sum": v,
            "crs": {"alpha1": s1.alpha, "beta1": s1.beta, "delta1": s1.delta, "xi1": s1.xi, "xi_t": s1.xi_t,
                    "sum_gamma": s1.sum_gamma, "sum_delta": s1.sum_delta, "beta2": s2.beta, "gamma2": s2.gamma,
                    "delta2": s2.delta, "xi2": s2.xi},
            "proof": {"a": pr.a, "b": pr.b, "c": pr.c},
            "inputs": pub, "verify": og.verify(B, (s1, s2), pub, pr)}


def main():
    one = FR.from_usize(1)
    cases = []
    rep = circuit.try_parse(FR, SIMPLE)
    wit = circuit.weights(FR, SIMPLE, [3, 2, 4])
    assert wit == [1, 2, 34, 6, 3, 4]
    cases.append(case("simple_zk", rep, wit, 1, "test_programs/simple.zk; circuit/mod.rs:230-526, 759-768 (BASELINE config 1)"))
    root = (-250) % P
    rep = circuit.DummyRep(u=[[], [], [(root, one)], []], v=[[], [], [], [(root, one)]], w=[[], [(root, one)], [], []],
                           roots=[root], input=2)
    cases.append(case("single_mult_honest_bn", rep, [1, 51, 3, 17], 2, "groth16/fr.rs:248-271"))
    n = 8
    w8 = synthetic.omega(3)
    rep = synthetic.horner_rep(FR, n, [pow(w8, k, P) for k in range(n)])
    rng = random.Random(33)
    wit = synthetic.horner_witness(FR, n, rng.randrange(1, P), [rng.randrange(P) for _ in range(n)])
    cases.append(case("horner8_omega", rep, wit, 3, "test_programs/deg_15.zk family (fr.rs:361-416) on the 8th roots of unity"))
    bad = list(wit)
    bad[5] = (bad[5] + 3) % P
    cases.append(case("horner8_omega_invalid_witness", rep, bad, 4, "same; h is the quotient with the remainder dropped (coefficient_poly.rs:155)"))
    rep = circuit.DummyRep(
        u=[[(3, one)], [(1, one), (2, one)], [], [], [], [], [], []],
        v=[[], [], [], [(1, one)], [(2, one)], [(3, one)], [(2, one)], [(3, one)]],
        w=[[], [], [(3, one)], [], [], [], [(1, one)], [(2, one)]], roots=[1, 2, 3], input=2)
    x, a, b, c = (rng.randrange(1, P) for _ in range(4))
    share = (a * x * x + b * x + c) % P
    wit = [1, x, share, a, b, c, a * x % P, x * ((a * x + b) % P) % P]
    cases.append(case("quad_share_roots_123", rep, wit, 5, "groth16/mod.rs:635-690 (qap_from_roots)"))
    out = os.path.join(ROOT, "tests", "golden", "proofs.json")
    with open(out, "w") as f:
        json.dump({"generator": "tests/gen_golden.py (Oracle A)", "field_modulus": P, "cases": cases}, f, indent=0)
    print(out, os.path.getsize(out), "bytes;", [(c["name"], c["verify"]) for c in cases])


if __name__ == "__main__":
    main()
