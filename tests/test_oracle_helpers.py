"""Pin the fast/closed-form oracle helpers (used by the full-size GPU parity tests) against the literal
restatement of the reference.  CPU only."""

import random

import pytest

from oracle import closed_form as cf, groth16 as g, poly, synthetic
from oracle.fields import FR, Z251

P = FR.p


@pytest.mark.parametrize("log_n", [0, 1, 2, 3, 5, 7])
def test_ntt_fast_equals_reference_dft(log_n):
    rng = random.Random(log_n)
    n = 1 << log_n
    w = synthetic.omega(log_n)
    x = [rng.randrange(P) for _ in range(n)]
    assert poly.ntt_fast(FR, x, w) == poly.dft(FR, x, w)        # field/mod.rs:508-520
    assert poly.intt_fast(FR, x, w) == poly.idft(FR, x, w)      # field/mod.rs:524-537


def test_ntt_fast_z251():
    # 5 has order 25 in F_251 (dft_test, field/mod.rs:606-623); 5^? of order 2^k: 250 = 2 * 125 -> only n = 2
    x = [17, 200]
    assert poly.ntt_fast(Z251, x, 250) == poly.dft(Z251, x, 250)


@pytest.mark.parametrize("n", [2, 4, 8])
def test_closed_form_proof_equals_literal_prove(n):
    rng = random.Random(n)
    log_n = n.bit_length() - 1
    w = synthetic.omega(log_n)
    roots = [pow(w, k, P) for k in range(n)]
    rep = synthetic.horner_rep(FR, n, roots)
    qap = g.qap_from_root_rep(FR, rep)
    wit = synthetic.horner_witness(FR, n, rng.randrange(1, P), [rng.randrange(P) for _ in range(n)])
    toxic = tuple(rng.randrange(1, P) for _ in range(5))
    r, s = rng.randrange(1, P), rng.randrange(1, P)
    B = g.BN254Backend()
    pr = g.prove(B, qap, g.setup(B, qap, toxic), wit, r, s)
    idx = {rt: k for k, rt in enumerate(roots)}
    rows = lambda mat: [[(idx[a], c) for a, c in row] for row in mat]
    want = cf.expected_proof(n, w, rows(rep.u), rows(rep.v), rows(rep.w), rep.input, wit, toxic, r, s)
    assert want == (pr.a, pr.b, pr.c)


def test_h_is_high_half_of_product_on_unity_domain():
    """on the omega domain t = x^n - 1 and the reference quotient is p[n:], for ANY witness."""
    n = 8
    rng = random.Random(5)
    u = [rng.randrange(P) for _ in range(n)]
    v = [rng.randrange(P) for _ in range(n)]
    wv = [rng.randrange(P) for _ in range(n)]
    t = [P - 1] + [0] * (n - 1) + [1]
    h = poly.poly_div(FR, poly.poly_sub(FR, poly.poly_mul(FR, u, v), wv), t)
    assert h == poly.poly_mul(FR, u, v)[n:]
