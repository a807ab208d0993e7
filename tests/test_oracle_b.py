"""Pin Oracle B (oracle/oracle_b.c, the timed CPU restatement) against Oracle A (Python), which in
turn reproduces the reference's golden vectors (tests/test_oracle_kats.py).  CPU only."""

import random

import pytest

from oracle import bn254 as bn, circuit, groth16 as g, oracle_b as ob, poly, synthetic
from oracle.fields import FR, Q_MODULUS

from test_oracle_kats import SIMPLE

P = FR.p


def test_field_and_group_primitives():
    rng = random.Random(1)
    for _ in range(20):
        a, b = rng.randrange(P), rng.randrange(P)
        assert ob.fr_mul(a, b) == a * b % P
        assert ob.fq_mul(a, b) == a * b % Q_MODULUS
    assert ob.fr_inv(5) == pow(5, -1, P) and ob.fr_mul(P - 1, P - 1) == 1
    for _ in range(3):
        a, b = rng.randrange(1, P), rng.randrange(1, P)
        A1, B1 = bn.g1_mul(bn.BASE_G1, a), bn.g1_mul(bn.BASE_G1, b)
        assert ob.g1_mul(bn.BASE_G1, a) == A1
        assert ob.g1_add(A1, B1) == bn.g1_add(A1, B1)
        assert ob.g1_add(A1, A1) == bn.g1_add(A1, A1) and ob.g1_add(A1, bn.g1_neg(A1)) is None
        assert ob.g1_add(A1, None) == A1 and ob.g1_add(None, None) is None
        A2, B2 = bn.g2_mul(bn.BASE_G2, a), bn.g2_mul(bn.BASE_G2, b)
        assert ob.g2_mul(bn.BASE_G2, a) == A2
        assert ob.g2_add(A2, B2) == bn.g2_add(A2, B2) and ob.g2_add(A2, A2) == bn.g2_add(A2, A2)
        assert ob.g2_add(A2, bn.g2_neg(A2)) is None
    assert ob.g1_mul(bn.BASE_G1, 0) is None and ob.g1_mul(bn.G1_GEN, 69) == bn.BASE_G1  # fr.rs:106-109
    assert ob.g2_mul(bn.G2_GEN, 96) == bn.BASE_G2  # fr.rs:110-113
    pts = [bn.g1_mul(bn.BASE_G1, k) for k in (3, 4, 5)] + [None]
    sc = [7, 0, P - 1, 9]
    assert ob.msm_g1(sc, pts) == bn.msm_g1(sc, pts)


def test_poly_ops_follow_reference_quirks():
    # coefficient_poly.rs:336-367 dummy_mul vector (values small enough to hold over Fr as well)
    assert ob.poly_mul([4, 5, 6], [1, 2, 3, 0]) == [4, 13, 28, 27, 18]
    assert ob.poly_mul([0, 0], [1, 2]) == [0]
    rng = random.Random(2)
    for _ in range(10):
        a = [rng.randrange(P) for _ in range(rng.randrange(1, 9))] + [0] * rng.randrange(0, 3)
        b = [rng.randrange(P) for _ in range(rng.randrange(1, 6))] + [0] * rng.randrange(0, 2)
        assert ob.poly_mul(a, b) == poly.poly_mul(FR, a, b)
        if any(b):
            assert ob.poly_div(a, b) == poly.poly_div(FR, a, b)
    assert ob.poly_div([1, 2], [1, 2, 3]) == [0]  # field/mod.rs:443-445
    with pytest.raises(ZeroDivisionError):  # #[should_panic] field/mod.rs:679-692
        ob.poly_div([1, 2, 3], [0, 0])


def _check_prove(qap, weights, seed):
    rng = random.Random(seed)
    B = g.BN254Backend()
    toxic = tuple(rng.randrange(1, P) for _ in range(5))
    r, s = rng.randrange(1, P), rng.randrange(1, P)
    sig = g.setup(B, qap, toxic)
    want = g.prove(B, qap, sig, weights, r, s)
    u, v, w = g.weighted_sums(FR, qap, weights)
    a, b, c, h = ob.prove(qap, sig, weights, r, s)
    assert (a, b, c) == (want.a, want.b, want.c)
    assert h == g.quotient_h(FR, qap, u, v, w)
    return h


def test_prove_simple_zk_config1():
    """BASELINE config #1: test_programs/simple.zk, bit-exact plumbing check, h = [r - 64]."""
    qap = g.qap_from_root_rep(FR, circuit.try_parse(FR, SIMPLE))
    weights = circuit.weights(FR, SIMPLE, [3, 2, 4])
    assert weights == [1, 2, 34, 6, 3, 4]
    assert _check_prove(qap, weights, 1) == [P - 64]


@pytest.mark.parametrize("valid", [True, False])
def test_prove_horner_omega_domain(valid):
    n = 4
    w = synthetic.omega(2)
    rep = synthetic.horner_rep(FR, n, [pow(w, k, P) for k in range(n)])
    qap = g.qap_from_root_rep(FR, rep)
    rng = random.Random(9)
    wit = synthetic.horner_witness(FR, n, rng.randrange(1, P), [rng.randrange(P) for _ in range(n)])
    if not valid:
        wit[5] = (wit[5] + 3) % P
    _check_prove(qap, wit, 3 + valid)
    _check_prove(qap, wit[:-1], 5)  # zip truncation
