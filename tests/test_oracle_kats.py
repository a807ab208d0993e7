"""Pin Oracle A against every golden vector / known answer the reference's own tests
hold for the generic code on the prove() path.  CPU only.

Each test cites the reference test it restates (paths relative to /root/reference).
"""

import random

import pytest

from oracle import circuit, groth16 as g
from oracle.fields import FR, Z251, FieldPanic
from oracle.poly import (degree, dft, evaluate, idft, lagrange_basis, poly_add, poly_div, poly_from_points,
                         poly_mul, poly_neg, poly_scale, poly_sub, poly_sum, polynomial_division,
                         powers, root_poly)

Z = Z251


def zs(xs):
    return [Z.from_usize(x) for x in xs]


# ---------------------------------------------------------------- src/field/z251.rs:103-150
def test_z251_exhaustive_inverse_and_neg():
    for a in range(251):
        assert Z.add(a, Z.neg(a)) == 0
        if a:
            assert Z.mul(a, Z.mul_inv(a)) == 1
    assert Z.neg(0) == 251  # non-canonical -0, z251.rs:24-28
    with pytest.raises(FieldPanic):
        Z.from_usize(251)  # z251.rs:80


# ---------------------------------------------------------------- src/field/mod.rs tests
def test_powers_test():  # field/mod.rs:591-604
    assert powers(Z, 9, 5) == [1, 9, 81, 227, 35]


DFT_GOLDEN = [6, 86, 169, 189, 203, 131, 237, 118, 115, 91, 248, 177, 8, 48, 34, 136, 177, 203,
              125, 57, 237, 81, 9, 30, 122]


def test_dft_test():  # field/mod.rs:606-623
    seq = [0] * 25
    seq[0], seq[1], seq[2] = 1, 2, 3
    assert dft(Z, seq, 5) == DFT_GOLDEN


def test_idft_test():  # field/mod.rs:625-635
    seq = [0] * 25
    seq[0], seq[1], seq[2] = 1, 2, 3
    assert idft(Z, dft(Z, seq, 5), 5) == seq


def test_degree_test():  # field/mod.rs:637-655
    assert degree(Z, zs([3, 0, 0, 0, 179, 0, 0, 6])) == 7
    assert degree(Z, zs([29, 112, 68])) == 2
    assert degree(Z, zs([3, 0, 0, 0, 179, 0, 0, 6] + [0] * 7)) == 7
    assert degree(Z, []) == 0 and degree(Z, [0, 0]) == 0  # field/mod.rs:291-297


def test_polynomial_division_test():  # field/mod.rs:657-677
    q, r = polynomial_division(Z, zs([3, 0, 0, 0, 179, 0, 0, 6]), zs([29, 112, 68]))
    assert q == [209, 207, 78, 1, 131, 37]
    assert r == [217, 207]


def test_polynomial_division_doc_example():  # field/mod.rs:414-427
    q, r = polynomial_division(Z, zs([1, 0, 3, 1]), zs([0, 0, 9, 1]))
    assert (q, r) == ([1], [1, 0, 245])


def test_polynomial_divisionby0_test():  # field/mod.rs:679-692 (#[should_panic])
    with pytest.raises(FieldPanic):
        polynomial_division(Z, zs([3, 0, 0, 0, 179, 0, 0, 6]), [0] * 8)


def test_division_early_return():  # field/mod.rs:443-445
    assert polynomial_division(Z, zs([1, 2]), zs([1, 2, 3])) == ([0], [0])


def test_evaluate_doc_examples():  # field/mod.rs:300-337
    assert evaluate(Z, [1, 1, 1], 2) == 7
    assert evaluate(Z, [1, 1, 4], 2) == 19
    assert evaluate(Z, [1, 2, 3, 4], 3) == 142


# ---------------------------------------------------------------- src/groth16/coefficient_poly.rs tests
def test_dummy_add():  # coefficient_poly.rs:221-258
    assert all(c == 0 for c in poly_add(Z, [], []))
    assert poly_add(Z, [], [1, 2, 3]) == [1, 2, 3]
    assert poly_add(Z, [0], [1, 2, 3]) == [1, 2, 3]
    assert poly_add(Z, [4, 5, 6], [1, 2, 3, 0]) == [5, 7, 9, 0]
    assert poly_add(Z, [234, 100, 6], [123, 234, 3]) == [106, 83, 9]


def test_dummy_neg_sub_sum():  # coefficient_poly.rs:260-318
    rng = random.Random(1)
    for _ in range(200):
        a = [rng.randrange(1, 251) for _ in range(3)]
        b = [rng.randrange(1, 251) for _ in range(3)]
        assert all(c == 0 for c in poly_add(Z, a, poly_neg(Z, a)))
        assert poly_add(Z, b, poly_sub(Z, a, b)) == a
    polys = [[rng.randrange(1, 251) for _ in range(3)] for _ in range(20)]
    acc = [0, 0, 0]
    for p in polys:
        acc = poly_add(Z, acc, p)
    assert poly_sum(Z, polys) == acc


def test_dummy_mul():  # coefficient_poly.rs:320-368
    assert all(c == 0 for c in poly_mul(Z, [], []))
    assert all(c == 0 for c in poly_mul(Z, [], [1, 2, 3]))
    assert all(c == 0 for c in poly_mul(Z, [0], [1, 2, 3]))
    assert poly_mul(Z, [4, 5, 6], [1, 2, 3, 0]) == [4, 13, 28, 27, 18]
    assert poly_mul(Z, [234, 100, 6], [123, 234, 3]) == [168, 39, 242, 198, 18]


def test_dummy_scalar_mul():  # coefficient_poly.rs:370-403
    assert poly_scale(Z, [], 69) == []
    assert poly_scale(Z, [0], 69) == [0]
    assert poly_scale(Z, [1, 2, 3], 69) == [69, 138, 207]
    assert poly_scale(Z, [20, 2, 3], 69) == [125, 138, 207]
    assert all(c == 0 for c in poly_scale(Z, [20, 2, 3], 0))


def test_dummy_div():  # coefficient_poly.rs:405-429
    rng = random.Random(2)
    for _ in range(300):
        a = [rng.randrange(1, 251) for _ in range(3)]
        b = [rng.randrange(1, 251) for _ in range(3)]
        assert poly_div(Z, poly_mul(Z, a, b), b) == a


def test_dummy_lagrange():  # coefficient_poly.rs:431-446
    for mx in range(2, 25):
        for i in range(1, mx):
            p = lagrange_basis(Z, zs(range(1, mx)), i)
            for j in range(1, mx):
                assert evaluate(Z, p, j) == (1 if i == j else 0)


def test_dummy_from_roots():  # coefficient_poly.rs:448-468
    for mask in range(1, 255):
        pts = [(i + 1, i + 2) for i in range(8) if (1 << i) & mask]
        p = poly_from_points(Z, zs(range(1, 9)), pts)
        for i in range(8):
            assert evaluate(Z, p, i + 1) == ((i + 2) if (1 << i) & mask else 0)


def test_dummy_root_poly():  # coefficient_poly.rs:470-479
    for i in range(2, 25):
        p = root_poly(Z, zs(range(1, i)))
        for j in range(1, i):
            assert evaluate(Z, p, j) == 0


# ---------------------------------------------------------------- parser, circuit/mod.rs:664-769
QUAD = """(in x a b c)
(out y)
(verify x y)

(program
    (= t1
        (* x a))
    (= t2
        (* x (+ t1 b)))
    (= y
        (* 1 (+ t2 c))))"""

SIMPLE = """(in a b c)
(out x)
(verify b x)

(program
    (= temp
        (* a b))
    (= x
        (* 1 (+ (* 4 temp) c 6))))"""


def test_try_parse_impl_test():  # circuit/mod.rs:664-720
    rep = circuit.try_parse(Z, QUAD)
    assert rep.u == [[(3, 1)], [(1, 1), (2, 1)], [], [], [], [], [], []]
    assert rep.v == [[], [], [], [(2, 1)], [(1, 1)], [(3, 1)], [(2, 1)], [(3, 1)]]
    assert rep.w == [[], [], [(3, 1)], [(1, 1)], [], [(2, 1)], [], []]
    assert rep.roots == [1, 2, 3] and rep.input == 2


def test_weights_test():  # circuit/mod.rs:746-768
    assert circuit.weights(Z, SIMPLE, [3, 2, 4]) == [1, 2, 34, 6, 3, 4]


# ---------------------------------------------------------------- groth16/mod.rs tests (Z251 fake curve)
QS_U = [[1, 124, 126], [0, 127, 125]] + [[0, 0, 0]] * 6
QS_V = [[0, 0, 0]] * 3 + [[3, 123, 126], [248, 4, 250], [1, 124, 126], [248, 4, 250], [1, 124, 126]]
QS_W = [[0, 0, 0]] * 2 + [[1, 124, 126]] + [[0, 0, 0]] * 3 + [[3, 123, 126], [248, 4, 250]]
QS_T = [245, 11, 245, 1]

QUAD_REP = circuit.DummyRep(  # groth16/mod.rs:637-670
    u=[[(3, 1)], [(1, 1), (2, 1)], [], [], [], [], [], []],
    v=[[], [], [], [(1, 1)], [(2, 1)], [(3, 1)], [(2, 1)], [(3, 1)]],
    w=[[], [], [(3, 1)], [], [], [], [(1, 1)], [(2, 1)]],
    roots=[1, 2, 3], input=2)


def _nz(rng):
    return rng.randrange(1, 251)


def test_single_mult_honest():  # groth16/mod.rs:383-426 (CRS structure asserts :403-416)
    B = g.Z251Backend()
    c = lambda k: [Z.from_usize(k)]
    qap = g.QAP(u=[c(0), c(0), c(1), c(0)], v=[c(0), c(0), c(0), c(1)], w=[c(0), c(1), c(0), c(0)],
                t=[250, 1], input=2, degree=1)
    weights = [1, 17, 100, 83]
    rng = random.Random(3)
    for _ in range(200):
        toxic = tuple(_nz(rng) for _ in range(5))
        s1, s2 = g.setup(B, qap, toxic)
        alpha, beta, gamma, delta = (Z.div(s1.alpha, 69), Z.div(s1.beta, 69), Z.div(s2.gamma, 69),
                                     Z.div(s1.delta, 69))
        assert (alpha, beta, gamma, delta) == toxic[:4]
        assert s1.xi == [B.encrypt_g1(1)]
        assert s1.sum_gamma == [B.encrypt_g1(0), B.encrypt_g1(Z.div(1, gamma)),
                                B.encrypt_g1(Z.div(beta, gamma))]
        assert s1.sum_delta == [B.encrypt_g1(Z.div(alpha, delta))]
        assert s1.xi_t == []
        assert s2.xi == [B.encrypt_g2(1)]
        proof = g.prove(B, qap, (s1, s2), weights, _nz(rng), _nz(rng))
        assert g.verify(B, (s1, s2), [17, 100], proof)


def test_qap_from_roots_equals_hardcoded():  # groth16/mod.rs:472-521 == :635-672
    qap = g.qap_from_root_rep(Z, QUAD_REP)
    # empty rows interpolate to [0] (Sum seed, coefficient_poly.rs:84-89); compare zero-padded
    pad = lambda rows: [list(r) + [0] * (3 - len(r)) for r in rows]
    assert pad(qap.u) == QS_U and pad(qap.v) == QS_V and pad(qap.w) == QS_W
    assert qap.t == QS_T and qap.degree == 3 and qap.input == 2


@pytest.mark.parametrize("source", ["hardcoded", "roots", "legacy", "ast"])
def test_quadratic_share_honest(source):  # groth16/mod.rs:472-542, 635-693, 695-722, 758-790
    B = g.Z251Backend()
    if source == "hardcoded":
        qap = g.QAP(u=QS_U, v=QS_V, w=QS_W, t=QS_T, input=2, degree=3)
    elif source == "roots":
        qap = g.qap_from_root_rep(Z, QUAD_REP)
    elif source == "legacy":
        legacy = "x y\na b c\nt1 t2\n\nt1 ( x ) ( a )\nt2 ( x ) ( t1 b )\ny ( 1 ) ( t2 c )"
        qap = g.qap_from_root_rep(Z, circuit.dummy_rep_from_legacy(Z, legacy))
    else:
        qap = g.qap_from_root_rep(Z, circuit.try_parse(Z, QUAD))
    rng = random.Random(4)
    for _ in range(200):
        x, a, b, c = (_nz(rng) for _ in range(4))
        share = (a * x * x + b * x + c) % 251
        ax = a * x % 251
        t2 = x * (ax + b) % 251
        if source == "ast":
            weights = [1, x, share, ax, a, t2, b, c]  # order of first appearance, mod.rs:774-777
        else:
            weights = [1, x, share, a, b, c, ax, t2]
        sigma = g.setup(B, qap, tuple(_nz(rng) for _ in range(5)))
        proof = g.prove(B, qap, sigma, weights, _nz(rng), _nz(rng))
        assert g.verify(B, sigma, [x, share], proof)
        assert not g.verify(B, sigma, [x, (share + 1) % 251], proof)


def test_random_proof_acceptance_rate():  # groth16/mod.rs:428-470: ~1/250 of random proofs verify
    B = g.Z251Backend()
    c = lambda k: [Z.from_usize(k)]
    qap = g.QAP(u=[c(0), c(0), c(1), c(0)], v=[c(0), c(0), c(0), c(1)], w=[c(0), c(1), c(0), c(0)],
                t=[250, 1], input=2, degree=1)
    rng = random.Random(5)
    total, count = 10000, 0
    for _ in range(total):
        sigma = g.setup(B, qap, tuple(_nz(rng) for _ in range(5)))
        if g.verify(B, sigma, [17, 100], g.Proof(_nz(rng), _nz(rng), _nz(rng))):
            count += 1
    assert 0.002 < count / total < 0.006


# ---------------------------------------------------------------- config #1: simple.zk hand trace
def test_simple_zk_known_answer_fr():
    rep = circuit.try_parse(FR, SIMPLE)
    assert rep.input == 2 and rep.roots == [1, 2]
    weights = circuit.weights(FR, SIMPLE, [3, 2, 4])
    assert weights == [1, 2, 34, 6, 3, 4]
    qap = g.qap_from_root_rep(FR, rep)
    u_sum, v_sum, w_sum = g.weighted_sums(FR, qap, weights)
    r = FR.p
    assert u_sum == [5, r - 2] and v_sum == [r - 30, 32] and w_sum == [r - 22, 28]
    assert qap.t == [2, r - 3, 1]
    assert g.quotient_h(FR, qap, u_sum, v_sum, w_sum) == [r - 64]
