"""Pin the BN254 layer of Oracle A ("parity unpinned" at the bn boundary):
published constants / known answers, group-law identities, and the only end-to-end
check the reference itself uses -- verify(...) == true (src/lib.rs:156-190,
src/groth16/fr.rs:248-302) -- through an independent optimal-ate pairing.  CPU only."""

import random

from oracle import bn254 as bn, circuit, groth16 as g, synthetic
from oracle.fields import FR

from test_oracle_kats import QUAD, SIMPLE


def test_constants_and_known_answers():
    q, r, u = bn.Q, bn.R_ORDER, bn.BN_U
    assert q == 36 * u**4 + 36 * u**3 + 24 * u**2 + 6 * u + 1
    assert r == 36 * u**4 + 36 * u**3 + 18 * u**2 + 6 * u + 1
    assert (r - 1) % (1 << 28) == 0 and (r - 1) % (1 << 29) != 0
    assert pow(5, (r - 1) // 2, r) == r - 1  # 5 is a non-residue
    assert synthetic.OMEGA_2_28 == 19103219067921713944291392827692070036145651957329286315305642004821462161904
    assert synthetic.omega(20) == 17220337697351015657950521176323262483320249231368149235373741788599650842711
    assert bn.g1_is_on_curve(bn.G1_GEN) and bn.g2_is_on_curve(bn.G2_GEN)
    assert bn.g1_mul(bn.G1_GEN, 2) == (
        1368015179489954701390400359078579693043519447331113978918064868415326638035,
        9918110051302171585080402603319702774565515993150576347155970296011118125764)
    assert bn.BASE_G1 == (
        5138697240077803445514669414784254799933862402946278134326199877546184124353,
        12587011617949543324467535889916856826666519601316494966427400843934921824601)
    assert bn.BASE_G2[0][0] == 6796222810610176583640253016064427379720569220532816811297156498654695659892
    assert bn.g1_add(bn.g1_mul(bn.G1_GEN, r - 1), bn.G1_GEN) is None
    assert bn.g2_add(bn.g2_mul(bn.G2_GEN, r - 1), bn.G2_GEN) is None


def test_group_laws():
    rng = random.Random(7)
    for _ in range(5):
        a, b = rng.randrange(1, bn.R_ORDER), rng.randrange(1, bn.R_ORDER)
        P, Q2 = bn.g1_mul(bn.BASE_G1, a), bn.g2_mul(bn.BASE_G2, a)
        assert bn.g1_mul(P, b) == bn.g1_mul(bn.BASE_G1, a * b % bn.R_ORDER)  # fr.rs:240-246
        assert bn.g2_mul(Q2, b) == bn.g2_mul(bn.BASE_G2, a * b % bn.R_ORDER)
        assert bn.g1_add(P, bn.g1_mul(bn.BASE_G1, b)) == bn.g1_mul(bn.BASE_G1, (a + b) % bn.R_ORDER)
        assert bn.g1_sub(P, P) is None and bn.g1_add(P, None) == P


def test_pairing_bilinear():
    e1 = bn.pairing(bn.G1_GEN, bn.G2_GEN)
    assert e1 != bn.Fq12.one() and e1 ** bn.R_ORDER == bn.Fq12.one()
    assert bn.pairing(bn.g1_mul(bn.G1_GEN, 5), bn.G2_GEN) == bn.pairing(bn.G1_GEN, bn.g2_mul(bn.G2_GEN, 5)) == e1 ** 5
    assert bn.pairing(None, bn.G2_GEN) == bn.Fq12.one()


def _rand_fr(rng):
    return rng.randrange(1, FR.p)


def test_simple_circuit_test():  # src/lib.rs:156-190
    B = g.BN254Backend()
    rng = random.Random(11)
    qap = g.qap_from_root_rep(FR, circuit.try_parse(FR, SIMPLE))
    weights = circuit.weights(FR, SIMPLE, [3, 2, 4])
    sigma = g.setup(B, qap, tuple(_rand_fr(rng) for _ in range(5)))
    proof = g.prove(B, qap, sigma, weights, _rand_fr(rng), _rand_fr(rng))
    assert g.verify(B, sigma, [2, 34], proof)
    assert not g.verify(B, sigma, [2, 25], proof)


def test_bn_encrypt_quad_test():  # src/groth16/fr.rs:273-302
    B = g.BN254Backend()
    rng = random.Random(12)
    qap = g.qap_from_root_rep(FR, circuit.try_parse(FR, QUAD))
    x, a, b, c = (_rand_fr(rng) for _ in range(4))
    share = (a * x * x + b * x + c) % FR.p
    weights = [1, x, share, a * x % FR.p, a, x * (a * x + b) % FR.p, b, c]
    assert weights == circuit.weights(FR, QUAD, [x, a, b, c])
    sigma = g.setup(B, qap, tuple(_rand_fr(rng) for _ in range(5)))
    proof = g.prove(B, qap, sigma, weights, _rand_fr(rng), _rand_fr(rng))
    assert g.verify(B, sigma, [x, share], proof)


def test_horner_family_matches_parser_and_verifies_on_omega_domain():
    """The synthetic family == what the reference parser yields for the same program
    (deg_15.zk structure, fr.rs:361-416), and it proves/verifies on the omega domain."""
    n = 8
    text = synthetic.horner_program_text(n)
    rep_parser = circuit.try_parse(FR, text)
    rep_family = synthetic.horner_rep(FR, n, [FR.from_usize(k) for k in range(1, n + 1)])
    assert rep_parser == rep_family
    rng = random.Random(13)
    x, cs = _rand_fr(rng), [_rand_fr(rng) for _ in range(n)]
    assert circuit.weights(FR, text, [x] + cs) == synthetic.horner_witness(FR, n, x, cs)

    w = synthetic.omega(3)
    rep = synthetic.horner_rep(FR, n, [pow(w, k, FR.p) for k in range(n)])
    qap = g.qap_from_root_rep(FR, rep)
    assert qap.t == [FR.p - 1] + [0] * (n - 1) + [1]  # x^n - 1
    B = g.BN254Backend()
    weights = synthetic.horner_witness(FR, n, x, cs)
    sigma = g.setup(B, qap, tuple(_rand_fr(rng) for _ in range(5)))
    proof = g.prove(B, qap, sigma, weights, _rand_fr(rng), _rand_fr(rng))
    assert g.verify(B, sigma, weights[1:3], proof)
