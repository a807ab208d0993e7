"""Published alt_bn128 vectors (tests/golden/alt_bn128_published.json: EIP-196 / EIP-197 precompile tests, libff / arkworks
constants) against Oracle A (Python), Oracle B (C) and -- under -m gpu -- the device.  The reference holds no fixed BN254
value (src/groth16/fr.rs:240-416 only assert verify == true), so this is what pins the `bn` layer outside this repository."""

import importlib
import json
import os

import pytest

from oracle import bn254 as bn, oracle_b as ob

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
V = json.load(open(os.path.join(ROOT, "tests", "golden", "alt_bn128_published.json")))


def _i(x):
    return int(x, 0)


def _g1(p):
    return (_i(p[0]), _i(p[1]))


def _pairs():
    w = [_i(x) for x in V["eip197_pairing_jeff1"]["words"]]
    out = []
    for o in (0, 6):
        out.append(((w[o], w[o + 1]), ((w[o + 3], w[o + 2]), (w[o + 5], w[o + 4]))))  # imaginary part first in EIP-197
    return out


def test_oracle_a_matches_published_vectors():
    assert bn.G1_GEN == _g1(V["g1_generator"])
    assert bn.g1_add(bn.G1_GEN, bn.G1_GEN) == _g1(V["g1_double_generator"]) == bn.g1_mul(bn.G1_GEN, 2)
    a = V["eip196_add_chfast1"]
    assert bn.g1_add(_g1(a["p1"]), _g1(a["p2"])) == _g1(a["sum"])
    m = V["eip196_mul_chfast1"]
    assert bn.g1_mul(_g1(m["p"]), _i(m["k"])) == _g1(m["product"])
    g2 = V["g2_generator"]
    assert bn.G2_GEN == (tuple(_i(x) for x in g2["x"]), tuple(_i(x) for x in g2["y"]))
    assert bn.g2_is_on_curve(bn.G2_GEN) and bn.g2_mul(bn.G2_GEN, bn.R_ORDER) is None
    assert tuple(bn.G2_B) == tuple(_i(x) for x in V["twist_coeff_b"])
    assert bn.BN_U == _i(V["bn_parameter_u"]) and bn.ATE_LOOP_COUNT == _i(V["ate_loop_count"]) == 6 * bn.BN_U + 2

    def f2pow(a_, e):
        r = bn.F2_ONE
        while e:
            if e & 1:
                r = bn.f2_mul(r, a_)
            a_ = bn.f2_mul(a_, a_)
            e >>= 1
        return r
    assert tuple(f2pow((9, 1), (bn.Q - 1) // 3)) == tuple(_i(x) for x in V["twist_mul_by_q_x"])
    assert tuple(f2pow((9, 1), (bn.Q - 1) // 2)) == tuple(_i(x) for x in V["twist_mul_by_q_y"])


def test_oracle_a_pairing_matches_eip197_vector():
    (p1, q1), (p2, q2) = _pairs()
    assert bn.g1_is_on_curve(p1) and bn.g2_is_on_curve(q1) and bn.g1_is_on_curve(p2) and bn.g2_is_on_curve(q2)
    prod = bn.pairing(p1, q1) * bn.pairing(p2, q2)
    assert (prod == bn.Fq12.one()) is V["eip197_pairing_jeff1"]["product_is_one"]
    assert bn.pairing(p1, q1) != bn.Fq12.one()  # not degenerate


def test_oracle_b_matches_published_vectors():
    ob.build()
    a = V["eip196_add_chfast1"]
    assert ob.g1_add(_g1(a["p1"]), _g1(a["p2"])) == _g1(a["sum"])
    assert ob.g1_add((1, 2), (1, 2)) == _g1(V["g1_double_generator"])
    m = V["eip196_mul_chfast1"]
    assert ob.g1_mul(_g1(m["p"]), _i(m["k"])) == _g1(m["product"])
    (p1, q1), _ = _pairs()
    assert ob.g2_add(q1, q1) == bn.g2_add(q1, q1)  # a published G2 point through both oracles


@pytest.mark.gpu
def test_device_matches_published_vectors():
    zk = importlib.import_module("zksnark-rs_b200")
    zg = importlib.import_module("zksnark-rs_b200.groth16")
    ctx = zk.Context(0)
    try:
        a = V["eip196_add_chfast1"]
        assert zg.points_sum(ctx, 1, [_g1(a["p1"]), _g1(a["p2"])]) == _g1(a["sum"])
        assert zg.points_sum(ctx, 1, [(1, 2), (1, 2)]) == _g1(V["g1_double_generator"])
        m = V["eip196_mul_chfast1"]
        bases = zk.Bases.upload(ctx, 1, [_g1(m["p"]), (1, 2)])           # on-curve validation of raw points included
        assert zk.msm(ctx, bases, [_i(m["k"]), 0]) == _g1(m["product"])
        assert zk.msm(ctx, bases, [0, 2]) == _g1(V["g1_double_generator"])
        (p1, q1), (p2, q2) = _pairs()
        zk.Bases.upload(ctx, 2, [q1, q2, bn.G2_GEN])                     # on the twist AND in the order-r subgroup
        gt = zk.pairing(ctx, [(p1, q1), (p2, q2)])                        # product of the two pairings
        assert gt == [1] + [0] * 11
        assert zk.pairing(ctx, [(p1, q1)]) != [1] + [0] * 11
    finally:
        ctx.close()
