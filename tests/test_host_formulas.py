"""CPU check of the product's shared host/device field and curve formulas (ff.cuh, ec.cuh compiled
with g++ by tests/hostcheck) against Oracle A.  Covers every special case of the group law the
reference's `bn` add handles (identity operands, P+P, P+(-P)); SURVEY.md section 7 "hard parts"."""

import ctypes
import os
import random
import subprocess

import pytest

from oracle import bn254 as bn
from oracle.fields import FR, Q_MODULUS

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def hc():
    src = os.path.join(HERE, "hostcheck", "hostcheck.cpp")
    so = os.path.join(HERE, "hostcheck", "_hostcheck.so")
    deps = [src] + [os.path.join(ROOT, "zksnark-rs_b200", "csrc", f) for f in ("ff.cuh", "ec.cuh", "pairing.cuh", "constants.h")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-Wno-unknown-pragmas",
                               "-I", os.path.join(ROOT, "zksnark-rs_b200", "csrc"), "-o", so, src])
    return ctypes.CDLL(so)


def pack(*vals):
    out = (ctypes.c_uint64 * (4 * len(vals)))()
    for i, v in enumerate(vals):
        for j in range(4):
            out[4 * i + j] = (v >> (64 * j)) & 0xFFFFFFFFFFFFFFFF
    return out


def unpack(buf, n):
    return [sum(buf[4 * i + j] << (64 * j) for j in range(4)) for i in range(n)]


def g1_pack(P):
    return pack(0, 0) if P is None else pack(P[0], P[1])


def g1_unpack(buf):
    x, y = unpack(buf, 2)
    return None if (x, y) == (0, 0) else (x, y)


def g2_pack(P):
    return pack(0, 0, 0, 0) if P is None else pack(P[0][0], P[0][1], P[1][0], P[1][1])


def g2_unpack(buf):
    a, b, c, d = unpack(buf, 4)
    return None if (a, b, c, d) == (0, 0, 0, 0) else ((a, b), (c, d))


def test_field_ops(hc):
    rng = random.Random(21)
    for field, p in ((0, FR.p), (1, Q_MODULUS)):
        cases = [(0, 0), (1, 1), (p - 1, p - 1), (p - 1, 1), (0, p - 1)]
        cases += [(rng.randrange(p), rng.randrange(p)) for _ in range(300)]
        for a, b in cases:
            out = pack(0)
            for op, ref in ((0, a * b % p), (1, (a + b) % p), (2, (a - b) % p)):
                hc.hc_field(field, op, pack(a), pack(b), out)
                assert unpack(out, 1)[0] == ref
        for a in [1, 2, p - 1] + [rng.randrange(1, p) for _ in range(5)]:
            out = pack(0)
            hc.hc_field(field, 3, pack(a), pack(0), out)
            assert unpack(out, 1)[0] == pow(a, -1, p)


def test_fq2_ops(hc):
    rng = random.Random(22)
    q = Q_MODULUS
    for _ in range(200):
        a = (rng.randrange(q), rng.randrange(q))
        b = (rng.randrange(q), rng.randrange(q))
        out = pack(0, 0)
        for op, ref in ((0, bn.f2_mul(a, b)), (1, bn.f2_add(a, b)), (2, bn.f2_sub(a, b)), (3, bn.f2_inv(a)),
                        (4, bn.f2_mul(a, a))):
            hc.hc_fq2(op, pack(*a), pack(*b), out)
            assert tuple(unpack(out, 2)) == ref


def _g1_cases(rng):
    P = bn.g1_mul(bn.BASE_G1, rng.randrange(1, bn.R_ORDER))
    Q = bn.g1_mul(bn.BASE_G1, rng.randrange(1, bn.R_ORDER))
    return [(P, Q), (P, P), (P, bn.g1_neg(P)), (None, Q), (P, None), (None, None)]


def test_g1_group_law(hc):
    rng = random.Random(23)
    for _ in range(3):
        for P, Q in _g1_cases(rng):
            o1, o2 = pack(0, 0), pack(0, 0)
            hc.hc_g1_add(g1_pack(P), g1_pack(Q), o1, o2)
            ref = bn.g1_add(P, Q)
            assert g1_unpack(o1) == ref and g1_unpack(o2) == ref
    P = bn.BASE_G1
    for k in [0, 1, 2, 69, bn.R_ORDER - 1, rng.randrange(bn.R_ORDER)]:
        out = pack(0, 0)
        hc.hc_g1_mul(g1_pack(P), pack(k), out)
        assert g1_unpack(out) == bn.g1_mul(P, k)


def test_g2_group_law(hc):
    rng = random.Random(24)
    P = bn.g2_mul(bn.BASE_G2, rng.randrange(1, bn.R_ORDER))
    Q = bn.g2_mul(bn.BASE_G2, rng.randrange(1, bn.R_ORDER))
    for A, B in [(P, Q), (P, P), (P, bn.g2_neg(P)), (None, Q), (P, None), (None, None)]:
        o1, o2 = pack(0, 0, 0, 0), pack(0, 0, 0, 0)
        hc.hc_g2_add(g2_pack(A), g2_pack(B), o1, o2)
        ref = bn.g2_add(A, B)
        assert g2_unpack(o1) == ref and g2_unpack(o2) == ref
    for k in [0, 1, 96, bn.R_ORDER - 1, rng.randrange(bn.R_ORDER)]:
        out = pack(0, 0, 0, 0)
        hc.hc_g2_mul(g2_pack(bn.G2_GEN), pack(k), out)
        assert g2_unpack(out) == bn.g2_mul(bn.G2_GEN, k)


def test_fast_inverse_matches_fermat_and_bigint(hc):
    """The safegcd inverse (ff.cuh: inverse) against the Fermat exponentiation and Python's pow on random and edge
    inputs, for both moduli; inverse(0) = 0."""
    rng = random.Random(99)
    for field, p in ((0, FR.p), (1, Q_MODULUS)):
        vals = [0, 1, 2, p - 1, p - 2, (p + 1) // 2, 1 << 253, (1 << 253) - 1, 3 << 200, p >> 1]
        vals += [1 << k for k in range(0, 254, 7)] + [p - (1 << k) for k in range(0, 250, 11)]
        vals += [rng.randrange(p) for _ in range(3000)] + [rng.randrange(1 << 64) for _ in range(100)]
        n = len(vals)
        out, out2 = (ctypes.c_uint64 * (4 * n))(), (ctypes.c_uint64 * (4 * n))()
        hc.hc_inverse_many(field, n, pack(*vals), out, out2)
        got, fermat = unpack(out, n), unpack(out2, n)
        want = [pow(v, -1, p) if v else 0 for v in vals]
        assert fermat == want
        assert got == want
