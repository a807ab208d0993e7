"""CPU check of the product's pairing formulas (csrc/pairing.cuh compiled with g++ by tests/hostcheck)
against Oracle A's dense-polynomial Fq12 and py_ecc-style Miller loop (oracle/bn254.py:195-370).
The reduced pairing is a canonical GT element, so the comparison is coefficient by coefficient after the
change of basis  sum_i (c0_i + c1_i u) w^i,  u = w^6 - 9  ->  sum_k f_k w^k."""

import ctypes
import os
import random
import subprocess

import pytest

from oracle import bn254 as bn
from oracle.fields import FR, Q_MODULUS as Q

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
R = FR.p


@pytest.fixture(scope="module")
def hc():
    src = os.path.join(HERE, "hostcheck", "hostcheck.cpp")
    so = os.path.join(HERE, "hostcheck", "_hostcheck.so")
    deps = [src] + [os.path.join(ROOT, "zksnark-rs_b200", "csrc", f) for f in ("ff.cuh", "ec.cuh", "pairing.cuh", "constants.h")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-Wno-unknown-pragmas",
                               "-I", os.path.join(ROOT, "zksnark-rs_b200", "csrc"), "-o", so, src])
    return ctypes.CDLL(so)


def limbs(vals):
    out = (ctypes.c_uint64 * (4 * len(vals)))()
    for i, v in enumerate(vals):
        for j in range(4):
            out[4 * i + j] = (v >> (64 * j)) & 0xFFFFFFFFFFFFFFFF
    return out


def ints(buf, n):
    return [sum(buf[4 * i + j] << (64 * j) for j in range(4)) for i in range(n)]


def g1_vals(P):
    return [0, 0] if P is None else [P[0], P[1]]


def g2_vals(P):
    return [0, 0, 0, 0] if P is None else [P[0][0], P[0][1], P[1][0], P[1][1]]


def tower_to_flat(c):
    """12 residues (c0, c1 of the w^i coefficient, i = 0..5) -> the oracle's dense w-polynomial."""
    f = [0] * 12
    for i in range(6):
        c0, c1 = c[2 * i], c[2 * i + 1]
        f[i] = (f[i] + c0 - 9 * c1) % Q
        f[i + 6] = (f[i + 6] + c1) % Q
    return bn.Fq12(f)


def flat_to_tower(f):
    c = []
    for i in range(6):
        c1 = f.c[i + 6]
        c += [(f.c[i] + 9 * c1) % Q, c1]
    return c


def hc_pairing(hc, stage, pairs):
    g1 = limbs([v for P, _ in pairs for v in g1_vals(P)])
    g2 = limbs([v for _, T in pairs for v in g2_vals(T)])
    out = (ctypes.c_uint64 * 48)()
    hc.hc_pairing(stage, len(pairs), g1, g2, out)
    return tower_to_flat(ints(out, 12))


def test_fq12_ops_match_oracle(hc):
    rng = random.Random(11)
    for _ in range(4):
        a = bn.Fq12([rng.randrange(Q) for _ in range(12)])
        b = bn.Fq12([rng.randrange(Q) for _ in range(12)])
        la, lb = limbs(flat_to_tower(a)), limbs(flat_to_tower(b))
        line = bn.Fq12([0] * 12)
        tb = flat_to_tower(b)
        sparse = tower_to_flat(tb[0:4] + [0, 0] + tb[6:8] + [0, 0, 0, 0])  # keep the w^0, w^1, w^3 coefficients
        want = [a * b, a * a, a.inv(), a ** (Q * Q), a ** (Q ** 6), a ** Q, None, a * sparse]
        for op, w in enumerate(want):
            if w is None:
                continue
            out = (ctypes.c_uint64 * 48)()
            hc.hc_fq12(op, la, lb, out)
            assert tower_to_flat(ints(out, 12)) == w, op
        # cyclotomic subgroup (after the easy part): Granger-Scott squaring and the u-power built on it
        c = (a ** (Q ** 6)) * a.inv()
        c = (c ** (Q * Q)) * c
        out = (ctypes.c_uint64 * 48)()
        hc.hc_fq12(9, la, lb, out)
        assert tower_to_flat(ints(out, 12)) == c
        lc = limbs(flat_to_tower(c))
        for op, w in ((8, c * c), (6, c ** bn.BN_U), (1, c * c)):
            hc.hc_fq12(op, lc, lc, out)
            assert tower_to_flat(ints(out, 12)) == w, op


def test_pairing_matches_oracle(hc):
    rng = random.Random(12)
    a, b = rng.randrange(1, R), rng.randrange(1, R)
    P, T = bn.g1_mul(bn.BASE_G1, a), bn.g2_mul(bn.BASE_G2, b)
    got = hc_pairing(hc, 1, [(P, T)])
    assert got == bn.pairing(P, T)
    # Miller values differ from the oracle's only by subfield factors: equal after the final exponentiation
    assert bn.final_exponentiation(hc_pairing(hc, 0, [(P, T)])) == got
    # the plain versions (affine steps with an inversion each; 761-bit hard exponent) define the fast ones
    assert hc_pairing(hc, 3, [(P, T)]) == got
    assert bn.final_exponentiation(hc_pairing(hc, 2, [(P, T)])) == got
    # bilinearity and the product form used by verify: e(aG, bH) e(-abG, H) == 1
    nP = bn.g1_neg(bn.g1_mul(bn.BASE_G1, a * b % R))
    assert hc_pairing(hc, 1, [(P, T), (nP, bn.BASE_G2)]) == bn.Fq12.one()
    assert hc_pairing(hc, 1, [(P, T), (bn.g1_neg(P), T)]) == bn.Fq12.one()
    # identity in either slot gives 1
    assert hc_pairing(hc, 1, [(None, T)]) == bn.Fq12.one()
    assert hc_pairing(hc, 1, [(P, None)]) == bn.Fq12.one()
