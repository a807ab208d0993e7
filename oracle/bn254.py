"""Oracle A -- BN254 (alt_bn128) group layer (test infrastructure only).

Restates, from the published curve definition, the part of crate ``bn`` 0.4.3
(zcash-hackworks/bn, pinned by /root/reference/Cargo.toml:15; NOT vendored in the
reference tree) that the reference reaches through src/groth16/fr.rs:

* fr.rs:106-113  ``G1::one()``, ``G2::one()``, ``Group * Fr``   -> g1_mul / g2_mul
* fr.rs:175-223  point + / - / Sum (fold from ``G::zero()``)     -> g1_add / g2_add ...
* fr.rs:120-122  ``bn::pairing``                                 -> pairing()
* fr.rs:225-231  GT "+" is Fq12 multiplication                   -> Fq12.__mul__

Parity is defined on canonical residues and AFFINE coordinates, never on the
crate's internal Montgomery/Jacobian representation, so plain affine formulas
are used here: points are ``None`` (identity) or ``(x, y)``.  G1 coordinates are
ints mod q; G2 coordinates are Fq2 elements ``(c0, c1)`` = c0 + c1*u, u^2 = -1.

"parity unpinned" for these values (see oracle/__init__.py): the constants below
were checked numerically (primality, on-curve, r*P = O) by tests/test_oracle_bn254.py.
"""

from __future__ import annotations

from .fields import Q_MODULUS as Q, FR

R_ORDER = FR.p
BN_U = 4965661367192848881  # BN parameter; q = 36u^4+36u^3+24u^2+6u+1

# ------------------------------------------------------------------ Fq helpers
def fq_inv(a: int) -> int:
    return pow(a, -1, Q)


# ------------------------------------------------------------------ Fq2 = Fq[u]/(u^2+1)
def f2_add(a, b):
    return ((a[0] + b[0]) % Q, (a[1] + b[1]) % Q)


def f2_sub(a, b):
    return ((a[0] - b[0]) % Q, (a[1] - b[1]) % Q)


def f2_neg(a):
    return ((-a[0]) % Q, (-a[1]) % Q)


def f2_mul(a, b):
    return ((a[0] * b[0] - a[1] * b[1]) % Q, (a[0] * b[1] + a[1] * b[0]) % Q)


def f2_scalar(a, k: int):
    return ((a[0] * k) % Q, (a[1] * k) % Q)


def f2_inv(a):
    n = fq_inv((a[0] * a[0] + a[1] * a[1]) % Q)
    return ((a[0] * n) % Q, (-a[1] * n) % Q)


F2_ZERO = (0, 0)
F2_ONE = (1, 0)

# ------------------------------------------------------------------ curve constants
G1_B = 3
G1_GEN = (1, 2)  # bn: G1::one()
# twist: y^2 = x^3 + 3/(9+u)
G2_B = f2_mul((3, 0), f2_inv((9, 1)))
G2_GEN = (  # bn: G2::one()
    (
        10857046999023057135944570762232829481370756359578518086990519993285655852781,
        11559732032986387107991004021392285783925812861821192530917403151452391805634,
    ),
    (
        8495653923123431417604973247489272438418190587263600148770280649306958101930,
        4082367875863433681332203403145435568316851327593401208105741076214120093531,
    ),
)


# ------------------------------------------------------------------ G1 (affine over Fq)
def g1_is_on_curve(P) -> bool:
    if P is None:
        return True
    x, y = P
    return (y * y - x * x * x - G1_B) % Q == 0


def g1_neg(P):
    return None if P is None else (P[0], (-P[1]) % Q)


def g1_add(P, R):
    if P is None:
        return R
    if R is None:
        return P
    x1, y1 = P
    x2, y2 = R
    if x1 == x2:
        if (y1 + y2) % Q == 0:
            return None
        lam = (3 * x1 * x1) * fq_inv(2 * y1) % Q
    else:
        lam = (y2 - y1) * fq_inv((x2 - x1) % Q) % Q
    x3 = (lam * lam - x1 - x2) % Q
    return (x3, (lam * (x1 - x3) - y1) % Q)


def g1_sub(P, R):
    return g1_add(P, g1_neg(R))


def g1_mul(P, k: int):
    """``Group * Fr`` (double-and-add); k is reduced mod r like an Fr value."""
    k %= R_ORDER
    acc = None
    while k:
        if k & 1:
            acc = g1_add(acc, P)
        P = g1_add(P, P)
        k >>= 1
    return acc


def g1_sum(points):
    """``Sum for G1Local`` fr.rs:191-198: fold from G1::zero()."""
    acc = None
    for p in points:
        acc = g1_add(acc, p)
    return acc


# ------------------------------------------------------------------ G2 (affine over Fq2)
def g2_is_on_curve(P) -> bool:
    if P is None:
        return True
    x, y = P
    return f2_sub(f2_mul(y, y), f2_add(f2_mul(f2_mul(x, x), x), G2_B)) == F2_ZERO


def g2_neg(P):
    return None if P is None else (P[0], f2_neg(P[1]))


def g2_add(P, R):
    if P is None:
        return R
    if R is None:
        return P
    x1, y1 = P
    x2, y2 = R
    if x1 == x2:
        if f2_add(y1, y2) == F2_ZERO:
            return None
        lam = f2_mul(f2_scalar(f2_mul(x1, x1), 3), f2_inv(f2_scalar(y1, 2)))
    else:
        lam = f2_mul(f2_sub(y2, y1), f2_inv(f2_sub(x2, x1)))
    x3 = f2_sub(f2_sub(f2_mul(lam, lam), x1), x2)
    return (x3, f2_sub(f2_mul(lam, f2_sub(x1, x3)), y1))


def g2_mul(P, k: int):
    k %= R_ORDER
    acc = None
    while k:
        if k & 1:
            acc = g2_add(acc, P)
        P = g2_add(P, P)
        k >>= 1
    return acc


def g2_sum(points):
    """``Sum for G2Local`` fr.rs:216-223."""
    acc = None
    for p in points:
        acc = g2_add(acc, p)
    return acc


# ------------------------------------------------------------------ reference base points
# fr.rs:106-113: encrypt_g1(s) = (G1::one()*69)*s ; encrypt_g2(s) = (G2::one()*96)*s
BASE_G1 = g1_mul(G1_GEN, 69)
BASE_G2 = g2_mul(G2_GEN, 96)


def msm_g1(scalars, points):
    """The reference's G1 'MSM' pattern (groth16/mod.rs:255-260): zip, scalar-mul, Sum."""
    return g1_sum(g1_mul(P, s) for s, P in zip(scalars, points))


def msm_g2(scalars, points):
    return g2_sum(g2_mul(P, s) for s, P in zip(scalars, points))


# ------------------------------------------------------------------ Fq12 and the pairing
# Fq12 = Fq[w]/(w^12 - 18 w^6 + 82); u = w^6 - 9 (so that (9+u) = w^6 is the sextic
# non-residue).  Dense degree-<12 polynomials over Fq; slow and simple on purpose.
_F12_MOD = (82, 0, 0, 0, 0, 0, -18 % Q, 0, 0, 0, 0, 0)  # low-order coeffs of the monic modulus


class Fq12:
    __slots__ = ("c",)

    def __init__(self, c):
        self.c = tuple(x % Q for x in c)

    @staticmethod
    def one():
        return Fq12((1,) + (0,) * 11)

    @staticmethod
    def zero():
        return Fq12((0,) * 12)

    def __eq__(self, o):
        return self.c == o.c

    def __add__(self, o):
        return Fq12(tuple(a + b for a, b in zip(self.c, o.c)))

    def __sub__(self, o):
        return Fq12(tuple(a - b for a, b in zip(self.c, o.c)))

    def __neg__(self):
        return Fq12(tuple(-a for a in self.c))

    def scale(self, k: int):
        return Fq12(tuple(a * k for a in self.c))

    def __mul__(self, o):
        t = [0] * 23
        for i, a in enumerate(self.c):
            if a:
                for j, b in enumerate(o.c):
                    t[i + j] += a * b
        for k in range(22, 11, -1):  # w^12 = 18 w^6 - 82
            top = t[k] % Q
            if top:
                t[k - 6] += 18 * top
                t[k - 12] -= 82 * top
            t[k] = 0
        return Fq12(t[:12])

    def __pow__(self, e: int):
        acc, base = Fq12.one(), self
        while e:
            if e & 1:
                acc = acc * base
            base = base * base
            e >>= 1
        return acc

    def inv(self):
        # polynomial extended Euclid over Fq against the modulus
        def deg(p):
            d = len(p) - 1
            while d and p[d] % Q == 0:
                d -= 1
            return d

        lm, hm = [1] + [0] * 12, [0] * 13
        low, high = list(self.c) + [0], list(_F12_MOD) + [1]
        while deg(low):
            # r = high / low (polynomial rounded division)
            dl, dh = deg(low), deg(high)
            temp = list(high)
            o = [0] * 13
            for i in range(dh - dl, -1, -1):
                o[i] = temp[dl + i] * fq_inv(low[dl]) % Q
                for c in range(dl + 1):
                    temp[c + i] = (temp[c + i] - o[i] * low[c]) % Q
            r = o
            nm, new = list(hm), list(high)
            for i in range(13):
                for j in range(13 - i):
                    nm[i + j] = (nm[i + j] - lm[i] * r[j]) % Q
                    new[i + j] = (new[i + j] - low[i] * r[j]) % Q
            lm, low, hm, high = nm, new, lm, low
        k = fq_inv(low[0])
        return Fq12(tuple(x * k for x in lm[:12]))


def _f12_from_fq(a: int) -> Fq12:
    return Fq12((a,) + (0,) * 11)


def _twist(P):
    """Map a G2 point (over Fq2) onto the curve y^2 = x^3 + 3 over Fq12."""
    (x0, x1), (y0, y1) = P
    # c0 + c1*u with u = w^6 - 9  ->  (c0 - 9 c1) + c1 w^6
    nx = Fq12(((x0 - 9 * x1),) + (0,) * 5 + (x1,) + (0,) * 5)
    ny = Fq12(((y0 - 9 * y1),) + (0,) * 5 + (y1,) + (0,) * 5)
    w = Fq12((0, 1) + (0,) * 10)
    return (nx * (w * w), ny * (w * w * w))


def _f12_double(P):
    x, y = P
    lam = (x * x).scale(3) * (y.scale(2)).inv()
    nx = lam * lam - x - x
    return (nx, lam * (x - nx) - y)


def _f12_add(P, R):
    x1, y1 = P
    x2, y2 = R
    if x1 == x2:
        if y1 == y2:
            return _f12_double(P)
        raise ValueError("unexpected P + (-P) inside the Miller loop")
    lam = (y2 - y1) * (x2 - x1).inv()
    nx = lam * lam - x1 - x2
    return (nx, lam * (x1 - nx) - y1)


def _linefunc(P1, P2, T):
    """Line through P1,P2 (tangent if equal) evaluated at T; all over Fq12."""
    x1, y1 = P1
    x2, y2 = P2
    xt, yt = T
    if not (x1 == x2):
        m = (y2 - y1) * (x2 - x1).inv()
        return m * (xt - x1) - (yt - y1)
    if y1 == y2:
        m = (x1 * x1).scale(3) * (y1.scale(2)).inv()
        return m * (xt - x1) - (yt - y1)
    return xt - x1


ATE_LOOP_COUNT = 6 * BN_U + 2  # 29793968203157093288
_FINAL_EXP = (Q ** 12 - 1) // R_ORDER


def miller_loop(Q2, P1) -> Fq12:
    """Optimal-ate Miller loop f_{6u+2,Q}(P) * l_{..}(P) (no final exponentiation).

    Q2 in G2 (affine over Fq2), P1 in G1 (affine).  Identity in either slot -> 1.
    """
    if Q2 is None or P1 is None:
        return Fq12.one()
    Qt = _twist(Q2)
    Pt = (_f12_from_fq(P1[0]), _f12_from_fq(P1[1]))
    Rp = Qt
    f = Fq12.one()
    for i in range(ATE_LOOP_COUNT.bit_length() - 2, -1, -1):
        f = f * f * _linefunc(Rp, Rp, Pt)
        Rp = _f12_double(Rp)
        if (ATE_LOOP_COUNT >> i) & 1:
            f = f * _linefunc(Rp, Qt, Pt)
            Rp = _f12_add(Rp, Qt)
    Q1 = (Qt[0] ** Q, Qt[1] ** Q)
    nQ2 = (Q1[0] ** Q, -(Q1[1] ** Q))
    f = f * _linefunc(Rp, Q1, Pt)
    Rp = _f12_add(Rp, Q1)
    f = f * _linefunc(Rp, nQ2, Pt)
    return f


def final_exponentiation(f: Fq12) -> Fq12:
    return f ** _FINAL_EXP


def pairing(P1, Q2) -> Fq12:
    """``bn::pairing(g1, g2)`` (fr.rs:120-122) up to the choice of GT generator.

    Any non-degenerate bilinear pairing gives the same truth value for the only
    thing the reference does with GT: the equality test in ``verify``
    (groth16/mod.rs:316-319).
    """
    return final_exponentiation(miller_loop(Q2, P1))
