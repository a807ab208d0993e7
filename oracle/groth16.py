"""Oracle A -- Groth16 protocol layer (test infrastructure only, see oracle/__init__.py).

Literal restatement of src/groth16/mod.rs: ``QAP`` (:60-102), ``setup`` (:134-197),
``prove`` (:213-296), ``verify`` (:299-320), generic over a back-end object that
plays the role of the ``T: EllipticEncryptable + Random + Field`` type parameter.

The reference draws its randomness from ``thread_rng`` inside ``setup``/``prove``
(:139-145, :231); here the caller passes it in (``toxic`` / ``r, s``) -- that is
the only seam that makes bit-exact comparison possible.
"""

from __future__ import annotations

from dataclasses import dataclass

from . import bn254 as bn
from .fields import FR, Z251, Field
from .poly import (degree, evaluate, poly_div, poly_from_points, poly_mul, poly_scale,
                   poly_sub, poly_sum, powers, root_poly)


# ------------------------------------------------------------------ back-ends
class Z251Backend:
    """The reference's fake curve over F_251, groth16/mod.rs:329-374 (G1=G2=GT=Z251)."""

    F: Field = Z251

    def encrypt_g1(self, s):
        return Z251.mul(s, Z251.from_usize(69))

    encrypt_g2 = encrypt_g1

    def exp_g1(self, s, g):
        return Z251.mul(s, g)

    exp_g2 = exp_g1

    def g1_add(self, a, b):
        return Z251.add(a, b)

    g2_add = g1_add

    def g1_sub(self, a, b):
        return Z251.sub(a, b)

    def g1_sum(self, it):
        acc = Z251.from_usize(0)
        for x in it:
            acc = Z251.add(acc, x)
        return acc

    g2_sum = g1_sum

    def pairing(self, g1, g2):
        return Z251.mul(g1, g2)

    def gt_add(self, a, b):
        return Z251.add(a, b)

    def gt_eq(self, a, b):
        return a == b


class BN254Backend:
    """``impl EllipticEncryptable for FrLocal``, groth16/fr.rs:101-123 + Sum impls :191-231."""

    F: Field = FR

    def encrypt_g1(self, s):
        return bn.g1_mul(bn.BASE_G1, s)  # (G1::one()*69)*s

    def encrypt_g2(self, s):
        return bn.g2_mul(bn.BASE_G2, s)  # (G2::one()*96)*s

    def exp_g1(self, s, g):
        return bn.g1_mul(g, s)

    def exp_g2(self, s, g):
        return bn.g2_mul(g, s)

    g1_add = staticmethod(bn.g1_add)
    g2_add = staticmethod(bn.g2_add)
    g1_sub = staticmethod(bn.g1_sub)
    g1_sum = staticmethod(bn.g1_sum)
    g2_sum = staticmethod(bn.g2_sum)

    def pairing(self, g1, g2):
        return bn.pairing(g1, g2)

    def gt_add(self, a, b):
        return a * b  # fr.rs:225-231

    def gt_eq(self, a, b):
        return a == b


# ------------------------------------------------------------------ data types
@dataclass
class QAP:
    """groth16/mod.rs:60-67 -- dense coefficient polynomials."""

    u: list
    v: list
    w: list
    t: list
    input: int
    degree: int


@dataclass
class SigmaG1:
    """groth16/mod.rs:105-113."""

    alpha: object
    beta: object
    delta: object
    xi: list
    sum_gamma: list
    sum_delta: list
    xi_t: list


@dataclass
class SigmaG2:
    """groth16/mod.rs:116-121."""

    beta: object
    gamma: object
    delta: object
    xi: list


@dataclass
class Proof:
    """groth16/mod.rs:124-128."""

    a: object
    b: object
    c: object


def qap_from_root_rep(F: Field, rep) -> QAP:
    """``From<RootRepresentation> for QAP``, groth16/mod.rs:69-102 == fr.rs:140-173."""
    u = [poly_from_points(F, rep.roots, pts) for pts in rep.u]
    v = [poly_from_points(F, rep.roots, pts) for pts in rep.v]
    w = [poly_from_points(F, rep.roots, pts) for pts in rep.w]
    assert len(u) == len(v) == len(w)
    t = root_poly(F, rep.roots)
    return QAP(u=u, v=v, w=w, t=t, input=rep.input, degree=degree(F, t))


# ------------------------------------------------------------------ protocol
def setup(B, qap: QAP, toxic):
    """groth16/mod.rs:134-197.  ``toxic`` = (alpha, beta, gamma, delta, x), all non-zero."""
    F = B.F
    alpha, beta, gamma, delta, x = toxic
    xi = powers(F, x, qap.degree)

    def lin(i):
        return F.add(F.add(F.mul(beta, evaluate(F, qap.u[i], x)),
                           F.mul(alpha, evaluate(F, qap.v[i], x))),
                     evaluate(F, qap.w[i], x))

    rows = min(len(qap.u), len(qap.v), len(qap.w))  # zip
    sum_gamma = [B.encrypt_g1(F.div(lin(i), gamma)) for i in range(min(rows, qap.input + 1))]
    sum_delta = [B.encrypt_g1(F.div(lin(i), delta)) for i in range(qap.input + 1, rows)]
    tx = evaluate(F, qap.t, x)
    xi_t = [B.encrypt_g1(F.div(F.mul(p, tx), delta)) for p in xi[: max(len(xi) - 1, 0)]]
    s1 = SigmaG1(alpha=B.encrypt_g1(alpha), beta=B.encrypt_g1(beta), delta=B.encrypt_g1(delta),
                 xi=[B.encrypt_g1(p) for p in xi], sum_gamma=sum_gamma, sum_delta=sum_delta,
                 xi_t=xi_t)
    s2 = SigmaG2(beta=B.encrypt_g2(beta), gamma=B.encrypt_g2(gamma), delta=B.encrypt_g2(delta),
                 xi=[B.encrypt_g2(p) for p in xi])
    return s1, s2


def weighted_sums(F: Field, qap: QAP, weights):
    """groth16/mod.rs:233-253."""
    u_sum = poly_sum(F, (poly_scale(F, p, a) for p, a in zip(qap.u, weights)))
    v_sum = poly_sum(F, (poly_scale(F, p, a) for p, a in zip(qap.v, weights)))
    w_sum = poly_sum(F, (poly_scale(F, p, a) for p, a in zip(qap.w, weights)))
    return u_sum, v_sum, w_sum


def quotient_h(F: Field, qap: QAP, u_sum, v_sum, w_sum):
    """groth16/mod.rs:277: h = (u_sum * v_sum - w_sum) / t."""
    return poly_div(F, poly_sub(F, poly_mul(F, u_sum, v_sum), w_sum), qap.t)


def prove(B, qap: QAP, sigma, weights, r, s) -> Proof:
    """groth16/mod.rs:213-296 with the two random scalars injected."""
    F = B.F
    s1, s2 = sigma
    u_sum, v_sum, w_sum = weighted_sums(F, qap, weights)
    a_g1 = B.g1_sum(B.exp_g1(a, x) for a, x in zip(u_sum, s1.xi))
    b_g1 = B.g1_sum(B.exp_g1(a, x) for a, x in zip(v_sum, s1.xi))
    b_g2 = B.g2_sum(B.exp_g2(a, x) for a, x in zip(v_sum, s2.xi))
    a = B.g1_add(B.g1_add(a_g1, s1.alpha), B.exp_g1(r, s1.delta))
    b = B.g2_add(B.g2_add(b_g2, s2.beta), B.exp_g2(s, s2.delta))
    h = quotient_h(F, qap, u_sum, v_sum, w_sum)
    c = B.g1_sum(B.exp_g1(c_, x) for c_, x in zip(h, s1.xi_t))
    c = B.g1_add(c, B.g1_sum(B.exp_g1(c_, x)
                             for c_, x in zip(weights[qap.input + 1:], s1.sum_delta)))
    c = B.g1_add(c, B.exp_g1(s, a))
    c = B.g1_add(c, B.exp_g1(r, B.g1_add(B.g1_add(s1.beta, b_g1), B.exp_g1(s, s1.delta))))
    c = B.g1_sub(c, B.exp_g1(F.mul(r, s), s1.delta))
    return Proof(a=a, b=b, c=c)


def verify(B, sigma, inputs, proof: Proof) -> bool:
    """groth16/mod.rs:299-320."""
    F = B.F
    s1, s2 = sigma
    sum_term = B.g1_sum(B.exp_g1(a, x) for x, a in zip(s1.sum_gamma, [F.one()] + list(inputs)))
    lhs = B.gt_add(B.gt_add(B.pairing(s1.alpha, s2.beta), B.pairing(sum_term, s2.gamma)),
                   B.pairing(proof.c, s2.delta))
    return B.gt_eq(lhs, B.pairing(proof.a, proof.b))
