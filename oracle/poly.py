"""Oracle A -- polynomial layer (test infrastructure only, see oracle/__init__.py).

Literal restatement of ``CoefficientPoly`` (src/groth16/coefficient_poly.rs:1-200)
and the ``Polynomial`` helpers of src/field/mod.rs.  A polynomial is a Python
list of field elements, little-endian in degree (index = exponent), exactly as
``CoefficientPoly.coeffs``.  Every function takes the ``Field`` first.
"""

from __future__ import annotations

from .fields import Field, FieldPanic


# ---- Polynomial trait helpers, src/field/mod.rs:256-355 -----------------------

def degree(F: Field, p: list) -> int:
    """field/mod.rs:291-297: count after skipping leading (high) zeros; empty/zero -> 0."""
    n = len(p)
    while n > 0 and F.eq(p[n - 1], F.zero()):
        n -= 1
    return 0 if n == 0 else n - 1


def remove_leading_zeros(F: Field, p: list) -> list:
    """field/mod.rs:344-355 (returns a new list; a zero poly becomes [])."""
    n = len(p)
    while n > 0 and F.eq(p[n - 1], F.zero()):
        n -= 1
    return list(p[:n])


def evaluate(F: Field, p: list, x: int) -> int:
    """Horner, field/mod.rs:338-343."""
    acc = F.zero()
    for c in reversed(p):
        acc = F.add(F.mul(acc, x), c)
    return acc


def powers(F: Field, x: int, count: int) -> list:
    """field/mod.rs:493-504: x^0, x^1, ... (first ``count`` of the infinite iterator)."""
    out = []
    cur = F.one()
    for _ in range(count):
        out.append(cur)
        cur = F.mul(cur, x)
    return out


def dft(F: Field, seq: list, root: int) -> list:
    """Naive O(n^2) DFT, field/mod.rs:508-520.  out[i] = sum_j seq[j]*root^(i*j)."""
    n = len(seq)
    out = []
    for ri in powers(F, root, n):
        acc = F.zero()
        for a, r in zip(seq, powers(F, ri, n)):
            acc = F.add(acc, F.mul(a, r))
        out.append(acc)
    return out


def idft(F: Field, seq: list, root: int) -> list:
    """field/mod.rs:524-537: DFT at root^-1, each output times len^-1."""
    n = len(seq)
    ninv = F.mul_inv(F.from_usize(n))
    out = []
    for ri in powers(F, F.mul_inv(root), n):
        acc = F.zero()
        for a, r in zip(seq, powers(F, ri, n)):
            acc = F.add(acc, F.mul(a, r))
        out.append(F.mul(acc, ninv))
    return out


def polynomial_division(F: Field, poly: list, dividend: list):
    """Long division, field/mod.rs:428-469.  Returns (quotient, remainder).

    ``dividend`` is the reference's (confusing) name for the divisor.
    Panics if the divisor is zero (:433-441); returns ([0],[0]) when
    deg(divisor) > deg(poly) (:443-445).
    """
    if all(F.eq(c, F.zero()) for c in dividend):
        raise FieldPanic("Dividend must be non-zero")
    if degree(F, dividend) > degree(F, poly):
        return [F.zero()], [F.zero()]
    poly = remove_leading_zeros(F, poly)
    dividend = remove_leading_zeros(F, dividend)
    q = [F.zero()] * (degree(F, poly) + 1 - degree(F, dividend))
    r = list(poly)
    d = degree(F, dividend)
    c = dividend[d]
    while degree(F, r) >= d and len(r) != 0:
        dr = degree(F, r)
        s = F.div(r[dr], c)
        q[dr - d] = s
        # zip(reversed r skipping leading zeros, reversed (dividend * s))
        top = len(r)
        while top > 0 and F.eq(r[top - 1], F.zero()):
            top -= 1
        scaled = [F.mul(a, s) for a in dividend]
        for k, b in enumerate(reversed(scaled)):
            idx = top - 1 - k
            if idx < 0:
                break
            r[idx] = F.sub(r[idx], b)
        r = remove_leading_zeros(F, r)
    return q, r


# ---- CoefficientPoly operators, src/groth16/coefficient_poly.rs ----------------

def poly_add(F: Field, a: list, b: list) -> list:
    """coefficient_poly.rs:24-49: result length = max(len)."""
    if len(a) < len(b):
        a, b = b, a
    out = list(a)
    for i, c in enumerate(b):
        out[i] = F.add(c, a[i])
    # entries of the longer operand beyond len(shorter) are (0 + a_i)
    for i in range(len(b), len(a)):
        out[i] = F.add(F.from_usize(0), a[i])
    return out


def poly_neg(F: Field, a: list) -> list:
    """coefficient_poly.rs:51-62."""
    return [F.neg(c) for c in a]


def poly_sub(F: Field, a: list, b: list) -> list:
    """coefficient_poly.rs:64-73: self + (-rhs)."""
    return poly_add(F, a, poly_neg(F, b))


def poly_sum(F: Field, polys) -> list:
    """coefficient_poly.rs:75-91: fold from [0] with Add."""
    acc = [F.from_usize(0)]
    for p in polys:
        acc = poly_add(F, acc, p)
    return acc


def poly_mul(F: Field, a: list, b: list) -> list:
    """Schoolbook product, coefficient_poly.rs:93-130.

    Both operands are stripped of leading zeros first (:102-103); the output has
    deg(a)+deg(b)+1 entries, so a zero/empty operand gives [0].
    """
    a = remove_leading_zeros(F, a)
    b = remove_leading_zeros(F, b)
    da, db = degree(F, a), degree(F, b)
    d = da + db + 1
    out = []
    for i in range(d):
        acc = F.from_usize(0)
        # sum_{j} a[i-j]*b[j] over the valid overlap; empty operands contribute nothing
        lo = max(i - da, 0)
        hi = min(i, len(b) - 1)
        for j in range(lo, hi + 1):
            if 0 <= i - j < len(a):
                acc = F.add(acc, F.mul(a[i - j], b[j]))
        out.append(acc)
    return out


def poly_scale(F: Field, a: list, s: int) -> list:
    """coefficient_poly.rs:132-146 (poly * scalar)."""
    return [F.mul(c, s) for c in a]


def poly_div(F: Field, a: list, b: list) -> list:
    """coefficient_poly.rs:148-157: quotient only."""
    return polynomial_division(F, a, b)[0]


def lagrange_basis(F: Field, roots: list, x: int) -> list:
    """coefficient_poly.rs:173-190."""
    acc = [F.from_usize(1)]
    for m in roots:
        if F.eq(m, x):
            continue
        lin = [F.neg(m), F.from_usize(1)]
        lin = poly_scale(F, lin, F.div(F.from_usize(1), F.sub(x, m)))
        acc = poly_mul(F, lin, acc)
    return acc


def poly_from_points(F: Field, roots: list, points) -> list:
    """coefficient_poly.rs:159-171: sum_points lagrange_basis(roots, x) * y."""
    return poly_sum(F, (poly_scale(F, lagrange_basis(F, roots, x), y) for x, y in points))


def root_poly(F: Field, roots: list) -> list:
    """coefficient_poly.rs:192-200: prod (x - r)."""
    acc = [F.from_usize(1)]
    for r in roots:
        acc = poly_mul(F, acc, [F.neg(r), F.from_usize(1)])
    return acc


def ntt_fast(F: Field, seq: list, root: int) -> list:
    """O(n log n) evaluation with the SAME convention as ``dft`` (field/mod.rs:508-520), for
    power-of-two lengths: out[i] = sum_j seq[j]*root^(i*j).  Test infrastructure: lets parity tests
    reach sizes the naive dft cannot; pinned against ``dft`` in tests/test_oracle_kats.py."""
    n = len(seq)
    if n == 1:
        return list(seq)
    assert n % 2 == 0
    r2 = F.mul(root, root)
    even = ntt_fast(F, seq[0::2], r2)
    odd = ntt_fast(F, seq[1::2], r2)
    out = [0] * n
    w = F.one()
    h = n // 2
    for i in range(h):
        t = F.mul(w, odd[i])
        out[i] = F.add(even[i], t)
        out[i + h] = F.sub(even[i], t)
        w = F.mul(w, root)
    return out


def intt_fast(F: Field, seq: list, root: int) -> list:
    """Inverse of ``ntt_fast`` with idft's 1/n scaling (field/mod.rs:524-537)."""
    ninv = F.mul_inv(F.from_usize(len(seq)))
    return [F.mul(x, ninv) for x in ntt_fast(F, seq, F.mul_inv(root))]
