"""Oracle A -- closed-form expectation of a Groth16 proof from the toxic waste (test infrastructure).

With the setup secrets (alpha, beta, gamma, delta, x) known, every CRS element is s * BASE for a known
scalar s (groth16/mod.rs:134-197), so the proof that ``prove`` (mod.rs:213-296) must return is

    A = (u(x) + alpha + r delta) * BASE_G1
    B = (v(x) + beta + s delta) * BASE_G2
    C = (h(x) t(x)/delta + sum_{i>l} a_i lin_i/delta + s A_s + r (beta + v(x) + s delta) - r s delta) * BASE_G1

where u(x) = sum_i a_i u_i(x) etc. and h is the reference's quotient.  On the roots-of-unity domain
u_i(x) = sum_k u_{i,k} L_k(x) with L_k(x) = (x^n - 1) w^k / (n (x - w^k)), which costs O(nnz) field
operations -- a size-independent parity check that reaches 2^16..2^20 constraints where the literal
O(n^2) restatement cannot run.  ``h(x)`` is taken as (u(x) v(x) - w(x)) / t(x), which equals the
reference's quotient evaluated at x exactly when the witness satisfies the circuit (remainder 0).
"""

from __future__ import annotations

from . import bn254 as bn
from .fields import FR

P = FR.p


def batch_inverse(vals: list) -> list:
    pref = [1] * (len(vals) + 1)
    for i, v in enumerate(vals):
        pref[i + 1] = pref[i] * v % P
    inv = pow(pref[-1], -1, P)
    out = [0] * len(vals)
    for i in range(len(vals) - 1, -1, -1):
        out[i] = inv * pref[i] % P
        inv = inv * vals[i] % P
    return out


def lagrange_at(n: int, omega: int, x: int) -> list:
    """[L_0(x), ..., L_{n-1}(x)] on the domain omega^k."""
    tx = (pow(x, n, P) - 1) % P
    ws = [1] * n
    for k in range(1, n):
        ws[k] = ws[k - 1] * omega % P
    den = batch_inverse([(n * (x - w)) % P for w in ws])
    return [tx * w % P * d % P for w, d in zip(ws, den)]


def lagrange_at_1_to_n(n: int, x: int):
    """([L_0(x), ..., L_{n-1}(x)], t(x)) on the parser's domain: gate k (0-based) <-> root k + 1 (circuit/mod.rs:517).
    L_k(x) = t(x) / ((x - r_k) t'(r_k)) with t'(k + 1) = k! (n - 1 - k)! (-1)^(n - 1 - k): O(n)."""
    fact = [1] * (n + 1)
    for i in range(1, n + 1):
        fact[i] = fact[i - 1] * i % P
    tx = 1
    for k in range(1, n + 1):
        tx = tx * (x - k) % P
    dens = []
    for k in range(n):
        tp = fact[k] * fact[n - 1 - k] % P
        if (n - 1 - k) & 1:
            tp = P - tp
        dens.append((x - (k + 1)) * tp % P)
    inv = batch_inverse(dens)
    return [tx * d % P for d in inv], tx


def row_evals(rows, L, index_of_root=None):
    """rows: per wire list of (gate index, coeff) -> [row_i(x)]."""
    return [sum(c * L[g] for g, c in row) % P for row in rows]


def expected_proof_scalars(n, omega, rows_u, rows_v, rows_w, n_input, weights, toxic, r, s):
    """Scalars (a, b, c) with A = a*BASE_G1, B = b*BASE_G2, C = c*BASE_G1.  rows_*: per wire list of
    (gate index, coeff).  Valid witnesses only (see module docstring).  omega = None: the parser's domain 1..=n."""
    alpha, beta, gamma, delta, x = toxic
    if omega is None:
        L, tx_generic = lagrange_at_1_to_n(n, x)
    else:
        L = lagrange_at(n, omega, x)
    ux, vx, wx = row_evals(rows_u, L), row_evals(rows_v, L), row_evals(rows_w, L)
    m = len(rows_u)
    a = list(weights) + [0] * max(0, m - len(weights))
    U = sum(ai * e for ai, e in zip(a, ux)) % P
    V = sum(ai * e for ai, e in zip(a, vx)) % P
    W = sum(ai * e for ai, e in zip(a, wx)) % P
    dinv = pow(delta, -1, P)
    hx_t = (U * V - W) % P  # h(x) * t(x): t(x) itself is not needed
    lin = [(beta * u + alpha * v + w) % P for u, v, w in zip(ux, vx, wx)]
    wit = sum(a[i] * lin[i] for i in range(n_input + 1, m)) % P * dinv % P
    A = (U + alpha + r * delta) % P
    B = (V + beta + s * delta) % P
    Cc = (hx_t * dinv + wit + s * A + r * ((beta + V + s * delta) % P) - r * s % P * delta) % P
    return A, B, Cc


def expected_proof(*args):
    a, b, c = expected_proof_scalars(*args)
    return bn.g1_mul(bn.BASE_G1, a), bn.g2_mul(bn.BASE_G2, b), bn.g1_mul(bn.BASE_G1, c)
