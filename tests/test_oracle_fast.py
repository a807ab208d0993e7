"""Pin Oracle F (oracle/oracle_fast.c: NTT + Pippenger on host threads, bench.py's `cpu_best_effort` context number)
against Oracle A (the literal restatement, itself pinned to the reference's golden vectors) and Oracle B.  CPU only."""

import random

import pytest

from oracle import bn254 as bn, groth16 as g, oracle_b as ob, oracle_fast as of, poly, synthetic
from oracle.fields import FR

P = FR.p


@pytest.mark.parametrize("log_n,threads", [(0, 1), (1, 1), (3, 1), (6, 3), (13, 4)])
def test_ntt_is_the_reference_dft(log_n, threads):
    rng = random.Random(log_n)
    n = 1 << log_n
    x = [rng.randrange(P) for _ in range(n)]
    w = synthetic.omega(log_n)
    y = of.ntt(x, threads=threads)
    if n <= 64:
        assert y == poly.dft(FR, x, w)  # field/mod.rs:508-523
    else:
        assert y[0] == sum(x) % P
        for i in (1, n // 2, n - 1):
            assert y[i] == sum(v * pow(w, i * j, P) for j, v in enumerate(x)) % P
    assert of.ntt(y, inverse=True, threads=threads) == x  # idft, :525-537


@pytest.mark.parametrize("n,threads", [(1, 1), (7, 2), (300, 1), (3000, 5)])
def test_pippenger_matches_per_term_fold(n, threads):
    rng = random.Random(n)
    ks = [rng.randrange(P) for _ in range(max(n, 8))]
    sc = [rng.choice([0, 1, P - 1, rng.randrange(P), rng.randrange(P)]) for _ in range(n)]
    pts = [bn.g1_mul(bn.BASE_G1, k) for k in ks[:8]]
    pts = (pts + [None, pts[0], bn.g1_neg(pts[0])] + pts * (n // 8 + 1))[:n]  # identity, repeated and opposite points
    if n <= 300:
        assert of.msm_g1(sc, pts, threads) == ob.msm_g1(sc, pts)
    else:  # closed form: every point is a known multiple of the base point
        kk = (ks[:8] + [0, ks[0], P - ks[0]] + ks[:8] * (n // 8 + 1))[:n]
        assert of.msm_g1(sc, pts, threads) == bn.g1_mul(bn.BASE_G1, sum(a * b for a, b in zip(sc, kk)) % P)
    m = min(n, 40)
    p2 = [bn.g2_mul(bn.BASE_G2, k) for k in ks[:4]]
    p2 = (p2 + [None, p2[1], bn.g2_neg(p2[1])] + p2 * (m // 4 + 1))[:m]
    assert of.msm_g2(sc[:m], p2, threads) == ob.msm_g2(sc[:m], p2)


@pytest.mark.parametrize("n,valid,threads", [(2, True, 1), (4, True, 2), (8, False, 3), (16, True, 8)])
def test_prove_equals_the_literal_restatement(n, valid, threads):
    log_n = n.bit_length() - 1
    w = synthetic.omega(log_n)
    roots = [pow(w, k, P) for k in range(n)]
    rep = synthetic.horner_rep(FR, n, roots)
    qap = g.qap_from_root_rep(FR, rep)
    rng = random.Random(n)
    wit = synthetic.horner_witness(FR, n, rng.randrange(1, P), [rng.randrange(P) for _ in range(n)])
    if not valid:
        wit[5] = (wit[5] + 3) % P
    toxic = tuple(rng.randrange(1, P) for _ in range(5))
    r, s = rng.randrange(1, P), rng.randrange(1, P)
    B = g.BN254Backend()
    sig = g.setup(B, qap, toxic)
    want = g.prove(B, qap, sig, wit, r, s)
    idx = {root: k for k, root in enumerate(roots)}
    ev = lambda rows: [sum(wit[i] * val for i, row in enumerate(rows) for (root, val) in row if idx[root] == k) % P for k in range(n)]
    a, b, c, h = of.prove(n, rep.input, ev(rep.u), ev(rep.v), sig, wit, r, s, threads)
    assert (a, b, c) == (want.a, want.b, want.c)
    u, v, ws = g.weighted_sums(FR, qap, wit)
    hq = g.quotient_h(FR, qap, u, v, ws)
    assert h[:len(hq)] == hq and not any(h[len(hq):])
    assert ob.prove(qap, sig, wit, r, s)[:3] == (a, b, c)


def test_timed_proof_is_deterministic_across_thread_counts():
    t1, parts, chk1 = of.time_prove(6, 1, seed=3)
    t2, _, chk2 = of.time_prove(6, 4, seed=3)
    assert chk1 == chk2 and chk1 != 0 and t1 > 0 and t2 > 0 and set(parts) == {"poly", "g1", "g2"}


def test_prove_512_gates_equals_closed_form():
    """Beyond the sizes the literal restatement reaches: a 512-gate Horner proof on 4 threads (several windows x point
    slices per MSM, block-local + global NTT stages) against the closed form from the toxic waste."""
    import types

    from oracle import closed_form as cf

    n, log_n, threads = 512, 9, 4
    w = synthetic.omega(log_n)
    roots = [pow(w, k, P) for k in range(n)]
    rep = synthetic.horner_rep(FR, n, roots)
    rng = random.Random(512)
    wit = synthetic.horner_witness(FR, n, rng.randrange(1, P), [rng.randrange(P) for _ in range(n)])
    alpha, beta, gamma, delta, x = toxic = tuple(rng.randrange(1, P) for _ in range(5))
    r, s = rng.randrange(1, P), rng.randrange(1, P)
    idx = {root: k for k, root in enumerate(roots)}
    by_gate = lambda rows: [[(idx[root], val) for (root, val) in row] for row in rows]
    ru, rv, rw = by_gate(rep.u), by_gate(rep.v), by_gate(rep.w)
    # the CRS of groth16/mod.rs:134-197 from its defining scalars (C scalar multiplications)
    L = cf.lagrange_at(n, w, x)
    ux, vx, wx = cf.row_evals(ru, L), cf.row_evals(rv, L), cf.row_evals(rw, L)
    dinv, tx = pow(delta, -1, P), (pow(x, n, P) - 1) % P
    g1 = lambda k: ob.g1_mul(bn.BASE_G1, k % P)
    g2 = lambda k: ob.g2_mul(bn.BASE_G2, k % P)
    xs = [pow(x, j, P) for j in range(n)]
    m = len(ru)
    s1 = types.SimpleNamespace(alpha=g1(alpha), beta=g1(beta), delta=g1(delta), xi=[g1(v) for v in xs],
                               xi_t=[g1(v * tx % P * dinv) for v in xs[:n - 1]],
                               sum_delta=[g1((beta * ux[i] + alpha * vx[i] + wx[i]) % P * dinv) for i in range(rep.input + 1, m)])
    s2 = types.SimpleNamespace(beta=g2(beta), delta=g2(delta), xi=[g2(v) for v in xs])
    ev = lambda rows: [sum(wit[i] * val for i, val in per_gate) % P for per_gate in rows]
    gates_u, gates_v = [[] for _ in range(n)], [[] for _ in range(n)]
    for i, row in enumerate(ru):
        for k, val in row:
            gates_u[k].append((i, val))
    for i, row in enumerate(rv):
        for k, val in row:
            gates_v[k].append((i, val))
    a, b, c, h = of.prove(n, rep.input, ev(gates_u), ev(gates_v), (s1, s2), wit, r, s, threads)
    assert (a, b, c) == cf.expected_proof(n, w, ru, rv, rw, rep.input, wit, toxic, r, s)
    assert len(h) == n - 1 and sum(hk * pow(x, k, P) for k, hk in enumerate(h)) % P * tx % P == \
        (sum(ai * e for ai, e in zip(wit, ux)) * sum(ai * e for ai, e in zip(wit, vx)) - sum(ai * e for ai, e in zip(wit, wx))) % P


@pytest.mark.parametrize("n,valid", [(4, True), (16, False), (64, True)])
def test_array_level_entry_points_equal_the_list_level_ones(n, valid):
    """of.qap_evals_np / qap_h_np / prove_np / ntt_np (uint64 limb arrays: what tests/test_gpu_parity_large.py compares the
    device with at 2^12 .. 2^22) against the list-level functions pinned above and the literal restatement."""
    import importlib

    import numpy as np
    zg = importlib.import_module("zksnark-rs_b200.groth16")
    log_n = n.bit_length() - 1
    w = synthetic.omega(log_n)
    roots = [pow(w, k, P) for k in range(n)]
    rep = synthetic.horner_rep(FR, n, roots)
    qap = g.qap_from_root_rep(FR, rep)
    rng = random.Random(7 * n)
    wit = synthetic.horner_witness(FR, n, rng.randrange(1, P), [rng.randrange(P) for _ in range(n)])
    if not valid:
        wit[3] = (wit[3] + 1) % P
        wit[-1] = rng.randrange(P)
    m, n_input, rows = zg.horner_qap_rows(n)
    assert m == len(rep.u) and n_input == rep.input
    wl = zg.fr_limbs(wit)
    A, B = of.qap_evals_np(m, n, rows, wl)
    idx = {root: k for k, root in enumerate(roots)}
    ev = lambda rws: [sum(wit[i] * val for i, row in enumerate(rws) for (root, val) in row if idx[root] == k) % P for k in range(n)]
    assert zg.limbs_to_ints(A) == ev(rep.u) and zg.limbs_to_ints(B) == ev(rep.v)
    u, v, h = of.qap_h_np(A, B, threads=3)
    uu, vv, ws = g.weighted_sums(FR, qap, wit)
    hq = g.quotient_h(FR, qap, uu, vv, ws)  # literal schoolbook Mul + long division, remainder discarded
    pad = lambda p, k: (list(p) + [0] * k)[:k]
    assert zg.limbs_to_ints(u) == pad(uu, n) and zg.limbs_to_ints(v) == pad(vv, n) and zg.limbs_to_ints(h) == pad(hq, n)
    # short witness: zip truncation
    A2, _ = of.qap_evals_np(m, n, rows, wl[: m - 2])
    assert zg.limbs_to_ints(A2) == ev([row if i < m - 2 else [] for i, row in enumerate(rep.u)])
    if n <= 16:
        toxic = tuple(rng.randrange(1, P) for _ in range(5))
        r, s = rng.randrange(1, P), rng.randrange(1, P)
        Bk = g.BN254Backend()
        s1, s2 = g.setup(Bk, qap, toxic)
        want = g.prove(Bk, qap, (s1, s2), wit, r, s)
        crs = {"alpha1": zg.g1_pack([s1.alpha]), "beta1": zg.g1_pack([s1.beta]), "delta1": zg.g1_pack([s1.delta]),
               "xi1": zg.g1_pack(s1.xi), "xi_t": zg.g1_pack(s1.xi_t), "sum_delta": zg.g1_pack(s1.sum_delta),
               "beta2": zg.g2_pack([s2.beta]), "delta2": zg.g2_pack([s2.delta]), "xi2": zg.g2_pack(s2.xi)}
        got = of.prove_np(n, n_input, A, B, crs, wl, r, s, threads=2)
        assert zg.g1_unpack(got[0:8])[0] == want.a and zg.g2_unpack(got[8:24])[0] == want.b and zg.g1_unpack(got[24:32])[0] == want.c
    x = [rng.randrange(P) for _ in range(n)]
    xl = zg.fr_limbs(x)
    assert zg.limbs_to_ints(of.ntt_np(xl, threads=2)) == poly.dft(FR, x, w)
    assert zg.limbs_to_ints(of.ntt_np(xl, inverse=True, threads=2)) == poly.idft(FR, x, w)
    sh = 7
    assert zg.limbs_to_ints(of.ntt_np(xl, coset_shift=sh)) == poly.dft(FR, [a * pow(sh, i, P) % P for i, a in enumerate(x)], w)
    assert np.array_equal(of.ntt_np(of.ntt_np(xl, coset_shift=sh), inverse=True, coset_shift=sh), xl)
