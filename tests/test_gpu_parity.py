"""GPU parity tests: the CUDA path (through the C ABI of libzkb200.so) against the CPU oracle.

Bit-exact comparisons on canonical residues and affine coordinates (integer arithmetic: no
tolerance anywhere).  Small sizes are compared with the literal restatement of the reference
(oracle.groth16 / oracle.poly / oracle.bn254); full sizes use size-independent properties
(closed-form proof from the toxic waste, sum_i s_i k_i collapse for MSMs, transform round trips).
"""

import importlib
import random

import numpy as np
import pytest

from oracle import bn254 as bn
from oracle import closed_form as cf
from oracle import groth16 as og
from oracle import circuit, poly, synthetic
from oracle.fields import FR

pytestmark = pytest.mark.gpu
zk = importlib.import_module("zksnark-rs_b200")
zg = importlib.import_module("zksnark-rs_b200.groth16")
P = FR.p


@pytest.fixture(scope="module")
def ctx():
    c = zk.Context(0)
    yield c
    c.close()


def rand_fr(rng, nonzero=False):
    return rng.randrange(1 if nonzero else 0, P)


# ------------------------------------------------------------------------------------------------
# NTT  (reference convention: field/mod.rs:508-537)
@pytest.mark.parametrize("log_n", [1, 2, 3, 5, 6])
def test_ntt_matches_naive_dft(ctx, log_n):
    rng = random.Random(100 + log_n)
    n = 1 << log_n
    x = [rand_fr(rng) for _ in range(n)]
    w = synthetic.omega(log_n)
    assert zk.ntt(ctx, x) == poly.dft(FR, x, w)
    assert zk.ntt(ctx, x, inverse=True) == poly.idft(FR, x, w)


@pytest.mark.parametrize("log_n", [9, 10, 11, 13, 15, 16])
def test_ntt_matches_fast_oracle(ctx, log_n):
    rng = random.Random(200 + log_n)
    n = 1 << log_n
    x = [rand_fr(rng) for _ in range(n)]
    w = synthetic.omega(log_n)
    y = zk.ntt(ctx, x)
    assert y == poly.ntt_fast(FR, x, w)
    assert zk.ntt(ctx, y, inverse=True) == x


def test_ntt_coset_and_edge_sizes(ctx):
    rng = random.Random(3)
    assert zk.ntt(ctx, [5]) == [5] and zk.ntt(ctx, [5], inverse=True) == [5]
    n, g = 64, 7
    x = [rand_fr(rng) for _ in range(n)]
    shifted = [a * pow(g, i, P) % P for i, a in enumerate(x)]
    y = zk.ntt(ctx, x, coset_shift=g)
    assert y == poly.dft(FR, shifted, synthetic.omega(6))
    assert zk.ntt(ctx, y, inverse=True, coset_shift=g) == x
    # all-zero and all-(r-1) vectors
    assert zk.ntt(ctx, [0] * 16) == [0] * 16
    top = [P - 1] * 16
    assert zk.ntt(ctx, top) == poly.dft(FR, top, synthetic.omega(4))


@pytest.mark.parametrize("log_n", [20, 22])
def test_ntt_roundtrip_and_linearity_full_size(ctx, log_n):
    """BASELINE configs 2^20 / 2^22: iNTT(NTT(x)) == x and NTT(x)[0] == sum(x), NTT(delta_1) == powers."""
    n = 1 << log_n
    rng = np.random.default_rng(log_n)
    a = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64((1 << 60) - 1)  # < 2^252 < r: canonical residues
    d = ctx.dev_alloc(a.nbytes)
    try:
        ctx.h2d(d, a)
        lib = ctx.lib
        import ctypes as C
        ctx.check(lib.zkb_ntt_fr(ctx.h, C.c_void_p(d), log_n, 0, None), "ntt")
        y = np.empty_like(a)
        ctx.d2h(y, d)
        # X[0] = sum of inputs (column sums of 32-bit halves stay below 2^64)
        lo = (a & np.uint64(0xFFFFFFFF)).sum(axis=0, dtype=np.uint64)
        hi = (a >> np.uint64(32)).sum(axis=0, dtype=np.uint64)
        tot = sum((int(lo[j]) + (int(hi[j]) << 32)) << (64 * j) for j in range(4)) % P
        assert zg.limbs_to_ints(y[:1])[0] == tot
        ctx.check(lib.zkb_ntt_fr(ctx.h, C.c_void_p(d), log_n, 1, None), "intt")
        ctx.d2h(y, d)
        assert np.array_equal(y, a)
        # delta at index 1 -> out[i] = omega^i
        e = np.zeros_like(a)
        e[1, 0] = 1
        ctx.h2d(d, e)
        ctx.check(lib.zkb_ntt_fr(ctx.h, C.c_void_p(d), log_n, 0, None), "ntt")
        ctx.d2h(y, d)
        w = synthetic.omega(log_n)
        for i in [0, 1, 2, 3, 1000, n // 2, n - 1]:
            assert zg.limbs_to_ints(y[i:i + 1])[0] == pow(w, i, P)
    finally:
        ctx.dev_free(d)


@pytest.mark.parametrize("log_n,log_g", [(4, 1), (10, 2), (13, 3), (3, 3), (11, 0)])
def test_ntt_outer_dimension_sharded(ctx, log_n, log_g):
    """Multi-GPU transform: G ranks transform the decimated subsequences x[g::G]; the all-gathered partial
    transforms are combined slice by slice (zkb_ntt_combine).  One device plays every rank here; the exchange itself
    is covered by tools/ntt_multi_gpu.py on real ranks.  Forward and inverse equal the single-device transform."""
    rng = random.Random(900 + log_n)
    n, G = 1 << log_n, 1 << log_g
    sub = n // G
    x = [rand_fr(rng) for _ in range(n)]
    for inverse in (False, True):
        want = zk.ntt(ctx, x, inverse=inverse)
        parts = []
        for g in range(G):
            xs = x[g::G]
            parts += zk.ntt(ctx, xs, inverse=inverse) if sub > 1 else xs
        d_parts = ctx.dev_alloc(32 * n)
        d_out = ctx.dev_alloc(32 * sub)
        ctx.h2d(d_parts, zg.fr_limbs(parts))
        got = []
        for r in range(G):
            zg.ntt_combine(ctx, d_parts, log_n, log_g, inverse, r * sub, sub, d_out)
            buf = np.zeros((sub, 4), dtype=np.uint64)
            ctx.d2h(buf, d_out)
            got += zg.limbs_to_ints(buf)
        ctx.dev_free(d_parts)
        ctx.dev_free(d_out)
        assert got == want


# ------------------------------------------------------------------------------------------------
# device field arithmetic (crate bn's Fr / Fq / Fq2 as reached through fr.rs:18-56)
@pytest.mark.parametrize("field", [0, 1])
def test_field_ops_match_bigint(ctx, field):
    """mul (CIOS), sqr (36-product wide square + REDC), a*b - c*d (two wide products, one REDC), add,
    sub, inverse on edge values and 4096 random elements, against Python integers."""
    p = FR.p if field == 0 else bn.Q
    rng = random.Random(300 + field)
    edge = [0, 1, 2, p - 1, p - 2, (p - 1) // 2, (p + 1) // 2, (1 << 253), (1 << 224) - 1, (1 << 32) - 1, 1 << 32]
    a = edge * len(edge) + [rng.randrange(p) for _ in range(4096)]
    b = [y for y in edge for _ in edge] + [rng.randrange(p) for _ in range(4096)]
    c = list(reversed(a))
    d = list(reversed(b))
    assert zg.field_op(ctx, field, 0, a, b) == [x * y % p for x, y in zip(a, b)]
    assert zg.field_op(ctx, field, 1, a) == [x * x % p for x in a]
    assert zg.field_op(ctx, field, 2, a, b, c, d) == [(x * y - z * w) % p for x, y, z, w in zip(a, b, c, d)]
    assert zg.field_op(ctx, field, 3, a, b) == [(x + y) % p for x, y in zip(a, b)]
    assert zg.field_op(ctx, field, 4, a, b) == [(x - y) % p for x, y in zip(a, b)]
    assert zg.field_op(ctx, field, 5, a[:200]) == [pow(x, p - 2, p) for x in a[:200]]
    # the product as wide product + stand-alone reduction (what the Fq2 arithmetic is built from) and its Karatsuba form
    half = [(1 << 128) - 1, 1 << 128, ((1 << 125) - 1) << 128, (1 << 253) | ((1 << 128) - 1)]
    aa, bb = a + [x % p for x in half for _ in half], b + [y % p for _ in half for y in half]
    for op in (6, 7):
        assert zg.field_op(ctx, field, op, aa, bb) == [x * y % p for x, y in zip(aa, bb)]


def test_fq2_ops_match_bigint(ctx):
    """Fq2 = Fq[u]/(u^2+1): Karatsuba on unreduced 512-bit products (3 products, 2 reductions)."""
    p = bn.Q
    rng = random.Random(310)
    edge = [0, 1, p - 1, (p - 1) // 2, 1 << 253]
    a = [(x, y) for x in edge for y in edge] + [(rng.randrange(p), rng.randrange(p)) for _ in range(2048)]
    b = list(reversed(a[:25])) + [(rng.randrange(p), rng.randrange(p)) for _ in range(2048)]
    mul = lambda x, y: ((x[0] * y[0] - x[1] * y[1]) % p, (x[0] * y[1] + x[1] * y[0]) % p)
    assert zg.field_op(ctx, 2, 0, a, b) == [mul(x, y) for x, y in zip(a, b)]
    assert zg.field_op(ctx, 2, 1, a) == [mul(x, x) for x in a]
    inv = zg.field_op(ctx, 2, 2, a[25:125])
    assert [mul(x, y) for x, y in zip(a[25:125], inv)] == [(1, 0)] * 100
    # a b - c d with two reductions (both Karatsuba products unreduced): the last step of the G2 mixed addition.
    # Edge operands drive every intermediate to the ends of its range (all-(q-1): re = 0 mod q with the 2 q^2 offset).
    c = [(y, x) for x, y in a[:25]] + [(rng.randrange(p), rng.randrange(p)) for _ in range(2048)]
    d = [a[(7 * i) % 25] for i in range(25)] + [(rng.randrange(p), rng.randrange(p)) for _ in range(2048)]
    sub = lambda x, y: ((x[0] - y[0]) % p, (x[1] - y[1]) % p)
    assert zg.field_op(ctx, 2, 3, a, b, c, d) == [sub(mul(x, y), mul(z, w)) for x, y, z, w in zip(a, b, c, d)]
    top = [(p - 1, p - 1)] * 8 + [(p - 1, 0), (0, p - 1), (0, 0), (1, p - 1)]
    for rot in range(4):
        ops = [top[rot:] + top[:rot], top, top[::-1], top[rot:] + top[:rot]]
        assert zg.field_op(ctx, 2, 3, *ops) == [sub(mul(x, y), mul(z, w)) for x, y, z, w in zip(*ops)]


# ------------------------------------------------------------------------------------------------
# fixed-base generation + MSM
def test_bases_generate_matches_oracle(ctx):
    rng = random.Random(5)
    ks = [0, 1, 2, P - 1, 69] + [rand_fr(rng) for _ in range(6)]
    g1 = zk.Bases.generate(ctx, 1, ks).download()
    g2 = zk.Bases.generate(ctx, 2, ks).download()
    assert g1 == [bn.g1_mul(bn.BASE_G1, k) for k in ks]
    assert g2 == [bn.g2_mul(bn.BASE_G2, k) for k in ks]
    assert g1[0] is None and g2[0] is None  # encrypt_g1(0) is the identity (mod.rs:407)


@pytest.mark.parametrize("group", [1, 2])
def test_msm_small_adversarial(ctx, group):
    """Zero scalars, 1, r-1, repeated points, P and -P, identity among the bases."""
    rng = random.Random(40 + group)
    base = bn.BASE_G1 if group == 1 else bn.BASE_G2
    mul = bn.g1_mul if group == 1 else bn.g2_mul
    neg = bn.g1_neg if group == 1 else bn.g2_neg
    ora = bn.msm_g1 if group == 1 else bn.msm_g2
    Pt = mul(base, 12345)
    pts = [Pt, Pt, neg(Pt), None, mul(base, 7), mul(base, 7), base, mul(base, P - 1)]
    pts += [mul(base, rand_fr(rng)) for _ in range(8)]
    scal = [5, 5, 10, 999, 0, 1, P - 1, P - 1] + [rand_fr(rng) for _ in range(8)]
    b = zk.Bases.upload(ctx, group, pts)
    for c in (0, 2, 3, 5, 8, 13):
        assert zk.msm(ctx, b, scal, window_bits=c) == ora(scal, pts), f"window {c}"
    # everything cancels -> identity
    assert zk.msm(ctx, b, [3, 4, 7] + [0] * 13) is None
    # zip truncation: fewer scalars than bases; and the empty sum is the identity (fr.rs:196)
    assert zk.msm(ctx, b, scal[:3]) == ora(scal[:3], pts[:3])
    assert zk.msm(ctx, b, []) is None
    # same scalar everywhere (one bucket per window takes all points)
    assert zk.msm(ctx, b, [scal[9]] * 16) == ora([scal[9]] * 16, pts)


@pytest.mark.parametrize("group,log_n", [(1, 8), (1, 12), (2, 10), (1, 16), (2, 14)])
def test_msm_collapse_property(ctx, group, log_n):
    """bases P_i = k_i * BASE  =>  sum s_i P_i = (sum s_i k_i) * BASE.  Size independent."""
    n = 1 << log_n
    rng = random.Random(log_n * 10 + group)
    ks = [rand_fr(rng) for _ in range(n)]
    ss = [rand_fr(rng) for _ in range(n)]
    # witness-like skew: a block of zeros and ones
    for i in range(0, n, 5):
        ss[i] = i % 2
    b = zk.Bases.generate(ctx, group, ks)
    e = sum(s * k for s, k in zip(ss, ks)) % P
    want = bn.g1_mul(bn.BASE_G1, e) if group == 1 else bn.g2_mul(bn.BASE_G2, e)
    assert zk.msm(ctx, b, ss) == want


def test_msm_2pow20_collapse(ctx):
    """BASELINE config 4 size (2^20 points), same collapse property, numpy-generated inputs."""
    n = 1 << 20
    rng = np.random.default_rng(20)
    k = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64)
    s = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64)
    k[:, 3] &= np.uint64((1 << 60) - 1)
    s[:, 3] &= np.uint64((1 << 60) - 1)
    b = zk.Bases.generate(ctx, 1, k)
    ki, si = zg.limbs_to_ints(k), zg.limbs_to_ints(s)
    e = sum(x * y for x, y in zip(ki, si)) % P
    assert zk.msm(ctx, b, s) == bn.g1_mul(bn.BASE_G1, e)


@pytest.mark.parametrize("group,log_n,world", [(1, 10, 2), (1, 12, 4), (2, 9, 3), (1, 6, 8)])
def test_msm_window_sharded_equals_single(ctx, group, log_n, world):
    """Window sharding (BASELINE config 4): rank g takes the table rows j = g (mod world); the `world` partial
    points fold (zkb_points_sum, what the ranks do after the NCCL all-gather) to the single-GPU result, including
    for more ranks than some scalars have non-zero windows."""
    n = 1 << log_n
    rng = random.Random(700 + log_n + group)
    ks = [rand_fr(rng) for _ in range(n)]
    ss = [rand_fr(rng) for _ in range(n)]
    for i in range(0, n, 7):
        ss[i] = i % 3  # witness-like small scalars: only window 0 is non-zero
    b = zk.Bases.generate(ctx, group, ks)
    full = zk.msm(ctx, b, ss)
    parts = [zk.msm(ctx, b, ss, windows=(g, world)) for g in range(world)]
    assert zg.points_sum(ctx, group, parts) == full
    e = sum(s * k for s, k in zip(ss, ks)) % P
    assert full == (bn.g1_mul(bn.BASE_G1, e) if group == 1 else bn.g2_mul(bn.BASE_G2, e))
    with pytest.raises(zk.ZkbError):
        zk.msm(ctx, b, ss, windows=(world, world))


def test_points_sum(ctx):
    pts = [bn.g1_mul(bn.BASE_G1, k) for k in (3, 5, 7)] + [None]
    assert zg.points_sum(ctx, 1, pts) == bn.g1_mul(bn.BASE_G1, 15)
    q = [bn.g2_mul(bn.BASE_G2, k) for k in (3, P - 3)]
    assert zg.points_sum(ctx, 2, q) is None


# ------------------------------------------------------------------------------------------------
# QAP / setup / prove against the literal restatement
def _horner_case(n, seed, valid=True):
    rng = random.Random(seed)
    log_n = n.bit_length() - 1
    w = synthetic.omega(log_n)
    roots = [pow(w, k, P) for k in range(n)]
    rep = synthetic.horner_rep(FR, n, roots)
    x, cs = rand_fr(rng, True), [rand_fr(rng) for _ in range(n)]
    wit = synthetic.horner_witness(FR, n, x, cs)
    if not valid:
        wit[3] = (wit[3] + 1) % P
        wit[-1] = rand_fr(rng)
    toxic = tuple(rand_fr(rng, True) for _ in range(5))
    r, s = rand_fr(rng, True), rand_fr(rng, True)
    return rep, wit, toxic, r, s


@pytest.mark.parametrize("n", [2, 4, 8, 16])
@pytest.mark.parametrize("valid", [True, False])
def test_h_matches_reference_quotient(ctx, n, valid):
    """u_sum, v_sum (mod.rs:233-246) and h = (u*v - w)/t (mod.rs:277) vs schoolbook Mul + long division,
    for satisfying AND non-satisfying witnesses (the reference discards the remainder)."""
    rep, wit, *_ = _horner_case(n, 7 * n + valid, valid)
    dense = og.qap_from_root_rep(FR, rep)
    u, v, w = og.weighted_sums(FR, dense, wit)
    h = og.quotient_h(FR, dense, u, v, w)
    q = zk.QAP.from_root_representation(ctx, rep)
    gu, gv, gh = zk.qap_h(ctx, q, wit)
    pad = lambda p, k: (list(p) + [0] * k)[:k]
    assert gu == pad(u, n) and gv == pad(v, n)
    assert gh == pad(h, n - 1)


def test_horner_rows_match_oracle_rep(ctx):
    n = 8
    rep, wit, toxic, r, s = _horner_case(n, 1)
    q1 = zk.QAP.from_root_representation(ctx, rep)
    q2 = zk.QAP.horner(ctx, n)
    assert zk.qap_h(ctx, q1, wit) == zk.qap_h(ctx, q2, wit)
    assert zg.horner_witness(n, wit[1], [wit[4], wit[6], wit[8], wit[10], wit[12], wit[14], wit[16], wit[17]]) == wit


@pytest.mark.parametrize("n", [2, 4, 8])
def test_setup_matches_reference(ctx, n):
    rep, wit, toxic, r, s = _horner_case(n, 50 + n)
    dense = og.qap_from_root_rep(FR, rep)
    s1, s2 = og.setup(og.BN254Backend(), dense, toxic)
    q = zk.QAP.from_root_representation(ctx, rep)
    crs = zk.setup(ctx, q, toxic).download()
    assert crs["alpha1"] == s1.alpha and crs["beta1"] == s1.beta and crs["delta1"] == s1.delta
    assert crs["xi1"] == s1.xi and crs["xi_t"] == s1.xi_t
    assert crs["sum_gamma"] == s1.sum_gamma and crs["sum_delta"] == s1.sum_delta
    assert crs["beta2"] == s2.beta and crs["gamma2"] == s2.gamma and crs["delta2"] == s2.delta
    assert crs["xi2"] == s2.xi
    assert len(crs["xi_t"]) == n - 1 and len(crs["sum_gamma"]) == rep.input + 1  # mod.rs:154,168


@pytest.mark.parametrize("n,valid", [(2, True), (4, True), (8, True), (8, False), (16, True)])
def test_prove_matches_reference(ctx, n, valid):
    """Bit-exact Proof{a,b,c} vs the literal restatement of groth16::prove, CRS uploaded from the
    oracle's setup() (so prove is tested independently of the device setup)."""
    rep, wit, toxic, r, s = _horner_case(n, 90 + n, valid)
    B = og.BN254Backend()
    dense = og.qap_from_root_rep(FR, rep)
    sig = og.setup(B, dense, toxic)
    want = og.prove(B, dense, sig, wit, r, s)
    q = zk.QAP.from_root_representation(ctx, rep)
    crs = zk.CRS.upload(ctx, sig[0], sig[1])
    got = zk.prove(ctx, q, crs, wit, r, s)
    assert (got.a, got.b, got.c) == (want.a, want.b, want.c)
    if valid:
        assert og.verify(B, sig, wit[1:rep.input + 1], og.Proof(got.a, got.b, got.c))
    assert zk.verify(ctx, crs, wit[1:rep.input + 1], got) == valid  # groth16::verify on the device (mod.rs:299-320)
    # zip truncation of the weights (mod.rs:237..288): a short witness == zero-padded witness
    short = wit[: len(wit) - 2]
    want2 = og.prove(B, dense, sig, short, r, s)
    got2 = zk.prove(ctx, q, crs, short, r, s)
    assert (got2.a, got2.b, got2.c) == (want2.a, want2.b, want2.c)


def test_single_gate_qap(ctx):
    """n = 1: xi_t is empty (mod.rs:168, asserted :412) and h = [0].  One gate y = x * c1 on the root {1} (the 1st
    root of unity and ASTParser's 1..=n coincide; served by the explicit-roots path)."""
    rng = random.Random(31)
    one = FR.from_usize(1)
    rep = circuit.DummyRep(u=[[], [(1, one)], [], []], v=[[], [], [], [(1, one)]], w=[[], [], [(1, one)], []], roots=[1], input=2)
    x, c1 = rand_fr(rng, True), rand_fr(rng, True)
    wit = [1, x, x * c1 % P, c1]
    toxic = tuple(rand_fr(rng, True) for _ in range(5))
    r, s = rand_fr(rng, True), rand_fr(rng, True)
    B = og.BN254Backend()
    dense = og.qap_from_root_rep(FR, rep)
    sig = og.setup(B, dense, toxic)
    assert sig[0].xi_t == []
    want = og.prove(B, dense, sig, wit, r, s)
    q = zk.QAP.from_root_representation(ctx, rep)
    crs = zk.setup(ctx, q, toxic)
    d = crs.download()
    assert d["xi_t"] == [] and d["xi1"] == sig[0].xi and d["sum_delta"] == sig[0].sum_delta and d["sum_gamma"] == sig[0].sum_gamma
    got = zk.prove(ctx, q, crs, wit, r, s)
    assert (got.a, got.b, got.c) == (want.a, want.b, want.c)
    assert zk.verify(ctx, crs, wit[1:3], got)
    assert not zk.verify(ctx, crs, [wit[1], (wit[2] + 1) % P], got)


def test_prove_with_identity_in_crs(ctx):
    """encrypt_g1(0) entries (identity points) are legal CRS members (mod.rs:407)."""
    n = 4
    rep, wit, toxic, r, s = _horner_case(n, 5)
    B = og.BN254Backend()
    dense = og.qap_from_root_rep(FR, rep)
    s1, s2 = og.setup(B, dense, toxic)
    s1.sum_delta[1] = None
    s1.xi[2] = None
    s2.xi[1] = None
    want = og.prove(B, dense, (s1, s2), wit, r, s)
    q = zk.QAP.from_root_representation(ctx, rep)
    got = zk.prove(ctx, q, zk.CRS.upload(ctx, s1, s2), wit, r, s)
    assert (got.a, got.b, got.c) == (want.a, want.b, want.c)


def _rows_from_csr(rows):
    out = []
    for ptr, gate, coeff in rows:
        cs = zg.limbs_to_ints(coeff)
        out.append([[(int(gate[e]), cs[e]) for e in range(int(ptr[i]), int(ptr[i + 1]))] for i in range(len(ptr) - 1)])
    return out


@pytest.mark.parametrize("log_n", [6, 10, 12, 16, 20])
def test_prove_closed_form_and_pairing(ctx, log_n):
    """Device setup + prove on the synthetic Horner QAP (BASELINE config 2 at 2^16, config 3 at 2^20) against the
    closed-form proof from the toxic waste (bit-exact) and, for 2^6, the reference's own
    acceptance criterion verify(...) == true (lib.rs:156-190) through the oracle pairing."""
    n = 1 << log_n
    rng = random.Random(log_n)
    m, n_input, rows = zg.horner_qap_rows(n)
    x, cs = rand_fr(rng, True), [rand_fr(rng) for _ in range(n)]
    wit = zg.horner_witness(n, x, cs)
    toxic = tuple(rand_fr(rng, True) for _ in range(5))
    r, s = rand_fr(rng, True), rand_fr(rng, True)
    q = zk.QAP(ctx, n, m, n_input, rows)
    crs = zk.setup(ctx, q, toxic)
    got = zk.prove(ctx, q, crs, wit, r, s)
    ru, rv, rw = _rows_from_csr(rows)
    want = cf.expected_proof(n, synthetic.omega(log_n), ru, rv, rw, n_input, wit, toxic, r, s)
    assert (got.a, got.b, got.c) == want
    # groth16::verify on the device at every size (the CRS stays resident; O(1) pairings)
    wrong = [wit[1], (wit[2] + 1) % P]  # wrong public output -> reject (lib.rs:182-189)
    assert zk.verify(ctx, crs, wit[1:3], got)
    assert not zk.verify(ctx, crs, wrong, got)
    if log_n == 6:
        d = crs.download()
        s1 = og.SigmaG1(d["alpha1"], d["beta1"], d["delta1"], d["xi1"], d["sum_gamma"], d["sum_delta"], d["xi_t"])
        s2 = og.SigmaG2(d["beta2"], d["gamma2"], d["delta2"], d["xi2"])
        assert og.verify(og.BN254Backend(), (s1, s2), wit[1:3], og.Proof(got.a, got.b, got.c))
        bad = list(wit)
        bad[2] = (bad[2] + 1) % P  # wrong public output -> reject (lib.rs:182-189)
        assert not og.verify(og.BN254Backend(), (s1, s2), bad[1:3], og.Proof(got.a, got.b, got.c))


def test_prove_sharded_equals_single(ctx):
    """Multi-GPU decomposition on one device: world-2 and world-3 shards, partial sums folded by
    zkb_prove_combine, must equal the unsharded proof bit for bit."""
    n = 64
    rng = random.Random(77)
    m, n_input, rows = zg.horner_qap_rows(n)
    wit = zg.horner_witness(n, rand_fr(rng, True), [rand_fr(rng) for _ in range(n)])
    toxic = tuple(rand_fr(rng, True) for _ in range(5))
    r, s = rand_fr(rng, True), rand_fr(rng, True)
    q = zk.QAP(ctx, n, m, n_input, rows)
    full = zk.prove(ctx, q, zk.setup(ctx, q, toxic), wit, r, s)
    for world in (2, 3):
        parts = []
        shards = [zk.setup(ctx, q, toxic, rank=k, world=world) for k in range(world)]
        for k in range(world):
            parts.append(zk.prove_partial(ctx, q, shards[k], wit, r, s))
        got = zk.prove_combine(ctx, np.stack(parts))
        assert (got.a, got.b, got.c) == (full.a, full.b, full.c)


# ------------------------------------------------------------------------------------------------
# generic root domain: the reference's own circuits (ASTParser numbers the gates 1..=n, circuit/mod.rs:517)
from test_oracle_kats import QUAD, SIMPLE  # noqa: E402

MIXED = """(in x a b)
(out y z)
(verify x y z)

(program
    (= t1 (* x x))
    (= t2 (* (+ t1 a) (+ x b 7)))
    (= y (* 1 (+ t2 t1 3)))
    (= z (* t2 (+ y x))))"""


def _parser_case(text, n_inputs, seed):
    rng = random.Random(seed)
    rep = circuit.try_parse(FR, text)
    wit = circuit.weights(FR, text, [rand_fr(rng, True) for _ in range(n_inputs)])
    toxic = tuple(rand_fr(rng, True) for _ in range(5))
    return rep, wit, toxic, rand_fr(rng, True), rand_fr(rng, True)


@pytest.mark.parametrize("name,text,n_inputs", [("simple", SIMPLE, 3), ("quad", QUAD, 4), ("mixed", MIXED, 3),
                                                ("deg_15", synthetic.horner_program_text(16), 17)])
@pytest.mark.parametrize("valid", [True, False])
def test_parser_circuits_generic_domain(ctx, name, text, n_inputs, valid):
    """Circuits as the reference's parser emits them (roots 1..=n: simple.zk of lib.rs:156-190, the quad
    share of fr.rs:273-302, deg_15 of fr.rs:361-416, a mixed one) through the device path: h, the
    device CRS and the proof are bit-exact vs the literal restatement, and the proof verifies."""
    rep, wit, toxic, r, s = _parser_case(text, n_inputs, sum(map(ord, name)))
    if not valid:
        wit = list(wit)
        wit[-1] = (wit[-1] + 1) % P
    assert rep.roots == list(range(1, len(rep.roots) + 1))
    dense = og.qap_from_root_rep(FR, rep)
    B = og.BN254Backend()
    sig = og.setup(B, dense, toxic)
    want = og.prove(B, dense, sig, wit, r, s)
    q = zk.QAP.from_root_representation(ctx, rep)
    # h(x) and the weighted sums against the reference's Mul / Div on coefficient vectors
    u_sum, v_sum, w_sum = og.weighted_sums(FR, dense, wit)
    h = og.quotient_h(FR, dense, u_sum, v_sum, w_sum)
    gu, gv, gh = zg.qap_h(ctx, q, wit)
    n = len(rep.roots)
    pad = lambda v: (list(v) + [0] * n)[:n]
    assert gu == pad(u_sum) and gv == pad(v_sum) and gh == pad(h)[: n - 1]
    # device setup == reference setup (same toxic waste), then prove over both CRS routes
    crs = zk.setup(ctx, q, toxic)
    d = crs.download()
    assert (d["alpha1"], d["beta1"], d["delta1"]) == (sig[0].alpha, sig[0].beta, sig[0].delta)
    assert d["xi1"] == sig[0].xi and d["xi_t"] == sig[0].xi_t
    assert d["sum_gamma"] == sig[0].sum_gamma and d["sum_delta"] == sig[0].sum_delta
    assert (d["beta2"], d["gamma2"], d["delta2"], d["xi2"]) == (sig[1].beta, sig[1].gamma, sig[1].delta, sig[1].xi)
    got = zk.prove(ctx, q, crs, wit, r, s)
    assert (got.a, got.b, got.c) == (want.a, want.b, want.c)
    got2 = zk.prove(ctx, q, zk.CRS.upload(ctx, sig[0], sig[1]), wit, r, s)
    assert (got2.a, got2.b, got2.c) == (want.a, want.b, want.c)
    n_pub = rep.input
    assert og.verify(B, sig, wit[1:1 + n_pub], og.Proof(got.a, got.b, got.c)) == valid
    assert zk.verify(ctx, crs, wit[1:1 + n_pub], got) == valid


def test_generic_domain_random_roots_and_limits(ctx):
    """Arbitrary pairwise-distinct roots (not 1..n, n not a power of two); repeated roots are refused
    (the reference's lagrange_basis would divide by zero); n > 32768 on an explicit domain is refused."""
    rng = random.Random(33)
    n = 11
    roots = rng.sample(range(2, 10 ** 6), n)
    rep = synthetic.horner_rep(FR, n, roots)
    wit = synthetic.horner_witness(FR, n, rand_fr(rng, True), [rand_fr(rng) for _ in range(n)])
    toxic = tuple(rand_fr(rng, True) for _ in range(5))
    r, s = rand_fr(rng, True), rand_fr(rng, True)
    dense = og.qap_from_root_rep(FR, rep)
    B = og.BN254Backend()
    sig = og.setup(B, dense, toxic)
    want = og.prove(B, dense, sig, wit, r, s)
    q = zk.QAP.from_root_representation(ctx, rep)
    got = zk.prove(ctx, q, zk.setup(ctx, q, toxic), wit, r, s)
    assert (got.a, got.b, got.c) == (want.a, want.b, want.c)
    bad = synthetic.horner_rep(FR, 4, [5, 6, 5, 7])
    with pytest.raises(zk.ZkbError):
        zk.QAP.from_root_representation(ctx, bad)
    m, n_input, rows = zg.horner_qap_rows(32769)
    with pytest.raises(zk.ZkbError):
        zk.QAP(ctx, 32769, m, n_input, rows, roots=list(range(1, 32770)))


def test_generic_domain_6000_gates_parser_numbering(ctx):
    """Past toy sizes on the parser's own domain (roots 1..=n, circuit/mod.rs:517; n = 6000, not a power of two): the dense
    device path against (1) the closed-form proof from the toxic waste, (2) u_sum / v_sum interpolating the gate
    evaluations at sampled roots (the defining property of mod.rs:233-246 + coefficient_poly.rs:159-190), (3) the division
    identity u(z) v(z) - w(z) = h(z) t(z) at a random point (field/mod.rs:428-469, remainder 0 for a satisfying witness),
    (4) verify == true / false on the device."""
    n = 6000
    rng = random.Random(6000)
    m, n_input, rows = zg.horner_qap_rows(n)
    wit = zg.horner_witness(n, rand_fr(rng, True), [rand_fr(rng) for _ in range(n)])
    toxic = tuple(rand_fr(rng, True) for _ in range(5))
    r, s = rand_fr(rng, True), rand_fr(rng, True)
    q = zk.QAP(ctx, n, m, n_input, rows, roots=list(range(1, n + 1)))
    crs = zk.setup(ctx, q, toxic)
    got = zk.prove(ctx, q, crs, wit, r, s)
    ru, rv, rw = _rows_from_csr(rows)
    assert (got.a, got.b, got.c) == cf.expected_proof(n, None, ru, rv, rw, n_input, wit, toxic, r, s)
    assert zk.verify(ctx, crs, wit[1:3], got) and not zk.verify(ctx, crs, [wit[1], (wit[2] + 1) % P], got)
    u, v, h = zk.qap_h(ctx, q, wit)

    def gate_evals(rows_):
        out = [0] * n
        for i, row in enumerate(rows_):
            for g, c in row:
                out[g] = (out[g] + c * wit[i]) % P
        return out
    A, Bv, Cw = gate_evals(ru), gate_evals(rv), gate_evals(rw)

    def horner(cs, x):
        acc = 0
        for c in reversed(cs):
            acc = (acc * x + c) % P
        return acc
    for k in [0, 1, n - 1] + rng.sample(range(n), 29):
        assert horner(u, k + 1) == A[k] and horner(v, k + 1) == Bv[k]
    z = rand_fr(rng, True)
    L, tz = cf.lagrange_at_1_to_n(n, z)
    wz = sum(c * l for c, l in zip(Cw, L)) % P
    assert (horner(u, z) * horner(v, z) - wz) % P == horner(h, z) * tz % P
    assert len(h) == n - 1


def test_reindexed_parser_circuit_on_the_fast_domain(ctx):
    """QAP.from_root_representation(rep, reindex=True): a parser-numbered circuit moved onto the roots of unity (next power
    of two, empty gates appended) proves and verifies under the CRS of that QAP; accepted / rejected as the reference's
    acceptance tests do (lib.rs:156-190)."""
    rng = random.Random(5)
    rep = circuit.try_parse(FR, MIXED)
    wit = circuit.weights(FR, MIXED, [rand_fr(rng, True) for _ in range(3)])
    q = zk.QAP.from_root_representation(ctx, rep, reindex=True)
    assert q.n == 4 and len(rep.roots) == 4
    rep5 = synthetic.horner_rep(FR, 5, [1, 2, 3, 4, 5])
    q5 = zk.QAP.from_root_representation(ctx, rep5, reindex=True)
    assert q5.n == 8
    for qq, ww, npub in ((q, wit, rep.input), (q5, synthetic.horner_witness(FR, 5, 3, [4, 5, 6, 7, 8]), 2)):
        crs = zk.setup(ctx, qq, tuple(rand_fr(rng, True) for _ in range(5)))
        pr = zk.prove(ctx, qq, crs, ww)
        assert zk.verify(ctx, crs, ww[1:1 + npub], pr)
        bad = list(ww[1:1 + npub])
        bad[-1] = (bad[-1] + 1) % P
        assert not zk.verify(ctx, crs, bad, pr)


def test_prove_batch_sharded_equals_single(ctx):
    """The N>1 bench path on one device: zkb_prove_batch over each rank's CRS shard (partial records),
    gathered to (world, count, 32) and folded by one zkb_prove_combine_batch launch."""
    n = 128
    rng = random.Random(78)
    m, n_input, rows = zg.horner_qap_rows(n)
    q = zk.QAP(ctx, n, m, n_input, rows)
    toxic = tuple(rand_fr(rng, True) for _ in range(5))
    wits = [zg.horner_witness(n, rand_fr(rng, True), [rand_fr(rng) for _ in range(n)]) for _ in range(3)]
    rs = [rand_fr(rng, True) for _ in range(3)]
    ss = [rand_fr(rng, True) for _ in range(3)]
    full_crs = zk.setup(ctx, q, toxic)
    want = [zk.prove(ctx, q, full_crs, w, r, s) for w, r, s in zip(wits, rs, ss)]
    for world in (2, 4):
        parts = []
        for k in range(world):
            shard = zk.setup(ctx, q, toxic, rank=k, world=world)
            rec = zk.prove_batch(ctx, q, shard, wits, rs, ss)
            assert rec.shape == (3, 32)
            parts.append(rec)
        got = zk.prove_combine_batch(ctx, np.stack(parts))
        assert [(p.a, p.b, p.c) for p in got] == [(p.a, p.b, p.c) for p in want]


def test_prove_batch_equals_single(ctx):
    """zkb_prove_batch (two proofs in flight on two stream pairs) == independent zkb_prove calls,
    for 1..5 proofs with different witnesses and (r, s); host and device-resident weights."""
    n = 256
    rng = random.Random(91)
    m, n_input, rows = zg.horner_qap_rows(n)
    q = zk.QAP(ctx, n, m, n_input, rows)
    crs = zk.setup(ctx, q, tuple(rand_fr(rng, True) for _ in range(5)))
    wits = [zg.horner_witness(n, rand_fr(rng, True), [rand_fr(rng) for _ in range(n)]) for _ in range(5)]
    wits[3][5] = (wits[3][5] + 1) % P  # one invalid witness in the middle
    rs = [rand_fr(rng, True) for _ in range(5)]
    ss = [rand_fr(rng, True) for _ in range(5)]
    single = [zk.prove(ctx, q, crs, w, r, s) for w, r, s in zip(wits, rs, ss)]
    for k in (1, 2, 3, 5):
        got = zk.prove_batch(ctx, q, crs, wits[:k], rs[:k], ss[:k])
        assert [(p.a, p.b, p.c) for p in got] == [(p.a, p.b, p.c) for p in single[:k]]
    dptrs = []
    for w in wits:
        a = zg.fr_limbs(w)
        d = ctx.dev_alloc(a.nbytes)
        ctx.h2d(d, a)
        dptrs.append(d)
    got = zk.prove_batch(ctx, q, crs, dptrs, rs, ss, on_device=True)
    assert [(p.a, p.b, p.c) for p in got] == [(p.a, p.b, p.c) for p in single]
    for d in dptrs:
        ctx.dev_free(d)
    assert zk.prove_batch(ctx, q, crs, [], [], []) == []


def test_error_paths(ctx):
    with pytest.raises(zk.ZkbError):
        zk.QAP(ctx, 6, 4, 1, [(np.zeros(5, dtype=np.uint64), np.zeros(0, dtype=np.uint32), np.zeros((0, 4), dtype=np.uint64))] * 3)
    n = 4
    m, n_input, rows = zg.horner_qap_rows(n)
    shifted = [(ptr + np.uint64(1), gate, coeff) for ptr, gate, coeff in rows]
    with pytest.raises(zk.ZkbError):
        zk.QAP(ctx, n, m, n_input, shifted)  # row offsets that do not start at 0
    q = zk.QAP(ctx, n, m, n_input, rows)
    with pytest.raises(zk.ZkbError):
        zk.setup(ctx, q, (1, 2, 0, 4, 5))  # zero secret: the reference's random_elem never yields 0
    q8 = zk.QAP.horner(ctx, 8)
    crs4 = zk.setup(ctx, q, (1, 2, 3, 4, 5))
    with pytest.raises(zk.ZkbError):
        zk.prove(ctx, q8, crs4, [1] * 18, 1, 1)  # CRS of another QAP


# ------------------------------------------------------------------------------------------------
# pairing / verify  (fr.rs:120-122, 225-231; groth16/mod.rs:299-320)
def _gt_to_flat(c):
    """Device GT layout ((c0, c1) of the w^i coefficient, i = 0..5; u = w^6 - 9) -> the oracle's dense w-polynomial."""
    f = [0] * 12
    for i in range(6):
        f[i] = (f[i] + c[2 * i] - 9 * c[2 * i + 1]) % bn.Q
        f[i + 6] = (f[i + 6] + c[2 * i + 1]) % bn.Q
    return bn.Fq12(f)


def test_pairing_matches_oracle(ctx):
    rng = random.Random(77)
    a, b, c = (rand_fr(rng, True) for _ in range(3))
    Pa, Qb = bn.g1_mul(bn.BASE_G1, a), bn.g2_mul(bn.BASE_G2, b)
    assert _gt_to_flat(zk.pairing(ctx, [(Pa, Qb)])) == bn.pairing(Pa, Qb)
    # GT "+" (fr.rs:225-231) is the Fq12 product: e(aG, bH) e(cG, H) == e((ab + c)G, H)
    lhs = _gt_to_flat(zk.pairing(ctx, [(Pa, Qb), (bn.g1_mul(bn.BASE_G1, c), bn.BASE_G2)]))
    rhs = _gt_to_flat(zk.pairing(ctx, [(bn.g1_mul(bn.BASE_G1, (a * b + c) % P), bn.BASE_G2)]))
    assert lhs == rhs and lhs != bn.Fq12.one()
    # identity in either slot, the empty product, P with -P
    one = bn.Fq12.one()
    assert _gt_to_flat(zk.pairing(ctx, [(None, Qb)])) == one
    assert _gt_to_flat(zk.pairing(ctx, [(Pa, None)])) == one
    assert _gt_to_flat(zk.pairing(ctx, [])) == one
    assert _gt_to_flat(zk.pairing(ctx, [(Pa, Qb), (bn.g1_neg(Pa), Qb)])) == one
    # a point off the curve is an argument error, not a silent value
    with pytest.raises(zk.ZkbError):
        zk.pairing(ctx, [((Pa[0], (Pa[1] + 1) % bn.Q), Qb)])


def test_verify_batch_and_edge_cases(ctx):
    """verify over a batch: valid proofs, a wrong public input, a tampered proof, an off-curve point, identity points;
    zip truncation of the inputs (mod.rs:312-316) -- every verdict equals the oracle's verify()."""
    n = 8
    rep, wit, toxic, r, s = _horner_case(n, 404)
    B = og.BN254Backend()
    dense = og.qap_from_root_rep(FR, rep)
    sig = og.setup(B, dense, toxic)
    q = zk.QAP.from_root_representation(ctx, rep)
    crs = zk.CRS.upload(ctx, sig[0], sig[1])
    good = zk.prove(ctx, q, crs, wit, r, s)
    good2 = zk.prove(ctx, q, crs, wit, r + 1, s + 5)
    pub = wit[1:rep.input + 1]
    tampered = zg.Proof(good.a, good.b, bn.g1_add(good.c, bn.BASE_G1))
    offcurve = zg.Proof((good.a[0], (good.a[1] + 1) % bn.Q), good.b, good.c)
    ident = zg.Proof(None, good.b, good.c)
    proofs = [good, good2, good, tampered, offcurve, ident]
    inputs = [pub, pub, [pub[0], (pub[1] + 1) % P], pub, pub, pub]
    got = zg.verify_batch(ctx, crs, inputs, proofs)
    want = [og.verify(B, sig, i, og.Proof(p.a, p.b, p.c)) if p is not offcurve else False for i, p in zip(inputs, proofs)]
    assert want == [True, True, False, False, False, False]
    assert got == want
    # zip truncation: extra inputs are ignored, missing ones end the sum early (and change the verdict)
    assert zk.verify(ctx, crs, pub + [5, 6], good) == og.verify(B, sig, pub + [5, 6], og.Proof(good.a, good.b, good.c)) is True
    assert zk.verify(ctx, crs, pub[:1], good) == og.verify(B, sig, pub[:1], og.Proof(good.a, good.b, good.c))
    assert zk.verify(ctx, crs, [], good) == og.verify(B, sig, [], og.Proof(good.a, good.b, good.c))
    # a sharded CRS verifies on every rank (fixed points and sum_gamma are replicated)
    crs1 = zk.CRS.upload(ctx, sig[0], sig[1], rank=1, world=2)
    assert zk.verify(ctx, crs1, pub, good) and not zk.verify(ctx, crs1, pub, tampered)


# ------------------------------------------------------------------------------------------------
# raw coordinates over the ABI: curve and subgroup membership (the reference can only hold bn-constructed group elements)
def _f2_pow(a, e):
    r = bn.F2_ONE
    while e:
        if e & 1:
            r = bn.f2_mul(r, a)
        a = bn.f2_mul(a, a)
        e >>= 1
    return r


def _f2_sqrt(a):
    """Square root in Fq2 = Fq[u]/(u^2 + 1), q = 3 mod 4 (complex method); None if a is not a square."""
    q = bn.Q
    a1 = _f2_pow(a, (q - 3) // 4)
    alpha = bn.f2_mul(a1, bn.f2_mul(a1, a))
    x0 = bn.f2_mul(a1, a)
    if alpha == ((q - 1) % q, 0):
        x = bn.f2_mul((0, 1), x0)
    else:
        b = _f2_pow(bn.f2_add(bn.F2_ONE, alpha), (q - 1) // 2)
        x = bn.f2_mul(b, x0)
    return x if bn.f2_mul(x, x) == (a[0] % q, a[1] % q) else None


def _twist_point_outside_g2(seed):
    """A point of the twist E'(Fq2) that is NOT in the order-r subgroup (the cofactor is ~2^254, so a random twist point
    almost never is): on the curve, but [r]P != O."""
    rng = random.Random(seed)
    while True:
        x = (rng.randrange(bn.Q), rng.randrange(bn.Q))
        y = _f2_sqrt(bn.f2_add(bn.f2_mul(x, bn.f2_mul(x, x)), bn.G2_B))
        if y is None:
            continue
        Pt = (x, y)
        assert bn.g2_is_on_curve(Pt)
        if bn.g2_add(bn.g2_mul(Pt, bn.R_ORDER - 1), Pt) is not None:  # [r]P != O (g2_mul reduces its scalar mod r)
            return Pt


def test_g2_points_outside_the_subgroup_are_rejected(ctx):
    """zkb_pairing / zkb_verify / zkb_crs_upload / zkb_bases_upload take raw coordinates; a twist point outside the
    order-r subgroup (on the curve!) must be refused, as EIP-197 demands of alt_bn128 verifiers."""
    bad = _twist_point_outside_g2(5)
    Pa = bn.g1_mul(bn.BASE_G1, 7)
    with pytest.raises(zk.ZkbError, match="subgroup"):
        zk.pairing(ctx, [(Pa, bad)])
    with pytest.raises(zk.ZkbError, match="subgroup"):
        zk.pairing(ctx, [(Pa, bn.BASE_G2), (Pa, bad)])
    with pytest.raises(zk.ZkbError, match="subgroup"):
        zk.Bases.upload(ctx, 2, [bn.BASE_G2, bad])
    with pytest.raises(zk.ZkbError, match="not on its curve"):
        zk.Bases.upload(ctx, 1, [bn.BASE_G1, (5, 7)])
    assert zk.Bases.upload(ctx, 2, [bn.BASE_G2, None, bn.g2_mul(bn.BASE_G2, 5)]).download()[2] == bn.g2_mul(bn.BASE_G2, 5)
    n = 4
    rep, wit, toxic, r, s = _horner_case(n, 606)
    B = og.BN254Backend()
    dense = og.qap_from_root_rep(FR, rep)
    s1, s2 = og.setup(B, dense, toxic)
    q = zk.QAP.from_root_representation(ctx, rep)
    crs = zk.CRS.upload(ctx, s1, s2)
    good = zk.prove(ctx, q, crs, wit, r, s)
    pub = wit[1:rep.input + 1]
    assert zk.verify(ctx, crs, pub, good)
    assert not zk.verify(ctx, crs, pub, zg.Proof(good.a, bad, good.c))
    got = zg.verify_batch(ctx, crs, [pub, pub, pub], [good, zg.Proof(good.a, bad, good.c), good])
    assert got == [True, False, True]
    import copy
    for field, val in (("xi", bad), ("beta", bad), ("gamma", bad), ("delta", bad)):
        t2 = copy.deepcopy(s2)
        if field == "xi":
            t2.xi[1] = val
        else:
            setattr(t2, field, val)
        with pytest.raises(zk.ZkbError, match="subgroup"):
            zk.CRS.upload(ctx, s1, t2)
    t1 = copy.deepcopy(s1)
    t1.sum_delta[0] = (t1.sum_delta[0][0], (t1.sum_delta[0][1] + 1) % bn.Q)
    with pytest.raises(zk.ZkbError, match="not on its curve"):
        zk.CRS.upload(ctx, t1, s2)
    # a CRS whose vectors do not fit the QAP (check_pair): one sum_delta entry short
    t1 = copy.deepcopy(s1)
    t1.sum_delta = t1.sum_delta[:-1]
    with pytest.raises(zk.ZkbError, match="does not belong"):
        zk.prove(ctx, q, zk.CRS.upload(ctx, t1, s2), wit, r, s)
