"""GPU parity of the batched-affine pair tree in front of the XYZZ bucket chain (csrc/affine_level.cuh, k_affine_level,
k_accumulate_points; switched per group with ZKB_AFF_G1 / ZKB_AFF_G2 = levels; off by default: measured not faster).
Group addition is exact, so every MSM and every proof must be bit-identical to the oracle and to the chain-only path,
whatever the number of levels: the adversarial bases of test_msm_small_adversarial (identity, P / -P, repeated points,
one bucket taking every record), the collapse property at sizes where several work items and levels run, skewed scalars,
and whole proofs (single, batch) compared with the closed form from the toxic waste."""

import importlib
import random

import numpy as np
import pytest

from oracle import bn254 as bn
from oracle import closed_form as cf
from oracle import synthetic
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


@pytest.fixture(params=[0, 1], ids=["plain", "staged"])
def var(request):
    """Level body: 0 = the host/device body the CPU test runs (affine_level.cuh), 1 = the staged device body (operands through
    cp.async and shared memory, msm_impl.cuh: affine_level_item_staged)."""
    return request.param


def set_mode(mp, g1, g2, var=0):
    mp.setenv("ZKB_AFF_G1", str(g1))
    mp.setenv("ZKB_AFF_G2", str(g2))
    mp.setenv("ZKB_AFF_VAR", str(var))


@pytest.mark.parametrize("group", [1, 2])
@pytest.mark.parametrize("levels", [1, 2, 3, 8])
def test_msm_adversarial_with_pair_tree(ctx, monkeypatch, var, group, levels):
    set_mode(monkeypatch, levels, levels, var)
    rng = random.Random(140 + group)
    base = bn.BASE_G1 if group == 1 else bn.BASE_G2
    mul = bn.g1_mul if group == 1 else bn.g2_mul
    neg = bn.g1_neg if group == 1 else bn.g2_neg
    ora = bn.msm_g1 if group == 1 else bn.msm_g2
    Pt = mul(base, 12345)
    pts = [Pt, Pt, neg(Pt), None, mul(base, 7), mul(base, 7), base, mul(base, P - 1)]
    pts += [mul(base, rand_fr(rng)) for _ in range(8)]
    scal = [5, 5, 10, 999, 0, 1, P - 1, P - 1] + [rand_fr(rng) for _ in range(8)]
    b = zk.Bases.upload(ctx, group, pts)
    for c in (0, 2, 3, 5, 8):
        assert zk.msm(ctx, b, scal, window_bits=c) == ora(scal, pts), f"window {c}"
    assert zk.msm(ctx, b, [3, 4, 7] + [0] * 13) is None              # everything cancels
    assert zk.msm(ctx, b, scal[:3]) == ora(scal[:3], pts[:3])
    assert zk.msm(ctx, b, []) is None
    assert zk.msm(ctx, b, [0] * 16) is None                          # no record at all
    assert zk.msm(ctx, b, [scal[9]] * 16, window_bits=3) == ora([scal[9]] * 16, pts)  # one bucket per window takes all
    # the same point 16 times with the same scalar: every pair of every level is a tangent
    b2 = zk.Bases.upload(ctx, group, [Pt] * 16)
    assert zk.msm(ctx, b2, [scal[10]] * 16, window_bits=4) == mul(Pt, 16 * scal[10] % P)
    # P and -P alternating with equal scalars: every pair of the first level cancels
    b3 = zk.Bases.upload(ctx, group, [Pt, neg(Pt)] * 8)
    assert zk.msm(ctx, b3, [scal[11]] * 16, window_bits=4) is None


@pytest.mark.parametrize("group,log_n,levels", [(1, 8, 2), (1, 12, 3), (1, 12, 5), (2, 10, 4), (2, 10, 6), (1, 16, 4), (2, 14, 5), (1, 16, 8)])
def test_msm_collapse_with_pair_tree(ctx, monkeypatch, var, group, log_n, levels):
    """bases P_i = k_i * BASE  =>  sum s_i P_i = (sum s_i k_i) * BASE, with witness-like skew (zeros and ones: one huge
    bucket) among random scalars; the chain-only result must be the same point."""
    n = 1 << log_n
    rng = random.Random(log_n * 10 + group)
    ks = [rand_fr(rng) for _ in range(n)]
    ss = [rand_fr(rng) for _ in range(n)]
    for i in range(0, n, 5):
        ss[i] = i % 2
    b = zk.Bases.generate(ctx, group, ks)
    e = sum(s * k for s, k in zip(ss, ks)) % P
    want = bn.g1_mul(bn.BASE_G1, e) if group == 1 else bn.g2_mul(bn.BASE_G2, e)
    set_mode(monkeypatch, 0, 0)
    assert zk.msm(ctx, b, ss) == want
    set_mode(monkeypatch, levels, levels, var)
    assert zk.msm(ctx, b, ss) == want
    # small scalars only: few buckets hold everything (deep trees, long pass-through tails)
    small = [rng.randrange(0, 9) for _ in range(n)]
    e = sum(s * k for s, k in zip(small, ks)) % P
    want = bn.g1_mul(bn.BASE_G1, e) if group == 1 else bn.g2_mul(bn.BASE_G2, e)
    assert zk.msm(ctx, b, small) == want


def test_msm_2pow20_with_pair_tree(ctx, monkeypatch, var):
    n = 1 << 20
    rng = np.random.default_rng(20)
    k = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64)
    s = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64)
    k[:, 3] &= np.uint64((1 << 60) - 1)
    s[:, 3] &= np.uint64((1 << 60) - 1)
    ki, si = zg.limbs_to_ints(k), zg.limbs_to_ints(s)
    e = sum(x * y for x, y in zip(ki, si)) % P
    for group, levels in ((1, 4), (2, 5)):
        set_mode(monkeypatch, levels, levels, var)
        b = zk.Bases.generate(ctx, group, k)
        want = bn.g1_mul(bn.BASE_G1, e) if group == 1 else bn.g2_mul(bn.BASE_G2, e)
        assert zk.msm(ctx, b, s) == want
        b.free()


def _rows_from_csr(rows):
    out = []
    for ptr, gate, coeff in rows:
        cs = zg.limbs_to_ints(coeff)
        out.append([[(int(gate[e]), cs[e]) for e in range(int(ptr[i]), int(ptr[i + 1]))] for i in range(len(ptr) - 1)])
    return out


@pytest.mark.parametrize("log_n,g1,g2", [(6, 2, 2), (10, 3, 3), (12, 0, 4), (12, 4, 0), (16, 4, 5)])
def test_prove_with_pair_tree(ctx, monkeypatch, var, log_n, g1, g2):
    """Setup + prove on the synthetic Horner QAP against the closed-form proof from the toxic waste (bit-exact), single
    proof and batch, valid and invalid witness (the invalid one against the chain-only proof)."""
    n = 1 << log_n
    rng = random.Random(300 + log_n)
    m, n_input, rows = zg.horner_qap_rows(n)
    x, cs = rand_fr(rng, True), [rand_fr(rng) for _ in range(n)]
    wit = zg.horner_witness(n, x, cs)
    bad = list(wit)
    bad[7] = (bad[7] + 1) % P
    toxic = tuple(rand_fr(rng, True) for _ in range(5))
    r, s = rand_fr(rng, True), rand_fr(rng, True)
    q = zk.QAP(ctx, n, m, n_input, rows)
    set_mode(monkeypatch, 0, 0)
    crs = zk.setup(ctx, q, toxic)
    plain_bad = zk.prove(ctx, q, crs, bad, r, s)
    set_mode(monkeypatch, g1, g2, var)
    got = zk.prove(ctx, q, crs, wit, r, s)
    if log_n <= 12:
        ru, rv, rw = _rows_from_csr(rows)
        want = cf.expected_proof(n, synthetic.omega(log_n), ru, rv, rw, n_input, wit, toxic, r, s)
        assert (got.a, got.b, got.c) == want
    assert zk.verify(ctx, crs, wit[1:3], got)
    got_bad = zk.prove(ctx, q, crs, bad, r, s)
    assert (got_bad.a, got_bad.b, got_bad.c) == (plain_bad.a, plain_bad.b, plain_bad.c)
    batch_out = zk.prove_batch(ctx, q, crs, [wit, bad, wit], [r, r, r], [s, s, s])
    assert [(p.a, p.b, p.c) for p in batch_out] == [(got.a, got.b, got.c), (got_bad.a, got_bad.b, got_bad.c), (got.a, got.b, got.c)]
