"""CPU check of the batched-affine pair tree (csrc/affine_level.cuh, the body of k_affine_level) compiled with g++ by
tests/hostcheck: every level's bucket sums against Oracle A's group law, on bucket shapes that exercise the pass-through
of odd elements, empty buckets, items that straddle buckets, and every special case of the affine addition inside a
batch (identity operands, P + P, P + (-P), repeated entries) -- the cases the reference's `bn` addition handles
(fr.rs:175-223)."""

import ctypes
import os
import random
import subprocess

import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from oracle import bn254 as bn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def hc():
    src = os.path.join(HERE, "hostcheck", "hostcheck.cpp")
    so = os.path.join(HERE, "hostcheck", "_hostcheck.so")
    deps = [src] + [os.path.join(ROOT, "zksnark-rs_b200", "csrc", f)
                    for f in ("ff.cuh", "ec.cuh", "pairing.cuh", "constants.h", "affine_level.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-Wno-unknown-pragmas",
                               "-I", os.path.join(ROOT, "zksnark-rs_b200", "csrc"), "-o", so, src])
    lib = ctypes.CDLL(so)
    lib.hc_pair_tree.restype = ctypes.c_uint32
    return lib


def limbs(vals):
    out = (ctypes.c_uint64 * (4 * len(vals)))()
    for i, v in enumerate(vals):
        for j in range(4):
            out[4 * i + j] = (v >> (64 * j)) & 0xFFFFFFFFFFFFFFFF
    return out


def unlimbs(buf, n):
    return [sum(buf[4 * i + j] << (64 * j) for j in range(4)) for i in range(n)]


class G1:
    group, width = 1, 2
    add, neg = staticmethod(bn.g1_add), staticmethod(bn.g1_neg)

    @staticmethod
    def points(rng, n):
        return [bn.g1_mul(bn.BASE_G1, rng.randrange(1, bn.R_ORDER)) for _ in range(n)]

    @staticmethod
    def flat(P):
        return [0, 0] if P is None else [P[0], P[1]]

    @staticmethod
    def unflat(v):
        return None if v == [0, 0] else (v[0], v[1])


class G2:
    group, width = 2, 4
    add, neg = staticmethod(bn.g2_add), staticmethod(bn.g2_neg)

    @staticmethod
    def points(rng, n):
        return [bn.g2_mul(bn.G2_GEN, rng.randrange(1, bn.R_ORDER)) for _ in range(n)]

    @staticmethod
    def flat(P):
        return [0, 0, 0, 0] if P is None else [P[0][0], P[0][1], P[1][0], P[1][1]]

    @staticmethod
    def unflat(v):
        return None if v == [0, 0, 0, 0] else ((v[0], v[1]), (v[2], v[3]))


def run_tree(hc, G, table, buckets, levels, batch):
    """buckets: list of lists of (entry, sign).  Returns the per-bucket element lists of the last level."""
    flat = []
    for P in table:
        flat += G.flat(P)
    offs, recs = [0], []
    for b in buckets:
        recs += [e | (s << 31) for e, s in b]
        offs.append(len(recs))
    nbk = len(buckets)
    offs0 = (ctypes.c_uint32 * (nbk + 1))(*offs)
    sorted_ = (ctypes.c_uint32 * max(1, len(recs)))(*recs)
    offs_out = (ctypes.c_uint32 * (nbk + 1))()
    elems = (ctypes.c_uint64 * (4 * G.width * max(1, len(recs))))()
    n = hc.hc_pair_tree(G.group, batch, len(table), limbs(flat), nbk, offs0, sorted_, levels, offs_out, elems)
    vals = unlimbs(elems, n * G.width)
    pts = [G.unflat(vals[G.width * i:G.width * (i + 1)]) for i in range(n)]
    oo = list(offs_out)
    assert oo[nbk] == n
    return [pts[oo[b]:oo[b + 1]] for b in range(nbk)]


def bucket_sum(G, table, bucket):
    acc = None
    for e, s in bucket:
        P = table[e]
        acc = G.add(acc, G.neg(P) if (s and P is not None) else P)
    return acc


def fold(G, pts):
    acc = None
    for P in pts:
        acc = G.add(acc, P)
    return acc


def shapes(rng, ntab, identity_at):
    """Bucket contents covering the edge cases; entry `identity_at` of the table is the identity."""
    r = lambda: (rng.randrange(1, ntab), rng.randrange(2))  # noqa: E731  (never the identity entry by accident: it is entry 0)
    b = []
    b.append([])                                            # empty bucket first
    b.append([r()])                                         # single element: passes through every level
    b.append([(3, 0), (3, 0)])                              # P + P
    b.append([(4, 0), (4, 1)])                              # P - P
    b.append([(5, 1), (5, 1), (5, 1), (5, 1)])              # 4 (-P): doubling at two levels
    b.append([(identity_at, 0), r()])                       # O + Q
    b.append([r(), (identity_at, 1)])                       # P + O
    b.append([(identity_at, 0), (identity_at, 0)])          # O + O
    b.append([(6, 0), (6, 1), r()])                         # (P - P) then + Q at the next level
    b.append([(7, 0), (7, 1), (8, 0), (8, 1)])              # O + O at level 2
    b.append([])
    b.append([])
    for k in (2, 3, 5, 7, 8, 9, 16, 17, 31, 33, 40):        # sizes around the powers of two and the batch sizes
        b.append([r() for _ in range(k)])
    b.append([(9, 0)] * 13)                                 # 13 P: tangents and chords mixed over the levels
    b.append([])
    return b


@pytest.mark.parametrize("G", [G1, G2], ids=["g1", "g2"])
@pytest.mark.parametrize("batch", [4, 16, 32])
def test_pair_tree_levels(hc, G, batch):
    rng = random.Random(1000 + batch + G.group)
    ntab = 24 if G is G1 else 12
    table = [None] + G.points(rng, ntab - 1)
    buckets = shapes(rng, ntab, 0)
    want = [bucket_sum(G, table, b) for b in buckets]
    for levels in (1, 2, 3, 6):
        got = run_tree(hc, G, table, buckets, levels, batch)
        for b, elems, w in zip(buckets, got, want):
            assert len(elems) == (len(b) + (1 << levels) - 1) >> levels
            assert fold(G, elems) == w
        if levels == 6:  # 40 records -> one element per non-empty bucket: the bucket sum itself
            assert [e[0] if e else None for e in got] == [w if b else None for b, w in zip(buckets, want)]


def test_pair_tree_random_sizes(hc):
    """Random bucket sizes (many empty) and random records over a small table, so equal and opposite entries meet."""
    rng = random.Random(77)
    table = [None] + G1.points(rng, 9)
    buckets = [[(rng.randrange(10), rng.randrange(2)) for _ in range(rng.choice([0, 0, 1, 2, 3, 4, 6, 11, 23]))] for _ in range(60)]
    want = [bucket_sum(G1, table, b) for b in buckets]
    for batch, levels in ((4, 2), (16, 3), (32, 5)):
        got = run_tree(hc, G1, table, buckets, levels, batch)
        for b, elems, w in zip(buckets, got, want):
            assert len(elems) == (len(b) + (1 << levels) - 1) >> levels
            assert fold(G1, elems) == w


_TABLE = None


def _table():
    global _TABLE
    if _TABLE is None:
        _TABLE = [None] + G1.points(random.Random(5), 7)
    return _TABLE


@settings(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(sizes=st.lists(st.integers(min_value=0, max_value=37), min_size=1, max_size=12),
       levels=st.integers(min_value=1, max_value=7), batch=st.sampled_from([4, 16, 32]), seed=st.integers(min_value=0, max_value=2**31))
def test_pair_tree_property(hc, sizes, levels, batch, seed):
    """Any bucket shape, any number of levels, any batch size: every bucket keeps its sum and has ceil(k / 2^levels) elements
    (entries drawn from an 8-point table with the identity in it, so equal, opposite and identity operands are common)."""
    rng = random.Random(seed)
    table = _table()
    buckets = [[(rng.randrange(len(table)), rng.randrange(2)) for _ in range(k)] for k in sizes]
    got = run_tree(hc, G1, table, buckets, levels, batch)
    for b, elems in zip(buckets, got):
        assert len(elems) == (len(b) + (1 << levels) - 1) >> levels
        assert fold(G1, elems) == bucket_sum(G1, table, b)
