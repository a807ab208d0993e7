"""Wire format of QAP / CRS / Proof (include/zkb200.h zkb_wire_*, csrc/wire.cu): the reference has no serialisation
(private fields, no accessors: groth16/mod.rs:60-128), so this surface is new.  CPU: round trips and
rejection of damaged records (host-only entry points: they work without a GPU).  GPU: a QAP and a CRS that went through
bytes prove the same proof."""

import importlib
import os
import random

import ctypes as C
import struct

import numpy as np
import pytest
from hypothesis import given, settings
from hypothesis import strategies as st

from oracle.fields import FR

zk = importlib.import_module("zksnark-rs_b200")
zg = importlib.import_module("zksnark-rs_b200.groth16")
P = FR.p


@pytest.fixture(scope="module", autouse=True)
def _lib():
    if not os.path.exists(zk.lib_path()):
        importlib.import_module("zksnark-rs_b200.build").build()
    zk.load_library()


def _rand_rows(rng, n, m):
    rows = []
    for _t in range(3):
        per = [sorted(rng.sample(range(n), rng.randrange(0, min(n, 4) + 1))) for _ in range(m)]
        ptr = np.zeros(m + 1, dtype=np.uint64)
        ptr[1:] = np.cumsum([len(r) for r in per])
        gate = np.asarray([g for r in per for g in r], dtype=np.uint32)
        coeff = zg.fr_limbs([rng.randrange(P) for _ in range(len(gate))]).reshape(-1, 4)
        rows.append((ptr, gate, coeff))
    return rows


@pytest.mark.parametrize("n,m,with_roots", [(4, 7, False), (5, 9, True), (1, 2, True), (64, 130, False)])
def test_qap_round_trip(n, m, with_roots):
    rng = random.Random(n * 100 + m)
    rows = _rand_rows(rng, n, m)
    roots = [rng.randrange(P) for _ in range(n)] if with_roots else None
    data = zg.qap_to_bytes(n, m, 2 if m > 2 else 1, rows, roots)
    assert zg.wire_kind(data) == 1 and len(data) % 8 == 0
    n2, m2, ni2, rows2, roots2 = zg.qap_from_bytes(data)
    assert (n2, m2, ni2) == (n, m, 2 if m > 2 else 1)
    for (a, b, c), (a2, b2, c2) in zip(rows, rows2):
        assert np.array_equal(a, a2) and np.array_equal(b, b2) and np.array_equal(c, c2)
    assert (roots2 is None) == (roots is None)
    if roots is not None:
        assert zg.limbs_to_ints(roots2) == roots
    assert zg.qap_to_bytes(n2, m2, ni2, rows2, roots2) == data  # canonical: one byte string per object


def test_crs_and_proof_round_trip():
    rng = np.random.default_rng(7)
    n, ng, nd = 6, 3, 11
    shapes = {"alpha1": (1, 8), "beta1": (1, 8), "delta1": (1, 8), "xi1": (n, 8), "xi_t": (n - 1, 8), "sum_gamma": (ng, 8),
              "sum_delta": (nd, 8), "beta2": (1, 16), "gamma2": (1, 16), "delta2": (1, 16), "xi2": (n, 16)}
    raw = {k: rng.integers(0, 1 << 63, size=s, dtype=np.uint64) for k, s in shapes.items()}
    data = zg.crs_raw_to_bytes(raw)
    assert zg.wire_kind(data) == 2 and len(data) == 64 + 8 * sum(a.size for a in raw.values())
    back = zg.crs_raw_from_bytes(data)
    assert all(np.array_equal(raw[k], back[k]) for k in shapes)
    one = {k: (v[:0] if k == "xi_t" else v[:1]) if k in ("xi1", "xi_t", "xi2") else v for k, v in raw.items()}  # n = 1: xi_t is empty
    assert all(np.array_equal(one[k], zg.crs_raw_from_bytes(zg.crs_raw_to_bytes(one))[k]) for k in shapes)
    pr = zg.Proof(a=(3, 4), b=((5, 6), (7, 8)), c=None)  # identity is a legal point (all-zero)
    pb = zg.proof_to_bytes(pr)
    assert len(pb) == zg.WIRE_PROOF_BYTES == 320 and zg.wire_kind(pb) == 3
    assert zg.proof_from_bytes(pb) == pr


def test_damaged_records_are_rejected():
    rng = random.Random(3)
    data = bytearray(zg.qap_to_bytes(4, 7, 2, _rand_rows(rng, 4, 7)))
    for pos, what in ((0, "magic"), (8, "version"), (len(data) - 1, "checksum"), (70, "checksum")):
        bad = bytearray(data)
        bad[pos] ^= 0x40
        with pytest.raises(zk.ZkbError, match=what):
            zg.qap_from_bytes(bytes(bad))
    with pytest.raises(zk.ZkbError, match="truncated"):
        zg.qap_from_bytes(bytes(data[:-8]))
    with pytest.raises(zk.ZkbError, match="kind"):
        zg.crs_raw_from_bytes(bytes(data))
    with pytest.raises(zk.ZkbError, match="shorter"):
        zg.proof_from_bytes(b"\0" * 16)


def _fnv1a(b):
    h = 0xcbf29ce484222325
    for x in b:
        h = ((h ^ x) * 0x100000001b3) & 0xFFFFFFFFFFFFFFFF
    return h


def _reseal(buf):
    """Recompute the checksum of a (tampered) record so that the reader gets past it and has to judge the structure."""
    total = struct.unpack_from("<Q", buf, 16)[0]
    if 64 <= total <= len(buf):
        struct.pack_into("<Q", buf, 56, _fnv1a(bytes(buf[64:total])))
    return buf


_BASE_QAP = None


def _base_qap():
    global _BASE_QAP
    if _BASE_QAP is None:
        rng = random.Random(11)
        _BASE_QAP = zg.qap_to_bytes(5, 9, 2, _rand_rows(rng, 5, 9), [rng.randrange(P) for _ in range(5)])
    return _BASE_QAP


@settings(max_examples=300, deadline=None)
@given(edits=st.lists(st.tuples(st.integers(min_value=16, max_value=400), st.integers(min_value=0, max_value=2**64 - 1)), min_size=1, max_size=4),
       small=st.booleans(), cut=st.integers(min_value=0, max_value=64))
def test_tampered_qap_records_with_valid_checksums_stay_in_bounds(edits, small, cut):
    """A record whose sizes, offsets or row counts were rewritten and whose checksum was recomputed (an adversary, not line
    noise) is either refused or yields views that lie inside the caller's buffer: the reader trusts no length field."""
    lib = zk.load_library()
    data = bytearray(_base_qap())
    for off, val in edits:
        off = (off // 8) * 8
        if off + 8 <= len(data) and not 56 <= off < 64:
            struct.pack_into("<Q", data, off, val % 64 if small else val)
    data = _reseal(data)
    if cut:
        data = _reseal(data[: max(0, len(data) - 8 * cut)])
    buf = zg._aligned_copy(bytes(data))
    host = zg._QapHost()
    rc = lib.zkb_wire_read_qap(zg._ptr(buf), len(data), C.byref(host))
    if rc != 0:
        return
    base = zg._ptr(buf).value
    end = base + len(data)
    m = int(host.m)
    for t in range(3):
        rp = C.cast(host.row_ptr[t], C.c_void_p).value
        assert base <= rp and rp + (m + 1) * 8 <= end
        nnz = int(np.ctypeslib.as_array(C.cast(host.row_ptr[t], C.POINTER(C.c_uint64)), shape=(m + 1,))[m])
        g, c = C.cast(host.gate[t], C.c_void_p).value, C.cast(host.coeff[t], C.c_void_p).value
        assert base <= g and g + nnz * 4 <= end and base <= c and c + nnz * 32 <= end
    if host.roots:
        r = C.cast(host.roots, C.c_void_p).value
        assert base <= r and r + int(host.n) * 32 <= end


@pytest.mark.gpu
def test_objects_that_went_through_bytes_prove_the_same_proof():
    ctx = zk.Context(0)
    try:
        n = 256
        rng = random.Random(9)
        m, n_input, rows = zg.horner_qap_rows(n)
        wit = zg.horner_witness(n, rng.randrange(1, P), [rng.randrange(P) for _ in range(n)])
        toxic = tuple(rng.randrange(1, P) for _ in range(5))
        r, s = rng.randrange(1, P), rng.randrange(1, P)
        q = zk.QAP(ctx, n, m, n_input, rows)
        crs = zk.setup(ctx, q, toxic)
        want = zk.prove(ctx, q, crs, wit, r, s)
        q2 = zg.qap_upload_bytes(ctx, zg.qap_to_bytes(n, m, n_input, rows))
        crs_bytes = zg.crs_raw_to_bytes(crs.download_raw())
        crs2 = zg.crs_upload_bytes(ctx, crs_bytes)
        got = zk.prove(ctx, q2, crs2, wit, r, s)
        assert got == want and zg.proof_from_bytes(zg.proof_to_bytes(got)) == want
        assert zk.verify(ctx, crs2, wit[1:3], got)
        # explicit roots survive the trip too (parser numbering, dense path)
        n3 = 12
        m3, ni3, rows3 = zg.horner_qap_rows(n3)
        q3 = zg.qap_upload_bytes(ctx, zg.qap_to_bytes(n3, m3, ni3, rows3, roots=list(range(1, n3 + 1))))
        q3b = zk.QAP(ctx, n3, m3, ni3, rows3, roots=list(range(1, n3 + 1)))
        w3 = zg.horner_witness(n3, 5, list(range(2, 2 + n3)))
        assert zk.prove(ctx, q3, zk.setup(ctx, q3, toxic), w3, r, s) == zk.prove(ctx, q3b, zk.setup(ctx, q3b, toxic), w3, r, s)
    finally:
        ctx.close()
