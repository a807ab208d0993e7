"""Committed golden fixtures (tests/golden/proofs.json, made by tests/gen_golden.py from Oracle A).

CPU (`-m "not gpu"`): Oracle A and Oracle B (C) re-derive every fixture -- the oracle cannot drift without this
failing -- and the pinned facts of the reference hold (config #1: h = [r - 64]; every honest proof verifies).
GPU (`-m gpu`): device setup, prove, h(x) and verify against the committed numbers, through the C ABI."""

import importlib
import json
import os

import pytest

from oracle import circuit, groth16 as og, oracle_b as ob, oracle_fast as of
from oracle.fields import FR

HERE = os.path.dirname(os.path.abspath(__file__))
P = FR.p
GOLD = json.load(open(os.path.join(HERE, "golden", "proofs.json")))
CASES = GOLD["cases"]


def _pt(v):  # JSON lists -> the tuples the oracle / host mirror use
    if v is None:
        return None
    return tuple(_pt(x) if isinstance(x, list) else x for x in v)


def _rep(c):
    r = c["rep"]
    rows = lambda m: [[(a, b) for a, b in row] for row in m]
    return circuit.DummyRep(u=rows(r["u"]), v=rows(r["v"]), w=rows(r["w"]), roots=list(r["roots"]), input=r["input"])


def test_fixture_file_is_sane():
    assert GOLD["field_modulus"] == P and len(CASES) == 5
    by = {c["name"]: c for c in CASES}
    assert by["simple_zk"]["h"] == [P - 64] and by["simple_zk"]["weights"] == [1, 2, 34, 6, 3, 4]
    assert by["single_mult_honest_bn"]["crs"]["xi_t"] == []  # mod.rs:168, :412
    assert all(c["verify"] == (c["name"] != "horner8_omega_invalid_witness") for c in CASES)


@pytest.mark.parametrize("c", CASES, ids=[c["name"] for c in CASES])
def test_oracles_reproduce_golden(c):
    B = og.BN254Backend()
    rep = _rep(c)
    dense = og.qap_from_root_rep(FR, rep)
    sig = og.setup(B, dense, tuple(c["toxic"]))
    assert list(sig[0].xi) == [_pt(p) for p in c["crs"]["xi1"]] and list(sig[1].xi) == [_pt(p) for p in c["crs"]["xi2"]]
    assert list(sig[0].sum_delta) == [_pt(p) for p in c["crs"]["sum_delta"]]
    pr = og.prove(B, dense, sig, c["weights"], c["r"], c["s"])
    want = (_pt(c["proof"]["a"]), _pt(c["proof"]["b"]), _pt(c["proof"]["c"]))
    assert (pr.a, pr.b, pr.c) == want
    a, b, cc, h = ob.prove(dense, sig, c["weights"], c["r"], c["s"])  # the C restatement
    assert (a, b, cc) == want and h == c["h"]
    if c["name"].startswith("horner8_omega"):  # Oracle F (NTT + Pippenger) only knows the roots-of-unity domain
        n, wts = len(rep.roots), c["weights"]
        idx = {root: k for k, root in enumerate(rep.roots)}
        ev = lambda rows: [sum(wts[i] * val for i, row in enumerate(rows) for (root, val) in row if idx[root] == k) % P
                           for k in range(n)]
        fa, fb, fc, fh = of.prove(n, rep.input, ev(rep.u), ev(rep.v), sig, wts, c["r"], c["s"], threads=2)
        assert (fa, fb, fc) == want and fh[:len(c["h"])] == c["h"] and not any(fh[len(c["h"]):])


@pytest.mark.gpu
@pytest.mark.parametrize("c", CASES, ids=[c["name"] for c in CASES])
def test_device_matches_golden(c):
    zk = importlib.import_module("zksnark-rs_b200")
    ctx = zk.Context(0)
    try:
        q = zk.QAP.from_root_representation(ctx, _rep(c))
        crs = zk.setup(ctx, q, tuple(c["toxic"]))
        d = crs.download()
        for k in ("alpha1", "beta1", "delta1", "beta2", "gamma2", "delta2"):
            assert d[k] == _pt(c["crs"][k]), k
        for k in ("xi1", "xi_t", "sum_gamma", "sum_delta", "xi2"):
            assert d[k] == [_pt(p) for p in c["crs"][k]], k
        u, v, h = zk.qap_h(ctx, q, c["weights"])
        n = len(c["rep"]["roots"])
        pad = lambda xs, k: list(xs) + [0] * (k - len(xs))
        assert u == pad(c["u_sum"], n) and v == pad(c["v_sum"], n)
        k = max(len(c["h"]), len(h))  # n = 1: the reference's quotient is [0], the device returns no coefficients
        assert pad(h, k) == pad(c["h"], k)
        got = zk.prove(ctx, q, crs, c["weights"], c["r"], c["s"])
        assert (got.a, got.b, got.c) == (_pt(c["proof"]["a"]), _pt(c["proof"]["b"]), _pt(c["proof"]["c"]))
        assert zk.verify(ctx, crs, c["inputs"], got) == c["verify"]
    finally:
        ctx.close()
