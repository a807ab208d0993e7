"""GPU parity at the sizes the bench reports: the CUDA path against Oracle F, VECTOR BY VECTOR.

tests/test_gpu_parity.py compares the device with the literal restatement of the reference up to 16 gates (the
literal schoolbook Mul / long division are O(n^2)) and checks larger sizes through one scalar identity.  This file
closes that gap with Oracle F (oracle/oracle_fast.c: inverse NTTs + one size-2n product + Pippenger on the host
threads), which tests/test_oracle_fast.py pins bit for bit against the literal restatement (Oracles A and B):

  * u_sum, v_sum (mod.rs:233-246) and h = (u_sum * v_sum - w_sum) / t (mod.rs:277; coefficient_poly.rs:93-157;
    field/mod.rs:428-469) as FULL VECTORS at 2^12 (two NTT passes), 2^16 and 2^20 (three passes), for a satisfying
    and for a NON-satisfying witness (the reference discards the remainder of the division: coefficient_poly.rs:155);
  * the proof {a, b, c} at the same sizes against Oracle F's proof over the same CRS (downloaded from the device),
    valid and invalid witnesses -- the invalid case has no closed form, so this is its only full-size check;
  * zkb_ntt_fr as full vectors at 2^20 and 2^22 (three-pass plans), forward, inverse and on a coset.

Oracle F's quotient is the high half of one size-2n product; the device computes h = (c - d g^-k) / 2 from two
size-n transforms on a coset (prove.cu) -- different algorithms, same vectors.
"""

import ctypes as C
import importlib
import random

import numpy as np
import pytest

from oracle import oracle_fast as of
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


def _case(log_n, valid, seed):
    n = 1 << log_n
    rng = random.Random(seed)
    m, n_input, rows = zg.horner_qap_rows(n)
    x, cs = rng.randrange(1, P), [rng.randrange(P) for _ in range(n)]
    wit = zg.horner_witness(n, x, cs)
    if not valid:  # break one inner gate, the output gate and a wire no gate multiplies
        wit[3] = (wit[3] + 1) % P
        wit[2 * (n // 2) + 1] = rng.randrange(P)
        wit[-1] = rng.randrange(P)
    toxic = tuple(rng.randrange(1, P) for _ in range(5))
    r, s = rng.randrange(1, P), rng.randrange(1, P)
    return n, m, n_input, rows, zg.fr_limbs(wit), toxic, r, s


@pytest.mark.parametrize("log_n", [12, 16, 20])
@pytest.mark.parametrize("valid", [True, False])
def test_h_vectors_and_proof_match_oracle_f(ctx, log_n, valid):
    n, m, n_input, rows, w, toxic, r, s = _case(log_n, valid, 1000 + 2 * log_n + valid)
    A, B = of.qap_evals_np(m, n, rows, w)
    u, v, h = of.qap_h_np(A, B)
    q = zk.QAP(ctx, n, m, n_input, rows)
    gu, gv, gh = zg.qap_h_raw(ctx, q, w)
    assert np.array_equal(gu, u), "u_sum differs from Oracle F"
    assert np.array_equal(gv, v), "v_sum differs from Oracle F"
    assert np.array_equal(gh, h), "h differs from Oracle F"
    assert not gh[n - 1].any()
    if not valid:  # the witness really is non-satisfying: u*v - w is not a multiple of t, i.e. A_k B_k != C_k somewhere
        wi = zg.limbs_to_ints(w[:8])
        assert (wi[1] * wi[4]) % P != wi[3]
    crs = zk.setup(ctx, q, toxic)
    got = zk.prove(ctx, q, crs, w, r, s)
    raw = crs.download_raw()
    want = of.prove_np(n, n_input, A, B, raw, w, r, s)
    got_limbs = np.concatenate([zg.g1_pack([got.a]).reshape(-1), zg.g2_pack([got.b]).reshape(-1), zg.g1_pack([got.c]).reshape(-1)])
    assert np.array_equal(got_limbs, want), "proof differs from Oracle F over the same CRS"
    pub = zg.limbs_to_ints(w[1:1 + n_input])
    assert zk.verify(ctx, crs, pub, got) == valid
    # batch (several proofs in flight) and device-resident witness give the same proof
    batch = zk.prove_batch(ctx, q, crs, [w, w, w], [r] * 3, [s] * 3)
    assert all((p.a, p.b, p.c) == (got.a, got.b, got.c) for p in batch)
    crs.free()
    q.free()


@pytest.mark.parametrize("log_n", [20, 22])
def test_ntt_full_vector_matches_oracle_f(ctx, log_n):
    """Every output of the three-pass transforms, not samples: forward, inverse, coset forward, coset inverse."""
    n = 1 << log_n
    rng = np.random.default_rng(4000 + log_n)
    a = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=(n, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64((1 << 60) - 1)  # < 2^252 < r: canonical residues
    a[0] = zg.fr_limbs([P - 1])[0]
    a[1] = 0
    g = 7
    d = ctx.dev_alloc(a.nbytes)
    y = np.empty_like(a)
    try:
        for inverse, shift in ((0, None), (1, None), (0, g), (1, g)):
            ctx.h2d(d, a)
            sh = zg.fr_limbs([shift]) if shift is not None else None
            ctx.check(ctx.lib.zkb_ntt_fr(ctx.h, C.c_void_p(d), log_n, inverse, sh.ctypes.data_as(C.c_void_p) if sh is not None else None),
                      "zkb_ntt_fr")
            ctx.d2h(y, d)
            want = of.ntt_np(a, inverse=bool(inverse), coset_shift=shift)
            assert np.array_equal(y, want), f"NTT 2^{log_n} inverse={inverse} coset={shift}: device != Oracle F"
    finally:
        ctx.dev_free(d)
