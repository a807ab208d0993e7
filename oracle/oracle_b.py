"""ctypes front-end of Oracle B (oracle/oracle_b.c) -- test infrastructure only."""

from __future__ import annotations

import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_build", "liboracle_b.so")
_lib = None


def build() -> str:
    """Compile oracle_b.c (content-hash stamp, not mtimes: snapshots do not preserve mtime order)."""
    import hashlib
    src = os.path.join(HERE, "oracle_b.c")
    digest = hashlib.sha256(open(src, "rb").read()).hexdigest()
    stamp = SO + ".sha256"
    if not (os.path.exists(SO) and os.path.exists(stamp) and open(stamp).read().strip() == digest):
        os.makedirs(os.path.dirname(SO), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-march=native", "-fPIC", "-Wall", "-Wno-unused-function", "-shared",
                               "-o", SO, src])
        with open(stamp, "w") as f:
            f.write(digest)
    return SO


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        for name in ("ob_time_g1_terms", "ob_time_mul_rows", "ob_time_div_steps", "ob_time_wsum_rows"):
            getattr(_lib, name).restype = C.c_double
        _lib.ob_time_g1_terms.argtypes = [C.c_size_t, C.c_uint64]
        _lib.ob_time_g2_terms.restype = C.c_double
        _lib.ob_time_g2_terms.argtypes = [C.c_size_t, C.c_uint64, C.c_void_p]
        for name in ("ob_time_mul_rows", "ob_time_div_steps", "ob_time_wsum_rows"):
            getattr(_lib, name).argtypes = [C.c_size_t, C.c_size_t, C.c_uint64]
    return _lib


def _pack(vals):
    buf = (C.c_uint64 * (4 * len(vals)))()
    for i, v in enumerate(vals):
        for j in range(4):
            buf[4 * i + j] = (v >> (64 * j)) & 0xFFFFFFFFFFFFFFFF
    return buf


def _unpack(buf, n):
    return [sum(buf[4 * i + j] << (64 * j) for j in range(4)) for i in range(n)]


def g1_pack(pts):
    flat = []
    for P in pts:
        flat += [0, 0] if P is None else [P[0], P[1]]
    return _pack(flat)


def g2_pack(pts):
    flat = []
    for P in pts:
        flat += [0, 0, 0, 0] if P is None else [P[0][0], P[0][1], P[1][0], P[1][1]]
    return _pack(flat)


def g1_unpack(buf, off=0):
    x, y = _unpack(buf, off // 4 + 2)[off // 4:]
    return None if x == 0 and y == 0 else (x, y)


def g2_unpack(buf, off=0):
    a, b, c, d = _unpack(buf, off // 4 + 4)[off // 4:]
    return None if not (a or b or c or d) else ((a, b), (c, d))


def fr_mul(a, b):
    out = (C.c_uint64 * 4)()
    lib().ob_fr_mul(_pack([a]), _pack([b]), out)
    return _unpack(out, 1)[0]


def fr_inv(a):
    out = (C.c_uint64 * 4)()
    lib().ob_fr_inv(_pack([a]), out)
    return _unpack(out, 1)[0]


def fq_mul(a, b):
    out = (C.c_uint64 * 4)()
    lib().ob_fq_mul(_pack([a]), _pack([b]), out)
    return _unpack(out, 1)[0]


def g1_add(P, Q):
    out = (C.c_uint64 * 8)()
    lib().ob_g1_add(g1_pack([P]), g1_pack([Q]), out)
    return g1_unpack(out)


def g1_mul(P, k):
    out = (C.c_uint64 * 8)()
    lib().ob_g1_mul(g1_pack([P]), _pack([k]), out)
    return g1_unpack(out)


def g2_add(P, Q):
    out = (C.c_uint64 * 16)()
    lib().ob_g2_add(g2_pack([P]), g2_pack([Q]), out)
    return g2_unpack(out)


def g2_mul(P, k):
    out = (C.c_uint64 * 16)()
    lib().ob_g2_mul(g2_pack([P]), _pack([k]), out)
    return g2_unpack(out)


def msm_g1(scalars, pts):
    out = (C.c_uint64 * 8)()
    lib().ob_msm_g1(_pack(scalars), g1_pack(pts), C.c_size_t(len(scalars)), out)
    return g1_unpack(out)


def msm_g2(scalars, pts):
    out = (C.c_uint64 * 16)()
    lib().ob_msm_g2(_pack(scalars), g2_pack(pts), C.c_size_t(len(scalars)), out)
    return g2_unpack(out)


def poly_mul(a, b):
    out = (C.c_uint64 * (4 * (len(a) + len(b) + 1)))()
    n = C.c_size_t()
    lib().ob_poly_mul(_pack(a), C.c_size_t(len(a)), _pack(b), C.c_size_t(len(b)), out, C.byref(n))
    return _unpack(out, n.value)


def poly_div(a, b):
    out = (C.c_uint64 * (4 * (len(a) + 2)))()
    n = C.c_size_t()
    rc = lib().ob_poly_div(_pack(a), C.c_size_t(len(a)), _pack(b), C.c_size_t(len(b)), out, C.byref(n))
    if rc != 0:
        raise ZeroDivisionError("Dividend must be non-zero")
    return _unpack(out, n.value)


class _ProveIn(C.Structure):
    _fields_ = [("m", C.c_size_t), ("stride", C.c_size_t), ("nt", C.c_size_t), ("n_input", C.c_size_t),
                ("u", C.c_void_p), ("v", C.c_void_p), ("w", C.c_void_p), ("t", C.c_void_p),
                ("ulen", C.c_void_p), ("vlen", C.c_void_p), ("wlen", C.c_void_p),
                ("n_xi", C.c_size_t), ("n_xit", C.c_size_t), ("n_sd", C.c_size_t),
                ("alpha1", C.c_void_p), ("beta1", C.c_void_p), ("delta1", C.c_void_p), ("xi1", C.c_void_p),
                ("xi_t", C.c_void_p), ("sum_delta", C.c_void_p), ("beta2", C.c_void_p), ("delta2", C.c_void_p),
                ("xi2", C.c_void_p)]


def prove(qap, sigma, weights, r, s):
    """groth16::prove on a dense oracle.groth16.QAP / (SigmaG1, SigmaG2); returns (a, b, c, h)."""
    s1, s2 = sigma
    m = min(len(qap.u), len(qap.v), len(qap.w))
    stride = max([1] + [len(p) for mat in (qap.u, qap.v, qap.w) for p in mat])
    keep = []

    def dense(mat):
        flat, lens = [], (C.c_size_t * m)()
        for i in range(m):
            lens[i] = len(mat[i])
            flat += list(mat[i]) + [0] * (stride - len(mat[i]))
        buf = _pack(flat)
        keep.extend([buf, lens])
        return C.addressof(buf), C.addressof(lens)

    pin = _ProveIn()
    pin.m, pin.stride, pin.nt, pin.n_input = m, stride, len(qap.t), qap.input
    pin.u, pin.ulen = dense(qap.u)
    pin.v, pin.vlen = dense(qap.v)
    pin.w, pin.wlen = dense(qap.w)
    bufs = {"t": _pack(qap.t), "alpha1": g1_pack([s1.alpha]), "beta1": g1_pack([s1.beta]), "delta1": g1_pack([s1.delta]),
            "xi1": g1_pack(s1.xi), "xi_t": g1_pack(s1.xi_t), "sum_delta": g1_pack(s1.sum_delta),
            "beta2": g2_pack([s2.beta]), "delta2": g2_pack([s2.delta]), "xi2": g2_pack(s2.xi)}
    for k, b in bufs.items():
        setattr(pin, k, C.addressof(b))
    pin.n_xi, pin.n_xit, pin.n_sd = len(s1.xi), len(s1.xi_t), len(s1.sum_delta)
    proof = (C.c_uint64 * 32)()
    hbuf = (C.c_uint64 * (4 * (2 * stride + 2)))()
    hlen = C.c_size_t()
    rc = lib().ob_prove(C.byref(pin), _pack(weights), C.c_size_t(len(weights)), _pack([r]), _pack([s]), proof, hbuf,
                        C.byref(hlen))
    if rc != 0:
        raise ZeroDivisionError(f"ob_prove rc={rc}")
    return g1_unpack(proof, 0), g2_unpack(proof, 8), g1_unpack(proof, 24), _unpack(hbuf, hlen.value)


class _FullIn(C.Structure):
    _fields_ = [("log_n", C.c_size_t), ("m", C.c_size_t), ("n_input", C.c_size_t),
                ("row_ptr", C.c_void_p * 3), ("gate", C.c_void_p * 3), ("coeff", C.c_void_p * 3), ("omega_inv", C.c_void_p),
                ("n_sd", C.c_size_t),
                ("alpha1", C.c_void_p), ("beta1", C.c_void_p), ("delta1", C.c_void_p), ("xi1", C.c_void_p),
                ("xi_t", C.c_void_p), ("sum_delta", C.c_void_p), ("beta2", C.c_void_p), ("delta2", C.c_void_p),
                ("xi2", C.c_void_p)]


def prove_full_omega(log_n, m, n_input, rows, crs_raw, weights_limbs, r, s):
    """The reference's prove() (mod.rs:213-296, its own O(m n + n^2) algorithms, one thread) run IN FULL and timed, on a
    roots-of-unity QAP given as sparse rows -- numpy (row_ptr u64, gate u32, coeff (nnz, 4) u64) for u, v, w -- and a
    CRS as uint64 limb arrays in the zkb_crs_host layout.  Returns (seconds of prove, seconds of building the dense QAP
    (not part of prove: QAP::from, fr.rs:140-173), proof limbs (32,) a | b | c)."""
    import numpy as np
    L = lib()
    L.ob_prove_full_omega.restype = C.c_double
    fin = _FullIn()
    fin.log_n, fin.m, fin.n_input = log_n, m, n_input
    keep = []
    for t, (ptr, gate, coeff) in enumerate(rows):
        a = [np.ascontiguousarray(ptr, dtype=np.uint64), np.ascontiguousarray(gate, dtype=np.uint32),
             np.ascontiguousarray(coeff, dtype=np.uint64)]
        keep += a
        fin.row_ptr[t], fin.gate[t], fin.coeff[t] = (x.ctypes.data for x in a)
    p = 21888242871839275222246405745257275088548364400416034343698204186575808495617
    winv = _pack([pow(pow(5, (p - 1) >> log_n, p), -1, p)])
    fin.omega_inv = C.addressof(winv)
    arrs = {k: np.ascontiguousarray(v, dtype=np.uint64) for k, v in crs_raw.items()}
    for k in ("alpha1", "beta1", "delta1", "xi1", "xi_t", "sum_delta", "beta2", "delta2", "xi2"):
        setattr(fin, k, arrs[k].ctypes.data)
    fin.n_sd = arrs["sum_delta"].shape[0]
    w = np.ascontiguousarray(weights_limbs, dtype=np.uint64)
    proof = np.zeros(32, dtype=np.uint64)
    build = C.c_double()
    sec = L.ob_prove_full_omega(C.byref(fin), C.c_void_p(w.ctypes.data), C.c_size_t(w.shape[0]), _pack([r]), _pack([s]),
                                C.c_void_p(proof.ctypes.data), C.byref(build))
    if sec < 0:
        raise RuntimeError(f"ob_prove_full_omega failed ({sec})")
    del keep
    return sec, build.value, proof


G2_GEN = [10857046999023057135944570762232829481370756359578518086990519993285655852781,
          11559732032986387107991004021392285783925812861821192530917403151452391805634,
          8495653923123431417604973247489272438418190587263600148770280649306958101930,
          4082367875863433681332203403145435568316851327593401208105741076214120093531]


def horner_rows_np(n):
    """Sparse rows (by wire) of the n-gate Horner circuit of SURVEY.md 8d (row order of ASTParser, circuit/mod.rs:230-526:
    0 unity, 1 x, 2 y, t_k -> 2k+1, c_k -> 2k+2, c_n -> 2n+1), as (row_ptr, gate, coeff) numpy triples for u, v, w."""
    import numpy as np
    m = 2 * n + 2
    rows_u = {0: [n - 1], 1: list(range(n - 1))}
    rows_v, rows_w = {}, {2: [n - 1]}
    for k in range(1, n + 1):
        c_row = 2 * k + 2 if k < n else 2 * n + 1
        rows_v.setdefault(c_row, []).append(k - 1)
        if k > 1:
            rows_v.setdefault(2 * (k - 1) + 1, []).append(k - 1)
        if k < n:
            rows_w.setdefault(2 * k + 1, []).append(k - 1)
    out = []
    for d in (rows_u, rows_v, rows_w):
        ptr = np.zeros(m + 1, dtype=np.uint64)
        gates = []
        for i in range(m):
            gates += d.get(i, [])
            ptr[i + 1] = len(gates)
        coeff = np.zeros((len(gates), 4), dtype=np.uint64)
        coeff[:, 0] = 1
        out.append((ptr, np.asarray(gates, dtype=np.uint32), coeff))
    return m, 2, out


def time_full_prove(log_n, seed=1):
    """One complete run of the reference's prove() at n = 2^log_n on the Horner QAP with stand-in CRS points (valid curve
    points; the running time does not depend on which) and a random witness.  Returns (seconds, seconds to build the dense QAP)."""
    import random

    import numpy as np
    n = 1 << log_n
    m, n_input, rows = horner_rows_np(n)
    L = lib()
    n1 = 3 + n + (n - 1) + (m - n_input - 1)
    g1 = np.zeros((n1, 8), dtype=np.uint64)
    g2 = np.zeros((2 + n, 16), dtype=np.uint64)
    L.ob_fill_points(C.c_void_p(g1.ctypes.data), C.c_size_t(n1), C.c_void_p(g2.ctypes.data), C.c_size_t(2 + n), _pack(G2_GEN))
    crs = {"alpha1": g1[0:1], "beta1": g1[1:2], "delta1": g1[2:3], "xi1": g1[3:3 + n], "xi_t": g1[3 + n:2 + 2 * n],
           "sum_delta": g1[2 + 2 * n:], "beta2": g2[0:1], "delta2": g2[1:2], "xi2": g2[2:]}
    rng = random.Random(seed)
    p = 21888242871839275222246405745257275088548364400416034343698204186575808495617
    w = np.frombuffer(b"".join(rng.randrange(p).to_bytes(32, "little") for _ in range(m)), dtype="<u8").reshape(m, 4).copy()
    sec, build, _ = prove_full_omega(log_n, m, n_input, rows, crs, w, rng.randrange(1, p), rng.randrange(1, p))
    return sec, build
