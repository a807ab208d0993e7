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
