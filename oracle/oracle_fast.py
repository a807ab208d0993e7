"""ctypes front-end of Oracle F (oracle/oracle_fast.c): the prove() path with fast algorithms (NTT, Pippenger) on all
host threads -- test infrastructure and bench.py's `cpu_best_effort` context number only."""

from __future__ import annotations

import ctypes as C
import hashlib
import os
import subprocess

from . import oracle_b as ob

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_build", "liboracle_fast.so")
_lib = None


def build() -> str:
    """Compile oracle_fast.c (which #includes oracle_b.c); content-hash stamp over both sources."""
    digest = hashlib.sha256(b"".join(open(os.path.join(HERE, f), "rb").read() for f in ("oracle_fast.c", "oracle_b.c"))).hexdigest()
    stamp = SO + ".sha256"
    if not (os.path.exists(SO) and os.path.exists(stamp) and open(stamp).read().strip() == digest):
        os.makedirs(os.path.dirname(SO), exist_ok=True)
        subprocess.check_call(["gcc", "-O3", "-march=native", "-fPIC", "-Wall", "-Wno-unused-function", "-pthread", "-shared",
                               "-o", SO, os.path.join(HERE, "oracle_fast.c")])
        with open(stamp, "w") as f:
            f.write(digest)
    return SO


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.of_time_prove.restype = C.c_double
        _lib.of_time_prove.argtypes = [C.c_uint, C.c_int, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
    return _lib


def threads_default() -> int:
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def ntt(values, inverse=False, threads=1):
    """dft / idft of field/mod.rs:508-537 at the 2^k-th root of unity 5^((r-1)/2^k), natural order."""
    n = len(values)
    assert n and n & (n - 1) == 0
    buf = ob._pack(values)
    lib().of_ntt(buf, C.c_uint(n.bit_length() - 1), C.c_int(1 if inverse else 0), C.c_int(threads))
    return ob._unpack(buf, n)


def msm_g1(scalars, pts, threads=1):
    out = (C.c_uint64 * 8)()
    lib().of_msm_g1(ob._pack(scalars), ob.g1_pack(pts), C.c_size_t(len(scalars)), C.c_int(threads), out)
    return ob.g1_unpack(out)


def msm_g2(scalars, pts, threads=1):
    out = (C.c_uint64 * 16)()
    lib().of_msm_g2(ob._pack(scalars), ob.g2_pack(pts), C.c_size_t(len(scalars)), C.c_int(threads), out)
    return ob.g2_unpack(out)


class _ProveIn(C.Structure):
    _fields_ = [("n", C.c_size_t), ("n_input", C.c_size_t), ("n_sd", C.c_size_t),
                ("alpha1", C.c_void_p), ("beta1", C.c_void_p), ("delta1", C.c_void_p), ("xi1", C.c_void_p),
                ("xi_t", C.c_void_p), ("sum_delta", C.c_void_p), ("beta2", C.c_void_p), ("delta2", C.c_void_p),
                ("xi2", C.c_void_p)]


def prove(n, n_input, a_evals, b_evals, sigma, weights, r, s, threads=1):
    """groth16::prove on the roots-of-unity domain from the evaluation vectors A_k = <weights, u(w^k)>, B_k likewise.
    Returns (a, b, c, h) with h the n-1 quotient coefficients (trailing zeros kept)."""
    s1, s2 = sigma
    assert len(s1.xi) == n and len(s1.xi_t) == n - 1 and len(a_evals) == n and len(b_evals) == n
    pin = _ProveIn()
    pin.n, pin.n_input, pin.n_sd = n, n_input, len(s1.sum_delta)
    bufs = {"alpha1": ob.g1_pack([s1.alpha]), "beta1": ob.g1_pack([s1.beta]), "delta1": ob.g1_pack([s1.delta]),
            "xi1": ob.g1_pack(s1.xi), "xi_t": ob.g1_pack(s1.xi_t), "sum_delta": ob.g1_pack(s1.sum_delta),
            "beta2": ob.g2_pack([s2.beta]), "delta2": ob.g2_pack([s2.delta]), "xi2": ob.g2_pack(s2.xi)}
    for k, b in bufs.items():
        setattr(pin, k, C.addressof(b))
    proof = (C.c_uint64 * 32)()
    hbuf = (C.c_uint64 * (4 * n))()
    rc = lib().of_prove(C.byref(pin), ob._pack(a_evals), ob._pack(b_evals), ob._pack(weights), C.c_size_t(len(weights)),
                        ob._pack([r]), ob._pack([s]), proof, hbuf, C.c_int(threads))
    if rc != 0:
        raise ValueError(f"of_prove rc={rc}")
    return ob.g1_unpack(proof, 0), ob.g2_unpack(proof, 8), ob.g1_unpack(proof, 24), ob._unpack(hbuf, n - 1)


# ---- array-level (numpy uint64 limbs) entry points for full-size parity tests --------------------------------
def _np():
    import numpy as np
    return np


def _arr(a, cols=4):
    np = _np()
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, cols)
    return a


def ntt_np(data, inverse=False, threads=None, coset_shift=None):
    """dft / idft of an (n, 4) uint64 limb array (canonical residues), natural order; returns a new array.
    coset_shift g: forward evaluates on g*w^i (input scaled by g^i); inverse undoes it (output scaled by g^-i)."""
    np = _np()
    a = _arr(data).copy()
    n = a.shape[0]
    assert n and n & (n - 1) == 0
    threads = threads or threads_default()
    L = lib()
    if coset_shift is not None and not inverse:
        L.of_scale_powers(a.ctypes.data_as(C.c_void_p), C.c_size_t(n), ob._pack([coset_shift]))
    L.of_ntt(a.ctypes.data_as(C.c_void_p), C.c_uint(n.bit_length() - 1), C.c_int(1 if inverse else 0), C.c_int(threads))
    if coset_shift is not None and inverse:
        from .fields import FR
        L.of_scale_powers(a.ctypes.data_as(C.c_void_p), C.c_size_t(n), ob._pack([pow(coset_shift, FR.p - 2, FR.p)]))
    return a


def qap_evals_np(m, n, rows, weights):
    """A_k = sum_i a_i u_i(w^k), B_k likewise (mod.rs:233-246 on the root domain) from the by-wire CSR rows
    [(row_ptr u64[m+1], gate u32[nnz], coeff u64[nnz, 4])] for u, v(, w) and an (nw, 4) limb array of weights."""
    np = _np()
    w = _arr(weights)
    A = np.zeros((n, 4), dtype=np.uint64)
    B = np.zeros((n, 4), dtype=np.uint64)
    keep = []
    args = [C.c_size_t(m), C.c_size_t(n)]
    for ptr, gate, coef in rows[:2]:
        ptr = np.ascontiguousarray(ptr, dtype=np.uint64)
        gate = np.ascontiguousarray(gate, dtype=np.uint32)
        coef = _arr(coef) if len(gate) else np.zeros((1, 4), dtype=np.uint64)
        assert len(ptr) == m + 1
        keep += [ptr, gate, coef]
        args += [ptr.ctypes.data_as(C.c_void_p), gate.ctypes.data_as(C.c_void_p), coef.ctypes.data_as(C.c_void_p)]
    args += [w.ctypes.data_as(C.c_void_p), C.c_size_t(w.shape[0]), A.ctypes.data_as(C.c_void_p), B.ctypes.data_as(C.c_void_p)]
    lib().of_qap_evals(*args)
    return A, B


def qap_h_np(A, B, threads=None):
    """(u_sum, v_sum, h) as (n, 4) limb arrays from the evaluation vectors; h[n-1] = 0 (n - 1 quotient coefficients)."""
    np = _np()
    A, B = _arr(A), _arr(B)
    n = A.shape[0]
    assert n >= 2 and n & (n - 1) == 0 and B.shape[0] == n
    u, v, h = (np.zeros((n, 4), dtype=np.uint64) for _ in range(3))
    rc = lib().of_qap_h(A.ctypes.data_as(C.c_void_p), B.ctypes.data_as(C.c_void_p), C.c_uint(n.bit_length() - 1),
                        u.ctypes.data_as(C.c_void_p), v.ctypes.data_as(C.c_void_p), h.ctypes.data_as(C.c_void_p),
                        C.c_int(threads or threads_default()))
    assert rc == 0
    return u, v, h


def prove_np(n, n_input, A, B, crs, weights, r, s, threads=None):
    """prove() from limb arrays: crs = dict of uint64 arrays in the zkb_crs_host layout (alpha1, beta1, delta1 (1, 8); xi1 (n, 8);
    xi_t (n-1, 8); sum_delta (k, 8); beta2, delta2 (1, 16); xi2 (n, 16)).  Returns the proof as a 32-limb array (a | b | c)."""
    np = _np()
    A, B, w = _arr(A), _arr(B), _arr(weights)
    pin = _ProveIn()
    pin.n, pin.n_input, pin.n_sd = n, n_input, crs["sum_delta"].shape[0]
    keep = {}
    for k in ("alpha1", "beta1", "delta1", "xi1", "xi_t", "sum_delta", "beta2", "delta2", "xi2"):
        keep[k] = np.ascontiguousarray(crs[k], dtype=np.uint64)
        setattr(pin, k, keep[k].ctypes.data if keep[k].size else None)
    proof = np.zeros(32, dtype=np.uint64)
    rc = lib().of_prove(C.byref(pin), A.ctypes.data_as(C.c_void_p), B.ctypes.data_as(C.c_void_p), w.ctypes.data_as(C.c_void_p),
                        C.c_size_t(w.shape[0]), ob._pack([r]), ob._pack([s]), proof.ctypes.data_as(C.c_void_p), None,
                        C.c_int(threads or threads_default()))
    if rc != 0:
        raise ValueError(f"of_prove rc={rc}")
    return proof


def time_prove(log_n, threads=None, seed=1):
    """Seconds for one full-size proof (synthetic shape m = 2n+2, input = 2) with fast algorithms on `threads` host threads.
    Returns (seconds, {"poly": s, "g1": s, "g2": s}, check) -- check = proof.a.x, equal across thread counts."""
    from .bn254 import G2_GEN
    threads = threads or threads_default()
    g2 = ob.g2_pack([G2_GEN])
    tm = (C.c_double * 3)()
    chk = (C.c_uint64 * 4)()
    t = lib().of_time_prove(C.c_uint(log_n), C.c_int(threads), C.c_uint64(seed), C.addressof(g2), tm, chk)
    return t, {"poly": tm[0], "g1": tm[1], "g2": tm[2]}, ob._unpack(chk, 1)[0]
