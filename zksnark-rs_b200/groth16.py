"""Host-side mirror of the reference's Groth16 interface for the prove() path, over libzkb200.so.

Names follow /root/reference/src/groth16/mod.rs: ``QAP`` (:60-67), ``SigmaG1``/``SigmaG2`` (:105-121,
held together here as one device-resident ``CRS``), ``Proof`` (:124-128), ``setup`` (:134-197),
``prove`` (:213-296).  Field elements are Python ints (canonical residues); G1 points are ``None``
(identity) or ``(x, y)``; G2 points are ``None`` or ``((x0, x1), (y0, y1))``.

Everything that computes goes through the CUDA library; nothing here does field or curve
arithmetic on the CPU (packing ints into limbs and building index arrays is all the host does).
"""

from __future__ import annotations

import ctypes as C
import os
import secrets
from dataclasses import dataclass

import numpy as np

FR_MODULUS = 21888242871839275222246405745257275088548364400416034343698204186575808495617
TWO_ADICITY = 28
_HERE = os.path.dirname(os.path.abspath(__file__))


class ZkbError(RuntimeError):
    """Raised for every non-zero status of the C ABI (the reference panics instead)."""


def lib_path() -> str:
    """The in-tree library; ZKB200_LIB overrides it (developer A/B builds of one kernel)."""
    return os.environ.get("ZKB200_LIB") or os.path.join(_HERE, "libzkb200.so")


class _QapHost(C.Structure):
    _fields_ = [("n", C.c_uint64), ("m", C.c_uint64), ("n_input", C.c_uint64),
                ("row_ptr", C.c_void_p * 3), ("gate", C.c_void_p * 3), ("coeff", C.c_void_p * 3), ("roots", C.c_void_p)]


class _CrsHost(C.Structure):
    _fields_ = [("n", C.c_uint64), ("n_sum_gamma", C.c_uint64), ("n_sum_delta", C.c_uint64),
                ("alpha1", C.c_void_p), ("beta1", C.c_void_p), ("delta1", C.c_void_p),
                ("xi1", C.c_void_p), ("xi_t", C.c_void_p), ("sum_gamma", C.c_void_p), ("sum_delta", C.c_void_p),
                ("beta2", C.c_void_p), ("gamma2", C.c_void_p), ("delta2", C.c_void_p), ("xi2", C.c_void_p)]


class _ProofC(C.Structure):
    _fields_ = [("a", C.c_uint64 * 8), ("b", C.c_uint64 * 16), ("c", C.c_uint64 * 8)]


# every symbol include/zkb200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
ABI = {
    "zkb_ctx_create": (C.c_int, [C.POINTER(_P), C.c_int]),
    "zkb_ctx_destroy": (None, [_P]),
    "zkb_last_error": (C.c_char_p, [_P]),
    "zkb_launch_count": (C.c_uint64, [_P]),
    "zkb_profile": (C.c_int, [_P, C.c_int]),
    "zkb_profile_read": (C.c_int, [_P, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "zkb_host_alloc": (C.c_int, [C.POINTER(_P), C.c_size_t]),
    "zkb_host_free": (None, [_P]),
    "zkb_dev_alloc": (C.c_int, [_P, C.POINTER(_P), C.c_size_t]),
    "zkb_dev_free": (None, [_P, _P]),
    "zkb_memcpy_h2d": (C.c_int, [_P, _P, _P, C.c_size_t]),
    "zkb_memcpy_d2h": (C.c_int, [_P, _P, _P, C.c_size_t]),
    "zkb_sync": (C.c_int, [_P]),
    "zkb_stream": (_P, [_P]),
    "zkb_qap_upload": (C.c_int, [_P, C.POINTER(_QapHost), C.POINTER(_P)]),
    "zkb_qap_free": (None, [_P, _P]),
    "zkb_crs_upload": (C.c_int, [_P, C.POINTER(_CrsHost), C.c_int, C.c_int, C.POINTER(_P)]),
    "zkb_setup": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.POINTER(_P)]),
    "zkb_crs_dims": (C.c_int, [_P, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "zkb_crs_download": (C.c_int, [_P, _P, C.POINTER(_CrsHost)]),
    "zkb_crs_free": (None, [_P, _P]),
    "zkb_prove": (C.c_int, [_P, _P, _P, _P, _P, _P, C.POINTER(_ProofC)]),
    "zkb_prove_dev": (C.c_int, [_P, _P, _P, _P, _P, _P, C.POINTER(_ProofC)]),
    "zkb_field_op": (C.c_int, [_P, C.c_int, C.c_int, _P, _P, _P, _P, _P, C.c_size_t]),
    "zkb_trace_dump": (C.c_int, [_P, C.c_char_p]),
    "zkb_prove_batch": (C.c_int, [_P, _P, _P, _P, C.c_int, _P, _P, C.c_size_t, _P]),
    "zkb_prove_partial": (C.c_int, [_P, _P, _P, _P, C.c_int, _P, _P, _P]),
    "zkb_prove_combine": (C.c_int, [_P, _P, C.c_int, C.POINTER(_ProofC)]),
    "zkb_prove_combine_batch": (C.c_int, [_P, _P, C.c_int, C.c_size_t, _P]),
    "zkb_qap_h": (C.c_int, [_P, _P, _P, _P, _P, _P]),
    "zkb_ntt_fr": (C.c_int, [_P, _P, C.c_uint32, C.c_int, _P]),
    "zkb_ntt_combine": (C.c_int, [_P, _P, C.c_uint32, C.c_uint32, C.c_int, C.c_uint64, C.c_uint64, _P]),
    "zkb_ntt_fr_raw": (C.c_int, [_P, _P, C.c_uint32, C.c_int]),
    "zkb_fr_to_mont": (C.c_int, [_P, _P, C.c_size_t, C.c_int]),
    "zkb_bases_upload": (C.c_int, [_P, C.c_int, _P, C.c_size_t, C.POINTER(_P)]),
    "zkb_bases_generate": (C.c_int, [_P, C.c_int, _P, C.c_size_t, C.POINTER(_P)]),
    "zkb_bases_download": (C.c_int, [_P, _P, _P]),
    "zkb_bases_free": (None, [_P, _P]),
    "zkb_msm": (C.c_int, [_P, _P, _P, C.c_int, C.c_size_t, C.c_int, _P]),
    "zkb_msm_windows": (C.c_int, [_P, _P, _P, C.c_int, C.c_size_t, C.c_int, C.c_int, _P]),
    "zkb_points_sum": (C.c_int, [_P, C.c_int, _P, C.c_size_t, _P]),
    "zkb_verify": (C.c_int, [_P, _P, _P, C.c_size_t, C.POINTER(_ProofC), C.POINTER(C.c_int)]),
    "zkb_verify_batch": (C.c_int, [_P, _P, _P, C.c_size_t, _P, C.c_size_t, _P]),
    "zkb_pairing": (C.c_int, [_P, _P, _P, C.c_size_t, _P]),
    "zkb_bench_modmul": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "zkb_comm_create": (C.c_int, [_P, C.c_int, C.c_int, C.c_uint32, C.POINTER(_P), _P]),
    "zkb_comm_connect": (C.c_int, [_P, _P]),
    "zkb_comm_destroy": (None, [_P]),
    "zkb_comm_info": (C.c_int, [_P, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "zkb_setup_shard": (C.c_int, [_P, _P, _P, _P, C.POINTER(_P)]),
    "zkb_crs_upload_shard": (C.c_int, [_P, _P, C.POINTER(_CrsHost), C.POINTER(_P)]),
    "zkb_prove_shard": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, _P, _P, C.POINTER(_ProofC)]),
    "zkb_prove_shard_enqueue": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, _P, _P, C.c_int]),
    "zkb_prove_shard_collect": (C.c_int, [_P, _P, C.c_int, C.POINTER(_ProofC)]),
    "zkb_prove_shard_batch": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, _P, _P, C.c_size_t, _P]),
    "zkb_ntt_shard": (C.c_int, [_P, _P, _P, C.c_uint32, C.c_int, C.c_int]),
    "zkb_wire_kind": (C.c_int, [_P, C.c_uint64, C.POINTER(C.c_int), C.POINTER(C.c_uint64)]),
    "zkb_wire_size_qap": (C.c_int, [C.POINTER(_QapHost), C.POINTER(C.c_uint64)]),
    "zkb_wire_write_qap": (C.c_int, [C.POINTER(_QapHost), _P, C.c_uint64]),
    "zkb_wire_read_qap": (C.c_int, [_P, C.c_uint64, C.POINTER(_QapHost)]),
    "zkb_wire_size_crs": (C.c_int, [C.POINTER(_CrsHost), C.POINTER(C.c_uint64)]),
    "zkb_wire_write_crs": (C.c_int, [C.POINTER(_CrsHost), _P, C.c_uint64]),
    "zkb_wire_read_crs": (C.c_int, [_P, C.c_uint64, C.POINTER(_CrsHost)]),
    "zkb_wire_write_proof": (C.c_int, [C.POINTER(_ProofC), _P, C.c_uint64]),
    "zkb_wire_read_proof": (C.c_int, [_P, C.c_uint64, C.POINTER(_ProofC)]),
    "zkb_witness_plan_create": (C.c_int, [_P, _P, _P, C.c_size_t, C.c_int, C.POINTER(_P)]),
    "zkb_witness_plan_info": (C.c_int, [_P, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "zkb_witness_generate": (C.c_int, [_P, _P, _P, C.c_size_t, C.c_int, _P, C.c_int]),
    "zkb_witness_plan_free": (None, [_P, _P]),
    "zkb_witness_levels": (C.c_int, [C.POINTER(_QapHost), _P, C.c_size_t, C.c_int, _P, C.POINTER(C.c_uint64)]),
}
WITNESS_PROGRAM_ORDER = 1
WIRE_PROOF_BYTES = 320
COMM_HANDLE_BYTES = 128

_lib = None


def load_library():
    """dlopen libzkb200.so and bind every symbol of include/zkb200.h.  Fails loudly when the CUDA
    library has not been built -- there is no fallback implementation."""
    global _lib
    if _lib is None:
        path = lib_path()
        if not os.path.exists(path):
            raise ZkbError(f"{path} not found: build it with `python zksnark-rs_b200/build.py` "
                           "(nvcc, sm_100a); this package has no CPU fallback")
        lib = C.CDLL(path)
        for name, (res, args) in ABI.items():
            fn = getattr(lib, name)  # AttributeError if the library lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


# ------------------------------------------------------------------------------------------------
# limb packing
def fr_limbs(values) -> np.ndarray:
    """ints -> (len, 4) uint64 little-endian limbs."""
    vals = list(values)
    buf = b"".join(int(v).to_bytes(32, "little") for v in vals)
    return np.frombuffer(buf, dtype="<u8").reshape(len(vals), 4).copy()


def limbs_to_ints(arr: np.ndarray) -> list:
    a = np.ascontiguousarray(arr, dtype="<u8").reshape(-1, 4)
    raw = a.tobytes()
    return [int.from_bytes(raw[32 * i:32 * i + 32], "little") for i in range(a.shape[0])]


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def g1_pack(points) -> np.ndarray:
    flat = []
    for P in points:
        flat += [0, 0] if P is None else [P[0], P[1]]
    return fr_limbs(flat).reshape(-1, 8)


def g2_pack(points) -> np.ndarray:
    flat = []
    for P in points:
        flat += [0, 0, 0, 0] if P is None else [P[0][0], P[0][1], P[1][0], P[1][1]]
    return fr_limbs(flat).reshape(-1, 16)


def g1_unpack(arr) -> list:
    v = limbs_to_ints(np.asarray(arr).reshape(-1, 4))
    out = []
    for i in range(0, len(v), 2):
        out.append(None if v[i] == 0 and v[i + 1] == 0 else (v[i], v[i + 1]))
    return out


def g2_unpack(arr) -> list:
    v = limbs_to_ints(np.asarray(arr).reshape(-1, 4))
    out = []
    for i in range(0, len(v), 4):
        q = v[i:i + 4]
        out.append(None if not any(q) else ((q[0], q[1]), (q[2], q[3])))
    return out


def omega(log_n: int) -> int:
    """Primitive 2^log_n-th root of unity of Fr used for the gate domain: 5^((r-1)/2^log_n)."""
    return pow(5, (FR_MODULUS - 1) >> log_n, FR_MODULUS)


def random_elem() -> int:
    """`Random::random_elem` for FrLocal (fr.rs:90-99): uniform and never zero."""
    return secrets.randbelow(FR_MODULUS - 1) + 1


# ------------------------------------------------------------------------------------------------
class Context:
    """One CUDA device (zkb_ctx)."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.zkb_ctx_create(C.byref(h), device)
        if rc != 0:
            raise ZkbError(f"zkb_ctx_create({device}) = {rc}: {self.lib.zkb_last_error(None).decode()}")
        self.h = h
        self.device = device

    def check(self, rc: int, what: str):
        if rc != 0:
            raise ZkbError(f"{what} = {rc}: {self.lib.zkb_last_error(self.h).decode()}")

    @property
    def launches(self) -> int:
        return int(self.lib.zkb_launch_count(self.h))

    def sync(self):
        self.check(self.lib.zkb_sync(self.h), "zkb_sync")

    def stream(self) -> int:
        return int(self.lib.zkb_stream(self.h) or 0)

    # raw device memory for resident inputs
    def dev_alloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        self.check(self.lib.zkb_dev_alloc(self.h, C.byref(p), nbytes), "zkb_dev_alloc")
        return p.value

    def dev_free(self, p: int):
        self.lib.zkb_dev_free(self.h, C.c_void_p(p))

    def h2d(self, dptr: int, arr: np.ndarray):
        arr = np.ascontiguousarray(arr)
        self.check(self.lib.zkb_memcpy_h2d(self.h, C.c_void_p(dptr), _ptr(arr), arr.nbytes), "zkb_memcpy_h2d")

    def d2h(self, arr: np.ndarray, dptr: int):
        self.check(self.lib.zkb_memcpy_d2h(self.h, _ptr(arr), C.c_void_p(dptr), arr.nbytes), "zkb_memcpy_d2h")

    def pinned(self, shape, dtype="<u8") -> np.ndarray:
        """numpy array backed by pinned host memory (for per-proof witness uploads)."""
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        self.check(self.lib.zkb_host_alloc(C.byref(p), nbytes), "zkb_host_alloc")
        buf = (C.c_uint8 * nbytes).from_address(p.value)
        arr = np.frombuffer(buf, dtype=dtype).reshape(shape)
        self._pinned = getattr(self, "_pinned", []) + [p]
        return arr

    def profile(self, enable):
        """False/0 off, True/1 tracked kernel classes, 2 = trace every launch (see trace_dump)."""
        self.check(self.lib.zkb_profile(self.h, int(enable)), "zkb_profile")

    def trace_dump(self, path: str):
        self.check(self.lib.zkb_trace_dump(self.h, path.encode()), "zkb_trace_dump")

    def profile_read(self, kind: int):
        """(total ms, launches, work units) of one tracked kernel class: 1 NTT passes, 2 G1 / 3 G2 bucket accumulation."""
        ms, cnt, units = C.c_double(), C.c_uint64(), C.c_uint64()
        self.check(self.lib.zkb_profile_read(self.h, kind, C.byref(ms), C.byref(cnt), C.byref(units)), "zkb_profile_read")
        return ms.value, cnt.value, units.value

    def bench_modmul(self, field: int = 1, iters: int = 2000):
        rate, ms = C.c_double(), C.c_double()
        self.check(self.lib.zkb_bench_modmul(self.h, field, iters, C.byref(rate), C.byref(ms)), "zkb_bench_modmul")
        return rate.value, ms.value

    def close(self):
        if getattr(self, "h", None):
            for p in getattr(self, "_pinned", []):
                self.lib.zkb_host_free(p)
            self._pinned = []
            self.lib.zkb_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ------------------------------------------------------------------------------------------------
def _csr(rows_of_pairs, n_rows):
    """list (per wire) of [(gate, coeff int)] -> row_ptr u64, gate u32, coeff limbs."""
    ptr = np.zeros(n_rows + 1, dtype=np.uint64)
    gates, coeffs = [], []
    for i, row in enumerate(rows_of_pairs):
        for g, c in row:
            gates.append(g)
            coeffs.append(c)
        ptr[i + 1] = len(gates)
    return ptr, np.asarray(gates, dtype=np.uint32), fr_limbs(coeffs).reshape(-1, 4)


def horner_qap_rows(n: int):
    """CSR rows (by wire) of the synthetic n-gate Horner circuit deg_{n-1} (generalises
    test_programs/deg_15.zk).  Row order as ASTParser produces it (circuit/mod.rs:230-526): 0 unity,
    1 x, 2 y, t_k -> 2k+1, c_k -> 2k+2 (k<n), c_n -> 2n+1.  Gate k (1-based) is gate index k-1.
    Returns (m, n_input, [(row_ptr, gate, coeff)] for u, v, w)."""
    assert n >= 2
    m = 2 * n + 2
    one = np.zeros((1, 4), dtype=np.uint64)
    one[0, 0] = 1
    k = np.arange(1, n, dtype=np.uint64)  # 1..n-1
    # u: row 0 -> gate n-1 ; row 1 -> gates 0..n-2
    pu = np.zeros(m + 1, dtype=np.uint64)
    pu[1] = 1
    pu[2:] = n
    gu = np.concatenate([[n - 1], np.arange(0, n - 1)]).astype(np.uint32)
    # v: rows 3..2n one entry each: row 2k+1 (t_k) -> gate k ; row 2k+2 (c_k) -> gate k-1 ; row 2n+1 (c_n) -> gate n-1
    pv = np.zeros(m + 1, dtype=np.uint64)
    pv[4:] = np.arange(1, m - 2, dtype=np.uint64)  # rows 3.. have one entry each
    gv = np.empty(m - 3, dtype=np.uint32)
    gv[0:2 * (n - 1):2] = k            # rows 3,5,.. = t_k
    gv[1:2 * (n - 1):2] = k - 1        # rows 4,6,.. = c_k
    gv[2 * (n - 1)] = n - 1            # row 2n+1 = c_n
    # w: row 2 (y) -> gate n-1 ; row 2k+1 (t_k) -> gate k-1
    cnt = np.zeros(m, dtype=np.uint64)
    cnt[2] = 1
    cnt[3:2 * n:2] = 1
    pw = np.zeros(m + 1, dtype=np.uint64)
    pw[1:] = np.cumsum(cnt)
    gw = np.concatenate([[n - 1], np.arange(0, n - 1)]).astype(np.uint32)
    rows = []
    for p, g in ((pu, gu), (pv, gv), (pw, gw)):
        rows.append((p, g, np.repeat(one, len(g), axis=0)))
    return m, 2, rows


def horner_witness(n: int, x: int, cs) -> list:
    """Wire assignment (row order of horner_qap_rows) for inputs x, c_1..c_n."""
    p = FR_MODULUS
    a = [0] * (2 * n + 2)
    a[0], a[1] = 1, x % p
    acc = x * cs[0] % p
    a[3], a[4] = acc, cs[0] % p
    for k in range(2, n):
        acc = x * (acc + cs[k - 1]) % p
        a[2 * k + 1], a[2 * k + 2] = acc, cs[k - 1] % p
    a[2 * n + 1] = cs[n - 1] % p
    a[2] = (acc + cs[n - 1]) % p
    return a


class QAP:
    """Device-resident `QAP<CoefficientPoly<FrLocal>>` (groth16/mod.rs:60-67), stored as sparse
    evaluation rows on the roots-of-unity domain instead of 3*m dense coefficient vectors."""

    def __init__(self, ctx: Context, n: int, m: int, n_input: int, rows, roots=None):
        """rows: (row_ptr, gate, coeff) CSR triples for u, v, w.  roots: None = the n-th roots of unity
        (fast NTT path); else the n explicit roots (ints), e.g. ASTParser's 1..=n (dense O(n^2) path)."""
        self.ctx, self.n, self.m, self.input = ctx, n, m, n_input
        self.degree = n
        host = _QapHost()
        host.n, host.m, host.n_input = n, m, n_input
        keep = []
        if roots is not None:
            ra = fr_limbs([r % FR_MODULUS for r in roots])
            assert ra.shape == (n, 4)
            keep.append(ra)
            host.roots = ra.ctypes.data
        else:
            host.roots = None
        for t, (ptr, gate, coeff) in enumerate(rows):
            ptr = np.ascontiguousarray(ptr, dtype=np.uint64)
            gate = np.ascontiguousarray(gate, dtype=np.uint32)
            coeff = np.ascontiguousarray(coeff, dtype=np.uint64)
            assert ptr.shape == (m + 1,) and coeff.size == gate.size * 4
            keep += [ptr, gate, coeff]
            host.row_ptr[t], host.gate[t], host.coeff[t] = ptr.ctypes.data, gate.ctypes.data, coeff.ctypes.data
        h = C.c_void_p()
        ctx.check(ctx.lib.zkb_qap_upload(ctx.h, C.byref(host), C.byref(h)), "zkb_qap_upload")
        self.h = h

    @classmethod
    def from_root_representation(cls, ctx: Context, rep, reindex: bool = False) -> "QAP":
        """`From<RootRepresentation> for QAP` (fr.rs:140-173).  ``rep`` has u, v, w (per wire: list of
        (root, value)), roots, input -- the DummyRep data model (circuit/dummy_rep.rs:7-13).  Roots that
        are omega_n^0 .. omega_n^(n-1) in order take the NTT path; any other pairwise-distinct roots (the
        parser's 1..=n, circuit/mod.rs:517) take the dense O(n^2) device path (n <= 32768), bit-identical to the
        reference on that QAP.  ``reindex=True`` instead moves gate k (whatever its root) to omega^k on the next power
        of two (extra gates are empty: 0 * 0 = 0), i.e. builds the QAP the reference would build for the SAME circuit
        had its gates been numbered with roots of unity (the reference lets the caller choose the roots:
        dummy_rep.rs:11, circuit/mod.rs:99-104): O(n log n) proofs at any size, verifying under the CRS of THAT
        QAP -- not bit-identical to proofs over the 1..=n numbering."""
        n = len(rep.roots)
        if not (len(rep.u) == len(rep.v) == len(rep.w)):
            raise ZkbError("QAP: u, v, w must have the same number of rows")  # assert at fr.rs:157-158
        m = len(rep.u)
        roots = [r % FR_MODULUS for r in rep.roots]
        index = {r: k for k, r in enumerate(roots)}
        fast = n >= 2 and n & (n - 1) == 0
        if fast:  # omega^0 .. omega^(n-1) in order -> NTT path
            w, acc = omega(n.bit_length() - 1), 1
            for k in range(n):
                if roots[k] != acc:
                    fast = False
                    break
                acc = acc * w % FR_MODULUS
        rows = []
        for mat in (rep.u, rep.v, rep.w):
            try:
                rows.append(_csr([[(index[r % FR_MODULUS], c % FR_MODULUS) for r, c in row] for row in mat], m))
            except KeyError as e:  # the reference's Lagrange interpolation would silently ignore nothing: a point off the domain is a bug
                raise ZkbError(f"QAP: row entry at {e} is not one of the roots") from None
        if reindex and not fast:
            n2 = max(2, 1 << (n - 1).bit_length())
            return cls(ctx, n2, m, rep.input, rows, roots=None)
        return cls(ctx, n, m, rep.input, rows, roots=None if fast else roots)

    @classmethod
    def horner(cls, ctx: Context, n: int) -> "QAP":
        m, n_input, rows = horner_qap_rows(n)
        return cls(ctx, n, m, n_input, rows)

    def free(self):
        if getattr(self, "h", None) and self.ctx.h:
            self.ctx.lib.zkb_qap_free(self.ctx.h, self.h)
        self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def _crs_host(sigma_g1, sigma_g2):
    """zkb_crs_host over packed copies of the reference-shaped objects; returns (struct, arrays to keep alive)."""
    host = _CrsHost()
    host.n, host.n_sum_gamma, host.n_sum_delta = len(sigma_g1.xi), len(sigma_g1.sum_gamma), len(sigma_g1.sum_delta)
    arrs = {
        "alpha1": g1_pack([sigma_g1.alpha]), "beta1": g1_pack([sigma_g1.beta]), "delta1": g1_pack([sigma_g1.delta]),
        "xi1": g1_pack(sigma_g1.xi), "xi_t": g1_pack(sigma_g1.xi_t), "sum_gamma": g1_pack(sigma_g1.sum_gamma),
        "sum_delta": g1_pack(sigma_g1.sum_delta),
        "beta2": g2_pack([sigma_g2.beta]), "gamma2": g2_pack([sigma_g2.gamma]), "delta2": g2_pack([sigma_g2.delta]),
        "xi2": g2_pack(sigma_g2.xi),
    }
    for k, a in arrs.items():
        setattr(host, k, a.ctypes.data if a.size else None)
    return host, arrs


class CRS:
    """Device-resident (SigmaG1, SigmaG2) (groth16/mod.rs:105-121)."""

    def __init__(self, ctx: Context, h, rank=0, world=1):
        self.ctx, self.h, self.rank, self.world = ctx, h, rank, world

    @classmethod
    def upload(cls, ctx: Context, sigma_g1, sigma_g2, rank=0, world=1) -> "CRS":
        """From the reference-shaped objects (attributes alpha, beta, delta, xi, sum_gamma, sum_delta,
        xi_t / beta, gamma, delta, xi)."""
        host, keep = _crs_host(sigma_g1, sigma_g2)
        h = C.c_void_p()
        ctx.check(ctx.lib.zkb_crs_upload(ctx.h, C.byref(host), rank, world, C.byref(h)), "zkb_crs_upload")
        del keep
        return cls(ctx, h, rank, world)

    def dims(self):
        n, g, d = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self.ctx.check(self.ctx.lib.zkb_crs_dims(self.h, C.byref(n), C.byref(g), C.byref(d)), "zkb_crs_dims")
        return n.value, g.value, d.value

    def download_raw(self) -> dict:
        """The CRS as uint64 limb arrays in the zkb_crs_host layout (canonical affine coordinates)."""
        n, ng, nd = self.dims()
        host = _CrsHost()
        host.n, host.n_sum_gamma, host.n_sum_delta = n, ng, nd
        shapes = {"alpha1": (1, 8), "beta1": (1, 8), "delta1": (1, 8), "xi1": (n, 8), "xi_t": (n - 1, 8),
                  "sum_gamma": (ng, 8), "sum_delta": (nd, 8), "beta2": (1, 16), "gamma2": (1, 16), "delta2": (1, 16),
                  "xi2": (n, 16)}
        arrs = {k: np.zeros(s, dtype=np.uint64) for k, s in shapes.items()}
        for k, a in arrs.items():
            setattr(host, k, a.ctypes.data if a.size else None)
        self.ctx.check(self.ctx.lib.zkb_crs_download(self.ctx.h, self.h, C.byref(host)), "zkb_crs_download")
        return arrs

    def download(self) -> dict:
        arrs = self.download_raw()
        out = {}
        for k, a in arrs.items():
            pts = g2_unpack(a) if a.shape[1] == 16 else g1_unpack(a)
            out[k] = pts[0] if k in ("alpha1", "beta1", "delta1", "beta2", "gamma2", "delta2") else pts
        return out

    def free(self):
        if getattr(self, "h", None) and self.ctx.h:
            self.ctx.lib.zkb_crs_free(self.ctx.h, self.h)
        self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


@dataclass
class Proof:
    """groth16/mod.rs:124-128."""

    a: object
    b: object
    c: object


def _proof(pc: _ProofC) -> Proof:
    return Proof(a=g1_unpack(np.array(pc.a[:], dtype=np.uint64))[0],
                 b=g2_unpack(np.array(pc.b[:], dtype=np.uint64))[0],
                 c=g1_unpack(np.array(pc.c[:], dtype=np.uint64))[0])


def setup(ctx: Context, qap: QAP, toxic=None, rank=0, world=1) -> CRS:
    """groth16::setup (mod.rs:134-197).  ``toxic`` = (alpha, beta, gamma, delta, x); drawn with
    random_elem() when omitted, as the reference does (:139-145)."""
    if toxic is None:
        toxic = tuple(random_elem() for _ in range(5))
    t = fr_limbs(toxic)
    h = C.c_void_p()
    ctx.check(ctx.lib.zkb_setup(ctx.h, qap.h, _ptr(t), rank, world, C.byref(h)), "zkb_setup")
    return CRS(ctx, h, rank, world)


def _weights_array(qap: QAP, weights) -> np.ndarray:
    if isinstance(weights, np.ndarray):
        w = np.ascontiguousarray(weights, dtype=np.uint64).reshape(-1, 4)
    else:
        w = fr_limbs(weights)
    if w.shape[0] != qap.m:  # every zip in prove() truncates to the shorter side (mod.rs:237..288)
        full = np.zeros((qap.m, 4), dtype=np.uint64)
        k = min(qap.m, w.shape[0])
        full[:k] = w[:k]
        w = full
    return w


def prove(ctx: Context, qap: QAP, crs: CRS, weights, r=None, s=None) -> Proof:
    """groth16::prove (mod.rs:213-296).  r, s are drawn with random_elem() (mod.rs:231) unless given."""
    r = random_elem() if r is None else r
    s = random_elem() if s is None else s
    w = _weights_array(qap, weights)
    rl, sl = fr_limbs([r]), fr_limbs([s])
    out = _ProofC()
    ctx.check(ctx.lib.zkb_prove(ctx.h, qap.h, crs.h, _ptr(w), _ptr(rl), _ptr(sl), C.byref(out)), "zkb_prove")
    return _proof(out)


def prove_dev(ctx: Context, qap: QAP, crs: CRS, d_weights: int, r: int, s: int) -> Proof:
    rl, sl = fr_limbs([r]), fr_limbs([s])
    out = _ProofC()
    ctx.check(ctx.lib.zkb_prove_dev(ctx.h, qap.h, crs.h, C.c_void_p(d_weights), _ptr(rl), _ptr(sl), C.byref(out)),
              "zkb_prove_dev")
    return _proof(out)


def prove_batch(ctx: Context, qap: QAP, crs: CRS, weights, rs, ss, on_device=False) -> list:
    """`len(weights)` proofs with several in flight (two at 2^20, up to four at small sizes) (zkb_prove_batch).  weights: list of witness vectors
    (host: anything _weights_array accepts, ideally pinned (m, 4) uint64 arrays; device: pointers)."""
    k = len(weights)
    if on_device:
        keep = None
        ptrs = (C.c_void_p * k)(*[C.c_void_p(w) for w in weights])
    else:
        keep = [_weights_array(qap, w) for w in weights]
        ptrs = (C.c_void_p * k)(*[a.ctypes.data for a in keep])
    rl, sl = fr_limbs(list(rs)), fr_limbs(list(ss))
    out = (_ProofC * k)()
    ctx.check(ctx.lib.zkb_prove_batch(ctx.h, qap.h, crs.h, ptrs, 1 if on_device else 0, _ptr(rl), _ptr(sl), k, out),
              "zkb_prove_batch")
    del keep
    if crs.world != 1:  # sharded CRS: the records are this rank's partial sums, (k, 32) limbs
        return np.frombuffer(bytes(out), dtype=np.uint64).reshape(k, PARTIAL_LIMBS).copy()
    return [_proof(out[i]) for i in range(k)]


PARTIAL_LIMBS = 32


def prove_partial(ctx: Context, qap: QAP, crs: CRS, weights, r: int, s: int, on_device=False) -> np.ndarray:
    """This rank's partial sums of A, B, C over its CRS shard, in the Proof layout: 32 limbs
    (a 8 | b 16 | c 8).  Rank 0's shard also carries the fixed-point terms (alpha1 + r delta1 ...)."""
    out = np.zeros(PARTIAL_LIMBS, dtype=np.uint64)
    if on_device:
        wp = C.c_void_p(weights)
    else:
        w = _weights_array(qap, weights)
        wp = _ptr(w)
    rl, sl = fr_limbs([r]), fr_limbs([s])
    ctx.check(ctx.lib.zkb_prove_partial(ctx.h, qap.h, crs.h, wp, 1 if on_device else 0, _ptr(rl), _ptr(sl), _ptr(out)),
              "zkb_prove_partial")
    return out


def prove_combine(ctx: Context, partials: np.ndarray) -> Proof:
    """Fold the gathered per-rank partial sums (world x 32 limbs) into the proof."""
    p = np.ascontiguousarray(partials, dtype=np.uint64).reshape(-1, PARTIAL_LIMBS)
    out = _ProofC()
    ctx.check(ctx.lib.zkb_prove_combine(ctx.h, _ptr(p), p.shape[0], C.byref(out)), "zkb_prove_combine")
    return _proof(out)


def _proof_c(proof: Proof) -> _ProofC:
    pc = _ProofC()
    pc.a[:] = [int(v) for v in g1_pack([proof.a]).reshape(-1)]
    pc.b[:] = [int(v) for v in g2_pack([proof.b]).reshape(-1)]
    pc.c[:] = [int(v) for v in g1_pack([proof.c]).reshape(-1)]
    return pc


def verify(ctx: Context, crs: CRS, inputs, proof: Proof) -> bool:
    """groth16::verify (mod.rs:299-320): e(alpha1, beta2) e(sum_term, gamma2) e(C, delta2) == e(A, B), with
    sum_term over zip(sum_gamma, [1] ++ inputs).  The CRS is borrowed (the reference consumes it)."""
    inp = fr_limbs(list(inputs))
    pc = _proof_c(proof)
    ok = C.c_int(-1)
    ctx.check(ctx.lib.zkb_verify(ctx.h, crs.h, _ptr(inp) if len(inp) else None, len(inp), C.byref(pc), C.byref(ok)), "zkb_verify")
    return bool(ok.value)


def verify_batch(ctx: Context, crs: CRS, inputs, proofs) -> list:
    """One verdict per (inputs[i], proofs[i]) against the same CRS (zkb_verify_batch); every inputs[i] has the
    same length."""
    k = len(proofs)
    n_in = len(inputs[0]) if k else 0
    assert all(len(x) == n_in for x in inputs)
    inp = fr_limbs([v for row in inputs for v in row])
    arr = (_ProofC * k)(*[_proof_c(p) for p in proofs])
    ok = np.full(k, -1, dtype=np.int32)
    ctx.check(ctx.lib.zkb_verify_batch(ctx.h, crs.h, _ptr(inp) if inp.size else None, n_in, arr, k, _ptr(ok)), "zkb_verify_batch")
    return [bool(v) for v in ok]


def pairing(ctx: Context, pairs) -> list:
    """prod_i e(P_i, Q_i) for pairs (G1 point, G2 point) -- `EllipticEncryptable::pairing` (fr.rs:120-122) and the GT
    product (fr.rs:225-231).  Returns the 12 Fq residues of the GT element: (c0, c1) of the w^i coefficient, i = 0..5,
    in Fq12 = Fq2[w]/(w^6 - (9 + u))."""
    g1 = g1_pack([P for P, _ in pairs])
    g2 = g2_pack([Q for _, Q in pairs])
    out = np.zeros((12, 4), dtype=np.uint64)
    ctx.check(ctx.lib.zkb_pairing(ctx.h, _ptr(g1) if len(pairs) else None, _ptr(g2) if len(pairs) else None, len(pairs), _ptr(out)),
              "zkb_pairing")
    return limbs_to_ints(out)


def field_op(ctx: Context, field: int, op: int, a, b=None, c=None, d=None) -> list:
    """Element-wise device field arithmetic (zkb_field_op).  Fr/Fq: lists of ints; Fq2: lists of (c0, c1)."""
    n = len(a)
    b = a if b is None else b
    c = a if c is None else c
    d = a if d is None else d
    if field == 2:
        arrs = [fr_limbs([x for pair in v for x in pair]) for v in (a, b, c, d)]
    else:
        arrs = [fr_limbs(v) for v in (a, b, c, d)]
    out = np.zeros_like(arrs[0])
    ctx.check(ctx.lib.zkb_field_op(ctx.h, field, op, _ptr(arrs[0]), _ptr(arrs[1]), _ptr(arrs[2]), _ptr(arrs[3]), _ptr(out), n),
              "zkb_field_op")
    vals = limbs_to_ints(out)
    return [(vals[2 * i], vals[2 * i + 1]) for i in range(n)] if field == 2 else vals


def prove_combine_batch(ctx: Context, partials: np.ndarray) -> list:
    """partials: (world, count, 32) limbs (all-gathered prove_batch records of a sharded CRS) -> count proofs."""
    p = np.ascontiguousarray(partials, dtype=np.uint64)
    world, count = p.shape[0], p.shape[1]
    out = (_ProofC * count)()
    ctx.check(ctx.lib.zkb_prove_combine_batch(ctx.h, _ptr(p), world, count, out), "zkb_prove_combine_batch")
    return [_proof(out[i]) for i in range(count)]


def qap_h_raw(ctx: Context, qap: QAP, weights):
    """(u_sum, v_sum, h) as (n, 4) uint64 limb arrays (h[n-1] = 0): mod.rs:233-246 and :277."""
    w = _weights_array(qap, weights)
    outs = [np.zeros((qap.n, 4), dtype=np.uint64) for _ in range(3)]
    ctx.check(ctx.lib.zkb_qap_h(ctx.h, qap.h, _ptr(w), _ptr(outs[0]), _ptr(outs[1]), _ptr(outs[2])), "zkb_qap_h")
    return outs


def qap_h(ctx: Context, qap: QAP, weights):
    """(u_sum, v_sum, h) coefficient lists: mod.rs:233-246 and :277."""
    u, v, h = (limbs_to_ints(o) for o in qap_h_raw(ctx, qap, weights))
    return u, v, h[: qap.n - 1]


def ntt(ctx: Context, values, inverse=False, coset_shift=None) -> list:
    """dft / idft (field/mod.rs:508-537) of a power-of-two-length list at root omega_n."""
    n = len(values)
    if n == 0 or n & (n - 1):
        raise ZkbError("ntt: length must be a power of two")
    a = fr_limbs(values)
    d = ctx.dev_alloc(a.nbytes)
    try:
        ctx.h2d(d, a)
        sh = fr_limbs([coset_shift]) if coset_shift is not None else None
        ctx.check(ctx.lib.zkb_ntt_fr(ctx.h, C.c_void_p(d), n.bit_length() - 1, 1 if inverse else 0,
                                     _ptr(sh) if sh is not None else None), "zkb_ntt_fr")
        ctx.d2h(a, d)
    finally:
        ctx.dev_free(d)
    return limbs_to_ints(a)


def ntt_dev(ctx: Context, d_data: int, log_n: int, inverse=False):
    """zkb_ntt_fr in place on a device vector of canonical residues (natural order in and out)."""
    ctx.check(ctx.lib.zkb_ntt_fr(ctx.h, C.c_void_p(d_data), log_n, 1 if inverse else 0, None), "zkb_ntt_fr")


def ntt_combine(ctx: Context, d_parts: int, log_n: int, log_g: int, inverse: bool, k0: int, count: int, d_out: int):
    """This rank's slice of a transform whose outer dimension is sharded over 2^log_g ranks (zkb_ntt_combine)."""
    ctx.check(ctx.lib.zkb_ntt_combine(ctx.h, C.c_void_p(d_parts), log_n, log_g, 1 if inverse else 0, k0, count, C.c_void_p(d_out)),
              "zkb_ntt_combine")


class Bases:
    """Device-resident vector of G1 (group=1) or G2 (group=2) points."""

    def __init__(self, ctx: Context, h, group: int, n: int):
        self.ctx, self.h, self.group, self.n = ctx, h, group, n

    @classmethod
    def upload(cls, ctx: Context, group: int, points) -> "Bases":
        arr = g1_pack(points) if group == 1 else g2_pack(points)
        h = C.c_void_p()
        ctx.check(ctx.lib.zkb_bases_upload(ctx.h, group, _ptr(arr) if arr.size else None, len(points), C.byref(h)),
                  "zkb_bases_upload")
        return cls(ctx, h, group, len(points))

    @classmethod
    def generate(cls, ctx: Context, group: int, scalars) -> "Bases":
        """P_i = encrypt_g1(k_i) / encrypt_g2(k_i) (fr.rs:106-113) on the device."""
        arr = scalars if isinstance(scalars, np.ndarray) else fr_limbs(scalars)
        arr = np.ascontiguousarray(arr, dtype=np.uint64).reshape(-1, 4)
        h = C.c_void_p()
        ctx.check(ctx.lib.zkb_bases_generate(ctx.h, group, _ptr(arr), arr.shape[0], C.byref(h)), "zkb_bases_generate")
        return cls(ctx, h, group, arr.shape[0])

    def download(self) -> list:
        arr = np.zeros((self.n, 8 if self.group == 1 else 16), dtype=np.uint64)
        self.ctx.check(self.ctx.lib.zkb_bases_download(self.ctx.h, self.h, _ptr(arr)), "zkb_bases_download")
        return g1_unpack(arr) if self.group == 1 else g2_unpack(arr)

    def free(self):
        if getattr(self, "h", None) and self.ctx.h:
            self.ctx.lib.zkb_bases_free(self.ctx.h, self.h)
        self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def msm(ctx: Context, bases: Bases, scalars, window_bits: int = 0, on_device=False, n=None, windows=None):
    """sum_i scalars[i] * bases[i] -- the `.zip().map(exp_encrypted_g*).sum()` pattern (mod.rs:255-272).
    windows=(rank, world): only the table rows (windows) j = rank (mod world) -- the per-GPU partial sum of a
    window-sharded MSM (zkb_msm_windows); the `world` partial points sum to the full result."""
    out = np.zeros(8 if bases.group == 1 else 16, dtype=np.uint64)
    if on_device:
        sp, cnt = C.c_void_p(scalars), n
    else:
        arr = scalars if isinstance(scalars, np.ndarray) else fr_limbs(scalars)
        arr = np.ascontiguousarray(arr, dtype=np.uint64).reshape(-1, 4)
        sp, cnt = (_ptr(arr) if arr.size else None), arr.shape[0]
    if windows is not None:
        ctx.check(ctx.lib.zkb_msm_windows(ctx.h, bases.h, sp, 1 if on_device else 0, cnt, windows[0], windows[1], _ptr(out)),
                  "zkb_msm_windows")
    else:
        ctx.check(ctx.lib.zkb_msm(ctx.h, bases.h, sp, 1 if on_device else 0, cnt, window_bits, _ptr(out)), "zkb_msm")
    return (g1_unpack(out) if bases.group == 1 else g2_unpack(out))[0]


def points_sum(ctx: Context, group: int, points):
    arr = g1_pack(points) if group == 1 else g2_pack(points)
    out = np.zeros(8 if group == 1 else 16, dtype=np.uint64)
    ctx.check(ctx.lib.zkb_points_sum(ctx.h, group, _ptr(arr) if arr.size else None, len(points), _ptr(out)),
              "zkb_points_sum")
    return (g1_unpack(out) if group == 1 else g2_unpack(out))[0]


# ------------------------------------------------------------------------------------------------
# one proof over several GPUs (include/zkb200.h: zkb_comm_*, zkb_prove_shard*)
class Comm:
    """One rank's end of a multi-GPU communicator.  `Comm.create` allocates the exchange window and returns the
    128-byte handle; after the host side has gathered all handles in rank order (dist.connect does it over
    torch.distributed), `connect` maps the peers' windows."""

    def __init__(self, ctx: Context, h, rank: int, world: int, handle: bytes):
        self.ctx, self.h, self.rank, self.world, self.handle = ctx, h, rank, world, handle

    @classmethod
    def create(cls, ctx: Context, rank: int, world: int, max_log_n: int) -> "Comm":
        h = C.c_void_p()
        buf = (C.c_uint8 * COMM_HANDLE_BYTES)()
        ctx.check(ctx.lib.zkb_comm_create(ctx.h, rank, world, max_log_n, C.byref(h), buf), "zkb_comm_create")
        return cls(ctx, h, rank, world, bytes(buf))

    def connect(self, handles):
        """handles: the `world` 128-byte handles in rank order (list of bytes or one bytes object)."""
        blob = handles if isinstance(handles, (bytes, bytearray)) else b"".join(handles)
        assert len(blob) == self.world * COMM_HANDLE_BYTES
        buf = (C.c_uint8 * len(blob)).from_buffer_copy(blob)
        self.ctx.check(self.ctx.lib.zkb_comm_connect(self.h, buf), "zkb_comm_connect")
        return self

    def status(self) -> int:
        st = C.c_int(0)
        self.ctx.check(self.ctx.lib.zkb_comm_info(self.h, None, None, C.byref(st)), "zkb_comm_info")
        return st.value

    def free(self):
        if getattr(self, "h", None) and self.ctx.h:
            self.ctx.lib.zkb_comm_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def setup_shard(ctx: Context, comm: Comm, qap: QAP, toxic) -> CRS:
    """groth16::setup for this rank's shard of a proof that runs over all ranks of `comm` (zkb_setup_shard)."""
    t = fr_limbs(toxic)
    h = C.c_void_p()
    ctx.check(ctx.lib.zkb_setup_shard(ctx.h, comm.h, qap.h, _ptr(t), C.byref(h)), "zkb_setup_shard")
    return CRS(ctx, h, comm.rank, comm.world)


def crs_upload_shard(ctx: Context, comm: Comm, sigma_g1, sigma_g2) -> CRS:
    """This rank's shard of a reference-shaped CRS (zkb_crs_upload_shard)."""
    host, keep = _crs_host(sigma_g1, sigma_g2)
    h = C.c_void_p()
    ctx.check(ctx.lib.zkb_crs_upload_shard(ctx.h, comm.h, C.byref(host), C.byref(h)), "zkb_crs_upload_shard")
    del keep
    return CRS(ctx, h, comm.rank, comm.world)


def _wptr(qap, weights, on_device):
    if on_device:
        return C.c_void_p(weights), None
    w = _weights_array(qap, weights)
    return _ptr(w), w


def prove_shard_enqueue(ctx: Context, comm: Comm, qap: QAP, crs: CRS, weights, r: int, s: int, lane=0, on_device=False):
    wp, keep = _wptr(qap, weights, on_device)
    rl, sl = fr_limbs([r]), fr_limbs([s])
    ctx.check(ctx.lib.zkb_prove_shard_enqueue(ctx.h, comm.h, qap.h, crs.h, wp, 1 if on_device else 0, _ptr(rl), _ptr(sl), lane),
              "zkb_prove_shard_enqueue")
    return keep  # the caller keeps host weights alive until collect


def prove_shard_collect(ctx: Context, comm: Comm, lane=0) -> Proof:
    out = _ProofC()
    ctx.check(ctx.lib.zkb_prove_shard_collect(ctx.h, comm.h, lane, C.byref(out)), "zkb_prove_shard_collect")
    return _proof(out)


def prove_shard(ctx: Context, comm: Comm, qap: QAP, crs: CRS, weights, r: int, s: int, on_device=False) -> Proof:
    """groth16::prove (mod.rs:213-296) as ONE proof over all ranks of `comm`; every rank gets the complete proof."""
    wp, keep = _wptr(qap, weights, on_device)
    rl, sl = fr_limbs([r]), fr_limbs([s])
    out = _ProofC()
    ctx.check(ctx.lib.zkb_prove_shard(ctx.h, comm.h, qap.h, crs.h, wp, 1 if on_device else 0, _ptr(rl), _ptr(sl), C.byref(out)),
              "zkb_prove_shard")
    del keep
    return _proof(out)


def prove_shard_batch(ctx: Context, comm: Comm, qap: QAP, crs: CRS, weights, rs, ss, on_device=False) -> list:
    k = len(weights)
    if on_device:
        keep = None
        ptrs = (C.c_void_p * k)(*[C.c_void_p(w) for w in weights])
    else:
        keep = [_weights_array(qap, w) for w in weights]
        ptrs = (C.c_void_p * k)(*[a.ctypes.data for a in keep])
    rl, sl = fr_limbs(list(rs)), fr_limbs(list(ss))
    out = (_ProofC * k)()
    ctx.check(ctx.lib.zkb_prove_shard_batch(ctx.h, comm.h, qap.h, crs.h, ptrs, 1 if on_device else 0, _ptr(rl), _ptr(sl), k, out),
              "zkb_prove_shard_batch")
    del keep
    return [_proof(out[i]) for i in range(k)]


def ntt_shard(ctx: Context, comm: Comm, d_local: int, log_n: int, inverse=False, wait=True):
    """zkb_ntt_shard on this rank's n/world canonical residues (layout D in, layout S out)."""
    ctx.check(ctx.lib.zkb_ntt_shard(ctx.h, comm.h, C.c_void_p(d_local), log_n, 1 if inverse else 0, 0 if wait else 1), "zkb_ntt_shard")


# ------------------------------------------------------------------------------------------------
# wire format (include/zkb200.h: zkb_wire_*; layout in csrc/wire.cu).  Host-only: works without a GPU.
def _wire_check(rc: int, what: str):
    if rc != 0:
        raise ZkbError(f"{what} = {rc}: {load_library().zkb_last_error(None).decode()}")


def _aligned_copy(data: bytes) -> np.ndarray:
    """8-byte aligned copy of `data` (the readers return views into the buffer)."""
    arr = np.zeros((len(data) + 7) // 8, dtype=np.uint64)
    arr.view(np.uint8)[: len(data)] = np.frombuffer(data, dtype=np.uint8)
    return arr


def _qap_host(n, m, n_input, rows, roots=None):
    host, keep = _QapHost(), []
    host.n, host.m, host.n_input = n, m, n_input
    if roots is not None:
        ra = roots if isinstance(roots, np.ndarray) else fr_limbs([r % FR_MODULUS for r in roots])
        ra = np.ascontiguousarray(ra, dtype=np.uint64).reshape(n, 4)
        keep.append(ra)
        host.roots = ra.ctypes.data
    else:
        host.roots = None
    for t, (ptr, gate, coeff) in enumerate(rows):
        a = [np.ascontiguousarray(ptr, dtype=np.uint64), np.ascontiguousarray(gate, dtype=np.uint32),
             np.ascontiguousarray(coeff, dtype=np.uint64)]
        keep += a
        host.row_ptr[t], host.gate[t], host.coeff[t] = (x.ctypes.data for x in a)
    return host, keep


def qap_to_bytes(n: int, m: int, n_input: int, rows, roots=None) -> bytes:
    """Serialise a QAP given as CSR rows (the arguments of QAP(...)): kind 1 record."""
    lib = load_library()
    host, keep = _qap_host(n, m, n_input, rows, roots)
    size = C.c_uint64()
    _wire_check(lib.zkb_wire_size_qap(C.byref(host), C.byref(size)), "zkb_wire_size_qap")
    buf = np.zeros(size.value, dtype=np.uint8)
    _wire_check(lib.zkb_wire_write_qap(C.byref(host), _ptr(buf), size.value), "zkb_wire_write_qap")
    del keep
    return buf.tobytes()


def qap_from_bytes(data: bytes):
    """-> (n, m, n_input, rows, roots or None) with numpy arrays copied out of the record."""
    lib = load_library()
    buf = _aligned_copy(data)
    host = _QapHost()
    _wire_check(lib.zkb_wire_read_qap(_ptr(buf), len(data), C.byref(host)), "zkb_wire_read_qap")
    n, m = int(host.n), int(host.m)
    rows = []
    for t in range(3):
        ptr = np.ctypeslib.as_array(C.cast(host.row_ptr[t], C.POINTER(C.c_uint64)), shape=(m + 1,)).copy()
        nnz = int(ptr[m])
        gate = np.ctypeslib.as_array(C.cast(host.gate[t], C.POINTER(C.c_uint32)), shape=(max(nnz, 1),))[:nnz].copy()
        coeff = np.ctypeslib.as_array(C.cast(host.coeff[t], C.POINTER(C.c_uint64)), shape=(max(nnz, 1) * 4,))[: nnz * 4].reshape(nnz, 4).copy()
        rows.append((ptr, gate, coeff))
    roots = None
    if host.roots:
        roots = np.ctypeslib.as_array(C.cast(host.roots, C.POINTER(C.c_uint64)), shape=(n * 4,)).reshape(n, 4).copy()
    return n, m, int(host.n_input), rows, roots


def qap_upload_bytes(ctx: Context, data: bytes) -> QAP:
    """Upload a serialised QAP (zkb_wire_read_qap view -> zkb_qap_upload, no intermediate copies of the arrays)."""
    buf = _aligned_copy(data)
    host = _QapHost()
    _wire_check(ctx.lib.zkb_wire_read_qap(_ptr(buf), len(data), C.byref(host)), "zkb_wire_read_qap")
    q = QAP.__new__(QAP)
    q.ctx, q.n, q.m, q.input, q.degree = ctx, int(host.n), int(host.m), int(host.n_input), int(host.n)
    h = C.c_void_p()
    ctx.check(ctx.lib.zkb_qap_upload(ctx.h, C.byref(host), C.byref(h)), "zkb_qap_upload")
    q.h = h
    return q


_CRS_SHAPES = (("alpha1", 8), ("beta1", 8), ("delta1", 8), ("xi1", 8), ("xi_t", 8), ("sum_gamma", 8), ("sum_delta", 8),
               ("beta2", 16), ("gamma2", 16), ("delta2", 16), ("xi2", 16))


def crs_raw_to_bytes(raw: dict) -> bytes:
    """Serialise a CRS given as limb arrays in the zkb_crs_host layout (CRS.download_raw()): kind 2 record."""
    lib = load_library()
    host = _CrsHost()
    host.n, host.n_sum_gamma, host.n_sum_delta = raw["xi1"].shape[0], raw["sum_gamma"].shape[0], raw["sum_delta"].shape[0]
    keep = {k: np.ascontiguousarray(raw[k], dtype=np.uint64) for k, _ in _CRS_SHAPES}
    for k, a in keep.items():
        setattr(host, k, a.ctypes.data if a.size else None)
    size = C.c_uint64()
    _wire_check(lib.zkb_wire_size_crs(C.byref(host), C.byref(size)), "zkb_wire_size_crs")
    buf = np.zeros(size.value, dtype=np.uint8)
    _wire_check(lib.zkb_wire_write_crs(C.byref(host), _ptr(buf), size.value), "zkb_wire_write_crs")
    return buf.tobytes()


def crs_raw_from_bytes(data: bytes) -> dict:
    lib = load_library()
    buf = _aligned_copy(data)
    host = _CrsHost()
    _wire_check(lib.zkb_wire_read_crs(_ptr(buf), len(data), C.byref(host)), "zkb_wire_read_crs")
    n, ng, nd = int(host.n), int(host.n_sum_gamma), int(host.n_sum_delta)
    counts = {"alpha1": 1, "beta1": 1, "delta1": 1, "xi1": n, "xi_t": max(n - 1, 0), "sum_gamma": ng, "sum_delta": nd,
              "beta2": 1, "gamma2": 1, "delta2": 1, "xi2": n}
    out = {}
    for k, w in _CRS_SHAPES:
        c = counts[k]
        p = getattr(host, k)
        out[k] = (np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint64)), shape=(c * w,)).reshape(c, w).copy() if c
                  else np.zeros((0, w), dtype=np.uint64))
    return out


def crs_upload_bytes(ctx: Context, data: bytes, rank=0, world=1, comm=None) -> CRS:
    """Upload a serialised CRS (whole, a contiguous shard, or -- with `comm` -- this rank's shard of a sharded proof)."""
    buf = _aligned_copy(data)
    host = _CrsHost()
    _wire_check(ctx.lib.zkb_wire_read_crs(_ptr(buf), len(data), C.byref(host)), "zkb_wire_read_crs")
    h = C.c_void_p()
    if comm is not None:
        ctx.check(ctx.lib.zkb_crs_upload_shard(ctx.h, comm.h, C.byref(host), C.byref(h)), "zkb_crs_upload_shard")
        return CRS(ctx, h, comm.rank, comm.world)
    ctx.check(ctx.lib.zkb_crs_upload(ctx.h, C.byref(host), rank, world, C.byref(h)), "zkb_crs_upload")
    return CRS(ctx, h, rank, world)


def proof_to_bytes(proof: Proof) -> bytes:
    pc = _proof_c(proof)
    buf = np.zeros(WIRE_PROOF_BYTES, dtype=np.uint8)
    _wire_check(load_library().zkb_wire_write_proof(C.byref(pc), _ptr(buf), WIRE_PROOF_BYTES), "zkb_wire_write_proof")
    return buf.tobytes()


def proof_from_bytes(data: bytes) -> Proof:
    buf = _aligned_copy(data)
    pc = _ProofC()
    _wire_check(load_library().zkb_wire_read_proof(_ptr(buf), len(data), C.byref(pc)), "zkb_wire_read_proof")
    return _proof(pc)


def wire_kind(data: bytes) -> int:
    """1 QAP, 2 CRS, 3 Proof (after the header / length / checksum checks)."""
    buf = _aligned_copy(data)
    k, total = C.c_int(), C.c_uint64()
    _wire_check(load_library().zkb_wire_kind(_ptr(buf), len(data), C.byref(k), C.byref(total)), "zkb_wire_kind")
    return k.value


# ------------------------------------------------------------------------------------------------
# witness generation on the device: `weights()` / `evaluate()` (circuit/mod.rs:529-656)
class WitnessPlan:
    """Levelised evaluation plan of a circuit: which wires the caller supplies (``free_wires``: the program's
    ``(in ...)`` variables as indices into the weight vector) and, per level, the gates whose inputs are known.
    ``program_order=True`` mirrors `weights()` exactly (a gate may only read wires assigned by EARLIER gates,
    circuit/mod.rs:598-621); False accepts any topological order (the builder's `evaluate`, builder/mod.rs:556-580)."""

    def __init__(self, ctx: Context, qap: QAP, free_wires, program_order: bool = True):
        self.ctx, self.qap = ctx, qap  # (the plan is self-contained; qap is kept for its row count)
        fw = np.ascontiguousarray(np.asarray(list(free_wires), dtype=np.int64).astype(np.uint32))
        self.n_free = int(fw.size)
        h = C.c_void_p()
        ctx.check(ctx.lib.zkb_witness_plan_create(ctx.h, qap.h, fw.ctypes.data if fw.size else None, fw.size,
                                                  WITNESS_PROGRAM_ORDER if program_order else 0, C.byref(h)), "zkb_witness_plan_create")
        self.h = h

    def info(self) -> dict:
        v = [C.c_uint64() for _ in range(4)]
        self.ctx.check(self.ctx.lib.zkb_witness_plan_info(self.h, *[C.byref(x) for x in v]), "zkb_witness_plan_info")
        return dict(zip(("n_gates", "n_levels", "max_width", "n_launches"), (int(x.value) for x in v)))

    def free(self):
        if getattr(self, "h", None) and self.ctx.h:
            self.ctx.lib.zkb_witness_plan_free(self.ctx.h, self.h)
        self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def witness_generate_raw(ctx: Context, plan: WitnessPlan, values) -> np.ndarray:
    """-> (m, 4) uint64 canonical limbs (host)."""
    vals = values if isinstance(values, np.ndarray) else fr_limbs([v % FR_MODULUS for v in values])
    vals = np.ascontiguousarray(vals, dtype=np.uint64).reshape(-1, 4)
    out = np.empty((plan.qap.m, 4), dtype=np.uint64)
    ctx.check(ctx.lib.zkb_witness_generate(ctx.h, plan.h, vals.ctypes.data if vals.size else None, vals.shape[0], 0,
                                           out.ctypes.data, 0), "zkb_witness_generate")
    return out


def weights(ctx: Context, plan: WitnessPlan, values) -> list:
    """`groth16::weights(code, values)` (circuit/mod.rs:529-637) with the parse already done: values of the `in`
    variables in plan order -> [1, every wire's assignment] as ints."""
    return limbs_to_ints(witness_generate_raw(ctx, plan, values))


def witness_generate_dev(ctx: Context, plan: WitnessPlan, d_values: int, n_values: int, d_out: int):
    """Values and result in device memory (canonical limbs); d_out feeds prove_dev / prove_batch(on_device=True)."""
    ctx.check(ctx.lib.zkb_witness_generate(ctx.h, plan.h, d_values, n_values, 1, d_out, 1), "zkb_witness_generate")


def layered_qap_rows(width: int, depth: int, fan_in: int = 2, seed: int = 1, unit_coeffs: bool = False):
    """Synthetic WIDE circuit for witness generation (the Horner family is a depth-n chain): `depth` layers of `width`
    gates; gate j of layer l multiplies two sums of `fan_in` wires each, drawn (seeded) from the previous layer's
    outputs -- layer 0 from the `width` free input wires -- with small literal weights (``unit_coeffs``: all 1, the shape of
    a parsed program, where variables and unweighted sums dominate).  Wire order: 0 unity,
    1..width inputs, then the gate outputs in gate order (gate k -> wire width + 1 + k); n = width * depth gates.
    Returns (n, m, n_input, rows, free_wires).  Vectorised: builds 2^20 gates in about a second."""
    rng = np.random.default_rng(seed)
    n = width * depth
    m = 1 + width + n
    gate = np.repeat(np.arange(n, dtype=np.int64), fan_in)
    layer = gate // width
    base = np.where(layer == 0, 1, 1 + width + (layer - 1) * width)  # first wire of the previous layer
    rows = []
    for _ in range(2):  # u, v
        wire = base + rng.integers(0, width, size=n * fan_in)
        coef = np.ones(n * fan_in, dtype=np.uint64) if unit_coeffs else rng.integers(1, 8, size=n * fan_in).astype(np.uint64)
        # merge duplicate (wire, gate) pairs: a by-wire row holds one entry per gate
        key = wire * n + gate
        uniq, inv = np.unique(key, return_inverse=True)
        csum = np.zeros(uniq.size, dtype=np.uint64)
        np.add.at(csum, inv, coef)
        w_u, g_u = uniq // n, uniq % n  # sorted by wire, then gate
        ptr = np.zeros(m + 1, dtype=np.uint64)
        np.add.at(ptr, w_u + 1, 1)
        ptr = np.cumsum(ptr).astype(np.uint64)
        limbs = np.zeros((uniq.size, 4), dtype=np.uint64)
        limbs[:, 0] = csum
        rows.append((ptr, g_u.astype(np.uint32), limbs))
    pw = np.zeros(m + 1, dtype=np.uint64)
    pw[1 + width + 1:] = np.arange(1, n + 1, dtype=np.uint64)
    one = np.zeros((n, 4), dtype=np.uint64)
    one[:, 0] = 1
    rows.append((pw, np.arange(n, dtype=np.uint32), one))
    return n, m, width, rows, list(range(1, width + 1))


def witness_levels(n: int, m: int, n_input: int, rows, free_wires, program_order: bool = True) -> list:
    """Host only (works without a GPU): the level the witness planner gives every gate (0: assigns nothing)."""
    lib = load_library()
    host, keep = _qap_host(n, m, n_input, rows)
    fw = np.ascontiguousarray(np.asarray(list(free_wires), dtype=np.int64).astype(np.uint32))
    out = np.zeros(n, dtype=np.uint32)
    nl = C.c_uint64()
    _wire_check(lib.zkb_witness_levels(C.byref(host), fw.ctypes.data if fw.size else None, fw.size,
                                       WITNESS_PROGRAM_ORDER if program_order else 0, out.ctypes.data, C.byref(nl)), "zkb_witness_levels")
    assert int(out.max(initial=0)) == nl.value
    return out.tolist()
