"""CPU-side checks of the product's host logic and of the C ABI surface (no compute calls: there
is no GPU here and the library has no CPU fallback)."""

import ctypes
import importlib
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from oracle import synthetic
from oracle.fields import FR

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
zk = importlib.import_module("zksnark-rs_b200")
zg = importlib.import_module("zksnark-rs_b200.groth16")
zd = importlib.import_module("zksnark-rs_b200.dist")


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(zk.lib_path()):
        importlib.import_module("zksnark-rs_b200.build").build()
    return zk.load_library()


def test_library_exports_every_declared_symbol(lib):
    header = open(os.path.join(ROOT, "include", "zkb200.h")).read()
    declared = set(re.findall(r"\b(zkb_[a-z0-9_]+)\s*\(", header))
    assert declared == set(zg.ABI), declared ^ set(zg.ABI)
    for name in declared:
        assert getattr(lib, name) is not None


def test_rust_extern_block_covers_the_whole_header():
    """integration/rust/zkb200_sys.rs is generated from include/zkb200.h (tools/gen_rust_sys.py): the committed file is
    up to date and declares exactly the header's functions, with as many parameters each."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    gen = importlib.import_module("gen_rust_sys")
    text, names = gen.generate()
    assert open(os.path.join(ROOT, "integration", "rust", "zkb200_sys.rs")).read() == text, "run python tools/gen_rust_sys.py"
    header = open(os.path.join(ROOT, "include", "zkb200.h")).read()
    declared = set(re.findall(r"\b(zkb_[a-z0-9_]+)\s*\(", gen.strip_comments(header)))
    rust = dict(re.findall(r"pub fn (zkb_[a-z0-9_]+)\(([^)]*)\)", text))
    assert set(rust) == declared == set(zg.ABI)
    for name, args in rust.items():
        assert len([a for a in args.split(",") if a.strip()]) == len(zg.ABI[name][1]), name


def test_no_cpu_fallback(lib):
    """Without a device every computing entry point must fail loudly, not fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = ctypes.c_void_p()
    assert lib.zkb_ctx_create(ctypes.byref(h), 0) == -1
    assert b"no CPU fallback" in lib.zkb_last_error(None)
    with pytest.raises(zk.ZkbError):
        zk.Context(0)


def test_product_never_imports_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may touch oracle/: not the package, not the headers,
    not the developer probes under tools/, not the reference-side bindings."""
    for top in ("zksnark-rs_b200", "include", "tools", "integration"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".rs", ".sh")):
                    src = open(os.path.join(dirpath, f)).read()
                    assert "import oracle" not in src and "from oracle" not in src and "oracle_b" not in src \
                        and "oracle_fast" not in src, os.path.join(top, f)


def test_limb_packing_roundtrip():
    vals = [0, 1, FR.p - 1, 2**64, 2**200 + 12345]
    arr = zg.fr_limbs(vals)
    assert arr.shape == (5, 4) and arr.dtype == np.uint64
    assert zg.limbs_to_ints(arr) == vals
    pts = [None, (3, 4)]
    assert zg.g1_unpack(zg.g1_pack(pts)) == pts
    q = [None, ((1, 2), (3, 4))]
    assert zg.g2_unpack(zg.g2_pack(q)) == q


@pytest.mark.parametrize("n", [2, 4, 8, 64])
def test_horner_rows_equal_oracle_rep(n):
    """The numpy CSR generator of the synthetic circuit == the oracle's DummyRep (parser row order)."""
    log_n = n.bit_length() - 1
    w = synthetic.omega(log_n)
    roots = [pow(w, k, FR.p) for k in range(n)]
    rep = synthetic.horner_rep(FR, n, roots)
    idx = {r: k for k, r in enumerate(roots)}
    m, n_input, rows = zg.horner_qap_rows(n)
    assert m == len(rep.u) and n_input == rep.input
    for (ptr, gate, coeff), mat in zip(rows, (rep.u, rep.v, rep.w)):
        cs = zg.limbs_to_ints(coeff)
        for i in range(m):
            got = sorted((int(gate[e]), cs[e]) for e in range(int(ptr[i]), int(ptr[i + 1])))
            assert got == sorted((idx[r], c) for r, c in mat[i])
    x, cs_ = 12345, list(range(7, 7 + n))
    assert zg.horner_witness(n, x, cs_) == synthetic.horner_witness(FR, n, x, cs_)
    assert zg.omega(log_n) == w


def test_shard_ranges_partition():
    for length in (0, 1, 7, 1024, (1 << 20) - 1):
        for world in (1, 2, 3, 8):
            spans = [zd.shard_range(length, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == length
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


def test_bench_mirrors_of_library_plans():
    sys.path.insert(0, ROOT)
    bench = importlib.import_module("bench")
    assert bench.ntt_plan(20) == [(0, 10), (10, 15), (15, 20)]
    assert bench.ntt_plan(8) == [(0, 8)]
    assert bench.msm_window((1 << 22) + 1) == 20 and bench.msm_window((1 << 20) + 2) == 17  # c1, c2 at 2^20
    assert 5.0 < bench.table_bytes(20) / 2**30 < 6.5 and bench.table_bytes(16) > 126e6


def test_reference_arm_line_and_best_effort_leg():
    """bench.py --impl reference (the reference's algorithm, Oracle B, sampled) prints the contract line without a GPU;
    the cpu_best_effort leg (Oracle F) reports a full small proof.  Small size so the CPU suite stays short."""
    import json
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--log-n", "10", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=240)
    assert out.returncode == 0, out.stdout + out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "groth16_proofs_per_sec" and line["unit"] == "proofs/s"
    assert line["value"] > 0 and line["higher_is_better"] is True and line["cpu_baseline"]["kind"] == "port"
    assert line["cpu_baseline"]["cores"] == 1 and line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]
    full = line["config"]["measured_full"]  # the same port run to completion at a size where it finishes
    assert full["log_n"] == 10 and 0.2 < full["seconds_per_proof"] < 120
    sys.path.insert(0, ROOT)
    bench = importlib.import_module("bench")
    fast = bench.best_effort_cpu(8)
    assert fast["kind"] == "port-fast-algorithms" and fast["value"] > line["value"] and fast["cores"] >= 1


_GLOO_WORKER = r"""
import importlib, os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, sys.argv[1])
zd = importlib.import_module("zksnark-rs_b200.dist")
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
part = (np.arange(32, dtype=np.uint64) + np.uint64(1000 * rank)) | np.uint64(1 << 63)  # exercise the sign bit
allp = zd.all_gather_partials(part)
assert allp.shape == (world, 32)
for r in range(world):
    assert np.array_equal(allp[r], (np.arange(32, dtype=np.uint64) + np.uint64(1000 * r)) | np.uint64(1 << 63))
dist.destroy_process_group()
print("ok", rank)
"""


def test_all_gather_partials_world2_gloo(tmp_path):
    """N>1 host path on CPU: world_size 2, gloo backend, rendezvous on 127.0.0.1."""
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29611", str(script), ROOT]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("ok") == 2


_GLOO_NTT_WORKER = """
import importlib, os, sys, random
import numpy as np
import torch.distributed as dist
sys.path.insert(0, sys.argv[1])
zd = importlib.import_module("zksnark-rs_b200.dist")
zg = importlib.import_module("zksnark-rs_b200.groth16")
from oracle import poly, synthetic          # the checker plays the two device steps (local transform, combine)
from oracle.fields import FR
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
P, log_n = FR.p, 6
n = 1 << log_n
rng = random.Random(7)                      # the same vector on every rank
x = [rng.randrange(P) for _ in range(n)]
w = synthetic.omega(log_n)
sl, k0, count = zd.ntt_shard_layout(log_n, rank, world)
assert (k0, count) == (rank * n // world, n // world)
y_mine = poly.dft(FR, x[sl], pow(w, world, P))                     # size n/G transform of the decimated subsequence
parts = zd.all_gather_limbs(zg.fr_limbs(y_mine))                   # (world, n/G, 4), rank order
ys = [zg.limbs_to_ints(parts[g]) for g in range(world)]
mine = [sum(pow(w, g * k, P) * ys[g][k % count] for g in range(world)) % P for k in range(k0, k0 + count)]
full = zd.all_gather_limbs(zg.fr_limbs(mine))
got = [v for g in range(world) for v in zg.limbs_to_ints(full[g])]
assert got == poly.dft(FR, x, w)
dist.destroy_process_group()
print("ok", rank)
"""


def test_ntt_sharded_protocol_world2_gloo(tmp_path):
    """The exchange and index conventions of the outer-dimension-sharded transform (dist.ntt_shard_layout,
    all_gather_limbs) on CPU with world_size 2 over gloo; the oracle stands in for the two device steps."""
    script = tmp_path / "worker_ntt.py"
    script.write_text(_GLOO_NTT_WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29612", str(script), ROOT]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("ok") == 2


def test_ntt_twiddle_tile_layout_matches_flat_table_index():
    """ntt.cu: the slot a butterfly reads from its block's TMA-staged twiddle tile (k_ntt_pass<.., true>) holds the same
    power of omega as the flat-table gather (k_ntt_pass<.., false>), for every pass of every plan (Python restatement
    of k_fill_tw_tiles and of the two branches of the `twiddle` lambda; the GPU suite checks the transforms bit for bit)."""
    import random
    sys.path.insert(0, ROOT)
    bench = importlib.import_module("bench")
    rng = random.Random(3)
    for log_n in (1, 2, 3, 5, 10, 11, 13, 16, 18, 20, 22, 26):
        for (lo, hi) in bench.ntt_plan(log_n):
            B = hi - lo
            logC = 0 if lo == 0 else 10 - B
            T, C = 1 << (B + logC), 1 << logC
            groups = (1 << lo) >> logC

            def tile_exponent(g, slot):  # k_fill_tw_tiles
                assert slot < T - C
                idx, c = slot >> logC, slot & (C - 1)
                ls = (idx + 1).bit_length() - 1
                m0 = idx + 1 - (1 << ls)
                e = (((m0 << lo) + (g << logC) + c) << (log_n - 1 - lo - ls))
                assert e < 1 << (log_n - 1)
                return e
            for _ in range(300):
                g, ls, e = rng.randrange(groups), rng.randrange(B), rng.randrange(T)
                e &= ~(1 << (ls + logC))  # lower element of a butterfly of stage ls: row bit ls clear
                s = lo + ls
                m0 = (e >> logC) & ((1 << ls) - 1)
                j = (m0 << lo) + (g << logC) + (e & (C - 1))
                flat = (0 if s == 0 else j & ((1 << s) - 1)) << (log_n - 1 - s)
                slot = (((1 << ls) - 1) << logC) + (e & ((1 << (ls + logC)) - 1))
                assert tile_exponent(g, slot) == flat
