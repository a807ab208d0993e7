"""CPU tests of the ONE-PROOF-OVER-G-RANKS polynomial stage (tests/shard_model.py = the steps csrc/shard.cu runs on the
device): index conventions of the two layouts, the distributed transforms against the reference's dft
(field/mod.rs:508-537) and u_sum / v_sum / h against the literal restatement of mod.rs:233-253, 277 -- in one process
for world 2, 4, 8 and over a real world-size-2 gloo group (torch.distributed all_to_all)."""

import importlib
import os
import random
import subprocess
import sys

import numpy as np
import pytest

import shard_model as sm
from oracle import groth16 as og, poly, synthetic
from oracle.fields import FR

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
zd = importlib.import_module("zksnark-rs_b200.dist")
P = FR.p


def test_layout_maps_partition_and_match_the_model():
    for n, G in ((16, 2), (64, 4), (128, 8), (1 << 12, 4)):
        seen_s, seen_d = [], []
        for r in range(G):
            s_idx = zd.layout_s_index(n, r, G)
            assert list(s_idx[:5]) == [sm.s_to_j(s, r, G, n) for s in range(5)]
            assert int(s_idx[-1]) == sm.s_to_j(n // G - 1, r, G, n)
            seen_s += list(s_idx)
            seen_d += list(zd.layout_d_index(n, r, G))
        assert sorted(seen_s) == list(range(n)) and sorted(seen_d) == list(range(n))
        # coefficient n-1 (absent from xi_t) is the last local index of the last rank
        assert int(zd.layout_s_index(n, G - 1, G)[-1]) == n - 1
    with pytest.raises(ValueError):
        zd.layout_s_index(8, 0, 4)


@pytest.mark.parametrize("n,G", [(16, 2), (16, 4), (128, 8)])
def test_distributed_transforms_equal_reference_dft(n, G):
    rng = random.Random(n + G)
    w = sm.omega(n.bit_length() - 1)
    x = [rng.randrange(P) for _ in range(n)]
    want = poly.dft(FR, x, w)
    res = sm.run_all_ranks(lambda r, ex: sm.dit_distributed(x[r::G], r, G, n, w, ex), G)
    for r in range(G):
        assert res[r] == [want[j] for j in zd.layout_s_index(n, r, G)]
    xs = [[x[j] for j in zd.layout_s_index(n, r, G)] for r in range(G)]
    res = sm.run_all_ranks(lambda r, ex: sm.dif_distributed(xs[r], r, G, n, w, ex), G)
    for r in range(G):
        assert res[r] == want[r::G]


def _horner_case(n, rng, valid):
    w = sm.omega(n.bit_length() - 1)
    roots = [pow(w, k, P) for k in range(n)]
    rep = synthetic.horner_rep(FR, n, roots)
    wit = synthetic.horner_witness(FR, n, rng.randrange(1, P), [rng.randrange(P) for _ in range(n)])
    if not valid:
        wit[5] = (wit[5] + 1) % P
    dense = og.qap_from_root_rep(FR, rep)
    us, vs, ws = og.weighted_sums(FR, dense, wit)
    h = og.quotient_h(FR, dense, us, vs, ws)
    pad = lambda v: list(v) + [0] * (n - len(v))
    idx = {r_: k for k, r_ in enumerate(roots)}
    A, B = [0] * n, [0] * n
    for i, row in enumerate(rep.u):
        for rt, c in row:
            A[idx[rt]] = (A[idx[rt]] + c * wit[i]) % P
    for i, row in enumerate(rep.v):
        for rt, c in row:
            B[idx[rt]] = (B[idx[rt]] + c * wit[i]) % P
    return A, B, pad(us), pad(vs), pad(h)


@pytest.mark.parametrize("n,G", [(16, 2), (32, 4), (128, 8)])
@pytest.mark.parametrize("valid", [True, False])
def test_sharded_poly_stage_equals_reference_quotient(n, G, valid):
    A, B, us, vs, h = _horner_case(n, random.Random(3 * n + G + valid), valid)
    res = sm.run_all_ranks(lambda r, ex: sm.poly_stage_rank(A[r::G], B[r::G], r, G, n, ex), G)
    for r in range(G):
        js = zd.layout_s_index(n, r, G)
        u, v, hh = res[r]
        assert u == [us[j] for j in js] and v == [vs[j] for j in js] and hh == [h[j] for j in js]


_GLOO_SHARD_WORKER = r"""
import importlib, os, sys, random
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import shard_model as sm
from test_shard_model import _horner_case
zd = importlib.import_module("zksnark-rs_b200.dist")
zg = importlib.import_module("zksnark-rs_b200.groth16")
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
def exchange(send):
    gathered = [None] * world
    dist.all_gather_object(gathered, send)   # gathered[g][d] = what rank g sends to rank d
    return [gathered[g][rank] for g in range(world)]
n = 32
A, B, us, vs, h = _horner_case(n, random.Random(11), valid=(sys.argv[2] == "1"))
u, v, hh = sm.poly_stage_rank(A[rank::world], B[rank::world], rank, world, n, exchange)
js = zd.layout_s_index(n, rank, world)
assert u == [us[j] for j in js] and v == [vs[j] for j in js] and hh == [h[j] for j in js]
dist.destroy_process_group()
print("ok", rank)
"""


@pytest.mark.parametrize("valid", ["1", "0"])
def test_sharded_poly_stage_world2_gloo(tmp_path, valid):
    """The N>1 protocol between real processes: world size 2, gloo, rendezvous on 127.0.0.1."""
    script = tmp_path / "worker_shard.py"
    script.write_text(_GLOO_SHARD_WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29613", str(script), ROOT, valid]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("ok") == 2
