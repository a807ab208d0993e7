"""ONE PROOF OVER SEVERAL RANKS on the device (csrc/shard.cu, zkb_prove_shard*, zkb_ntt_shard): the NTT's outer
dimension sharded, exchanges by kernel stores into the peers' windows, partial sums folded on the device.

The ranks of a communicator are one process per GPU (tools/shard_multi_gpu.py and bench.py --gpus N run that on real
multi-GPU boxes).  The GPU test box has ONE GPU, so these tests start world = 2, 4, 8 PROCESSES on it (torchrun, gloo
for the window handles): every rank maps its peers' windows with cudaIpcOpenMemHandle exactly as in a multi-GPU job and
the GPU time-slices the contexts.  tests/shard_worker.py checks every rank's result bit for bit against the one-GPU
proof (pinned to the reference restatement and to Oracle F by tests/test_gpu_parity*.py), against the literal
restatement directly (uploaded CRS) and, for the transform, against Oracle F.
(Several ranks inside one process would be simpler, but a rank's first allocations would then wait for another rank's
exchange-wait kernel: cudaMalloc synchronises the device.)
"""

import importlib
import os
import random
import subprocess
import sys

import pytest

from oracle.fields import FR

pytestmark = pytest.mark.gpu
zk = importlib.import_module("zksnark-rs_b200")
zg = importlib.import_module("zksnark-rs_b200.groth16")
P = FR.p
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _witness(n, rng, valid=True):
    wit = zg.horner_witness(n, rng.randrange(1, P), [rng.randrange(P) for _ in range(n)])
    if not valid:
        wit[3] = (wit[3] + 1) % P
        wit[-1] = rng.randrange(P)
    return zg.fr_limbs(wit)


def _run_ranks(world, cases, port, timeout=600):
    env = dict(os.environ, ZKB_COMM_TIMEOUT_MS="30000")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "shard_worker.py"), ROOT, "same"] + cases
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert out.stdout.count("ok") == world


def test_shard_world2():
    """n = 2*world^2 (the smallest allowed) up to 2^16; lanes in flight; uploaded reference CRS; transforms to 2^18."""
    _run_ranks(2, ["prove:3", "prove:4", "prove:10", "prove:16", "lanes:6", "lanes:13", "upload:4", "ntt:3", "ntt:14", "ntt:18"], 29614)


def test_shard_world4():
    _run_ranks(4, ["prove:5", "prove:6", "prove:12", "lanes:5", "lanes:12", "upload:5", "ntt:5", "ntt:9", "ntt:18"], 29615)


def test_shard_world8():
    _run_ranks(8, ["prove:7", "prove:14", "lanes:10", "ntt:7", "ntt:12", "ntt:16"], 29616)


def test_shard_argument_checks():
    ctx = zk.Context(0)
    try:
        with pytest.raises(zk.ZkbError):
            zk.Comm.create(ctx, 0, 3, 10)       # world must be a power of two
        with pytest.raises(zk.ZkbError):
            zk.Comm.create(ctx, 0, 4, 4)        # n >= 2 world^2
        cm = zk.Comm.create(ctx, 0, 2, 8)
        q = zk.QAP.horner(ctx, 16)
        crs = zk.setup(ctx, q, (3, 5, 7, 11, 13))
        with pytest.raises(zk.ZkbError, match="not connected"):
            zk.prove_shard(ctx, cm, q, crs, [1] * q.m, 5, 7)
        with pytest.raises(zk.ZkbError):
            cm.connect(cm.handle * 2)           # second handle does not describe rank 1
        one = zk.Comm.create(ctx, 0, 1, 8)      # world 1: connected by construction; same code path, no peers
        crs1 = zk.setup_shard(ctx, one, q, (3, 5, 7, 11, 13))
        w = _witness(16, random.Random(5))
        a = zk.prove_shard(ctx, one, q, crs1, w, 5, 7)
        b = zk.prove(ctx, q, crs, w, 5, 7)
        assert (a.a, a.b, a.c) == (b.a, b.b, b.c)
        with pytest.raises(zk.ZkbError, match="zkb_setup_shard"):
            zk.prove_shard(ctx, one, q, crs, w, 5, 7)   # an unsharded-layout CRS
    finally:
        ctx.close()


def test_exchange_timeout_is_an_error_not_a_hang(monkeypatch):
    """A rank whose peer never arrives fails with ZKB_ERR_COMM after the timeout instead of spinning forever."""
    monkeypatch.setenv("ZKB_COMM_TIMEOUT_MS", "200")
    ctxs = [zk.Context(0), zk.Context(0)]
    comms = [zk.Comm.create(c, r, 2, 6) for r, c in enumerate(ctxs)]  # two contexts of this process: plain pointers, no IPC
    try:
        for cm in comms:
            cm.connect([c.handle for c in comms])
        n = 64
        q = zk.QAP.horner(ctxs[0], n)
        crs = zk.setup_shard(ctxs[0], comms[0], q, (3, 5, 7, 11, 13))
        with pytest.raises(zk.ZkbError, match="did not arrive"):
            zk.prove_shard(ctxs[0], comms[0], q, crs, _witness(n, random.Random(1)), 5, 7)  # rank 1 never calls
        assert comms[0].status() == 1
    finally:
        for cm in comms:
            cm.free()
        for c in ctxs:
            c.close()
