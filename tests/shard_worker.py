"""Worker of tests/test_gpu_shard.py: one rank of a multi-process communicator (torchrun; gloo carries the window
handles).  All ranks may sit on the same GPU (the test box has one): the peers' windows are then mapped through CUDA IPC
exactly as between the processes of a multi-GPU job, and the GPU time-slices the contexts.  Test infrastructure (it
imports the oracle); the product path is zksnark-rs_b200 + libzkb200.

usage: shard_worker.py <repo root> <device: 'same' | 'local'> <case> [<case> ...]
  case = prove:<log_n> | lanes:<log_n> | upload:<log_n> | ntt:<log_n>
"""
import importlib
import os
import random
import sys

import numpy as np
import torch.distributed as dist

ROOT = sys.argv[1]
sys.path.insert(0, ROOT)
zk = importlib.import_module("zksnark-rs_b200")
zg = importlib.import_module("zksnark-rs_b200.groth16")
zd = importlib.import_module("zksnark-rs_b200.dist")
from oracle import groth16 as og, oracle_fast as of, synthetic  # noqa: E402
from oracle.fields import FR  # noqa: E402

P = FR.p


def witness(n, rng, valid=True):
    wit = zg.horner_witness(n, rng.randrange(1, P), [rng.randrange(P) for _ in range(n)])
    if not valid:
        wit[3] = (wit[3] + 1) % P
        wit[-1] = rng.randrange(P)
    return zg.fr_limbs(wit)


def same(a, b):
    return (a.a, a.b, a.c) == (b.a, b.b, b.c)


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = 0 if sys.argv[2] == "same" else int(os.environ.get("LOCAL_RANK", "0"))
    cases = [c.split(":") for c in sys.argv[3:]]
    max_log = max(int(c[1]) for c in cases)
    ctx = zk.Context(dev)
    comm = zd.connect(ctx, max_log)
    for kind, lg in cases:
        log_n = int(lg)
        n = 1 << log_n
        rng = random.Random(1000 * log_n + world)  # the same inputs on every rank
        toxic = tuple(rng.randrange(1, P) for _ in range(5))
        if kind == "prove":  # sharded == one GPU, valid and non-satisfying witness, host and device witness
            qap = zk.QAP.horner(ctx, n)
            crs = zk.setup_shard(ctx, comm, qap, toxic)
            crs1 = zk.setup(ctx, qap, toxic)
            for valid in (True, False):
                w = witness(n, rng, valid)
                r, s = rng.randrange(1, P), rng.randrange(1, P)
                want = zk.prove(ctx, qap, crs1, w, r, s)
                got = zk.prove_shard(ctx, comm, qap, crs, w, r, s)
                assert same(got, want), (kind, log_n, valid, "host witness")
                d_w = ctx.dev_alloc(w.nbytes)
                ctx.h2d(d_w, w)
                got = zk.prove_shard(ctx, comm, qap, crs, d_w, r, s, on_device=True)
                assert same(got, want), (kind, log_n, valid, "device witness")
                ctx.dev_free(d_w)
                if valid:
                    assert zk.verify(ctx, crs1, [int.from_bytes(w[i].tobytes(), "little") for i in (1, 2)], got)
            crs.free(); crs1.free(); qap.free()
        elif kind == "lanes":  # several sharded proofs in flight (one lane / exchange channel each), twice
            qap = zk.QAP.horner(ctx, n)
            crs = zk.setup_shard(ctx, comm, qap, toxic)
            crs1 = zk.setup(ctx, qap, toxic)
            for _round in range(2):
                ws = [witness(n, rng) for _ in range(7)]
                rs = [rng.randrange(1, P) for _ in range(7)]
                ss = [rng.randrange(1, P) for _ in range(7)]
                want = zk.prove_batch(ctx, qap, crs1, ws, rs, ss)
                got = zk.prove_shard_batch(ctx, comm, qap, crs, ws, rs, ss)
                assert all(same(g, w_) for g, w_ in zip(got, want)), (kind, log_n)
                keep = [zk.prove_shard_enqueue(ctx, comm, qap, crs, ws[l], rs[l], ss[l], lane=l) for l in range(4)]
                for l in (3, 1, 0, 2):  # collected in any order
                    assert same(zk.prove_shard_collect(ctx, comm, lane=l), want[l]), (kind, log_n, l)
                del keep
            crs.free(); crs1.free(); qap.free()
        elif kind == "upload":  # CRS made by the literal restatement of setup (mod.rs:134-197), uploaded shard by shard
            w_ = synthetic.omega(log_n)
            rep = synthetic.horner_rep(FR, n, [pow(w_, k, P) for k in range(n)])
            dense = og.qap_from_root_rep(FR, rep)
            B = og.BN254Backend()
            sig = og.setup(B, dense, toxic)
            qap = zk.QAP.horner(ctx, n)
            crs = zk.crs_upload_shard(ctx, comm, sig[0], sig[1])
            wit = synthetic.horner_witness(FR, n, rng.randrange(1, P), [rng.randrange(P) for _ in range(n)])
            r, s = rng.randrange(1, P), rng.randrange(1, P)
            want = og.prove(B, dense, sig, wit, r, s)  # the literal restatement of mod.rs:213-296
            got = zk.prove_shard(ctx, comm, qap, crs, wit, r, s)
            assert (got.a, got.b, got.c) == (want.a, want.b, want.c), (kind, log_n)
            crs.free(); qap.free()
        elif kind == "ntt":  # zkb_ntt_shard == Oracle F (pinned to the reference's dft / idft), forward and inverse, repeated
            g = np.random.default_rng(log_n)
            x = g.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64)
            x[:, 3] &= np.uint64((1 << 60) - 1)
            loc = np.ascontiguousarray(x[zd.layout_d_index(n, rank, world)])
            d = ctx.dev_alloc(loc.nbytes)
            for inverse in (False, True):
                want = of.ntt_np(x, inverse=inverse)[zd.layout_s_index(n, rank, world)]
                for rep in range(3):
                    ctx.h2d(d, loc)
                    zk.ntt_shard(ctx, comm, d, log_n, inverse=inverse, wait=(rep != 1))
                    ctx.sync()
                    out = np.zeros_like(loc)
                    ctx.d2h(out, d)
                    assert np.array_equal(out, want), (kind, log_n, inverse, rep)
            ctx.dev_free(d)
        else:
            raise SystemExit(f"unknown case {kind}")
        assert comm.status() == 0, (kind, log_n, "exchange status")
        dist.barrier()
    comm.free()
    ctx.close()
    dist.destroy_process_group()
    print("ok", rank, flush=True)


if __name__ == "__main__":
    main()
