"""Latency / throughput of groth16::verify on the device (zkb_verify, zkb_verify_batch).
  python tools/verify_bench.py [log_n] [batch sizes ...]"""
import importlib
import os
import random
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
zk = importlib.import_module("zksnark-rs_b200")
zg = importlib.import_module("zksnark-rs_b200.groth16")
P = zg.FR_MODULUS


def main():
    log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    batches = [int(a) for a in sys.argv[2:]] or [1, 32, 1024, 8192]
    n = 1 << log_n
    rng = random.Random(5)
    ctx = zk.Context(0)
    qap = zk.QAP.horner(ctx, n)
    crs = zk.setup(ctx, qap, tuple(rng.randrange(1, P) for _ in range(5)))
    wit = zg.horner_witness(n, rng.randrange(1, P), [rng.randrange(P) for _ in range(n)])
    proof = zk.prove(ctx, qap, crs, wit, 3, 4)
    pub = wit[1:3]
    assert zk.verify(ctx, crs, pub, proof)
    for _ in range(2):
        t0 = time.perf_counter()
        ok = zk.verify(ctx, crs, pub, proof)
        dt = time.perf_counter() - t0
    print(f"verify (1 proof): {dt * 1e3:.2f} ms  ok={ok}", flush=True)
    for b in batches:
        proofs, inputs = [proof] * b, [pub] * b
        zg.verify_batch(ctx, crs, inputs[:1], proofs[:1])
        t0 = time.perf_counter()
        oks = zg.verify_batch(ctx, crs, inputs, proofs)
        dt = time.perf_counter() - t0
        assert all(oks)
        print(f"verify_batch({b}): {dt * 1e3:.1f} ms wall incl. host packing -> {b / dt:.0f} proofs/s", flush=True)


if __name__ == "__main__":
    main()
