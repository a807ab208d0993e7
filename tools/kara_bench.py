"""A/B of the Fq multiplications (CIOS, wide product + reduction, Karatsuba wide product) on the device:
results against Python integers, then the modmul rate of each (profiles/r02_notes.md)."""
import importlib, random, sys, os
sys.path.insert(0, os.getcwd())
zk = importlib.import_module("zksnark-rs_b200"); zg = importlib.import_module("zksnark-rs_b200.groth16")
ctx = zk.Context(0)
p = 21888242871839275222246405745257275088696311157297823662689037894645226208583  # Fq modulus (python integers are the check)
rng = random.Random(1)
edge = [0, 1, p - 1, p - 2, (1 << 253), (1 << 128) - 1, ((1 << 125) - 1) << 128, (1 << 128), (1<<253) | ((1<<128)-1)]
edge = [e % p for e in edge]
a = [x for x in edge for _ in edge] + [rng.randrange(p) for _ in range(4096)]
b = [y for _ in edge for y in edge] + [rng.randrange(p) for _ in range(4096)]
want = [x * y % p for x, y in zip(a, b)]
for op in (0, 6, 7):
    print("op", op, zg.field_op(ctx, 1, op, a, b) == want, flush=True)
for f in (1, 2, 3, 1, 2, 3):
    print("field", f, ctx.bench_modmul(f, 2000), flush=True)
