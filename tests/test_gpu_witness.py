"""GPU parity tests of witness generation on the device: zkb_witness_plan_create / zkb_witness_generate
against `weights()` / `evaluate()` (/root/reference/src/groth16/circuit/mod.rs:529-656) as restated by the oracle
(oracle/circuit.py: `weights` on program text, `weights_from_rows` on the DummyRep rows).  Bit-exact on canonical
residues; the reference's own golden vector (`weights_test`, circuit/mod.rs:746-769) is reproduced on the device."""

import importlib
import random
import time

import numpy as np
import pytest

from oracle import circuit, synthetic
from oracle.fields import FR
from test_oracle_kats import QUAD, SIMPLE
from test_oracle_witness import MIXED, TWO_GATES, swapped_two_gates

pytestmark = pytest.mark.gpu
zk = importlib.import_module("zksnark-rs_b200")
zg = importlib.import_module("zksnark-rs_b200.groth16")
P = FR.p


@pytest.fixture(scope="module")
def ctx():
    c = zk.Context(0)
    yield c
    c.close()


def test_reference_weights_vector(ctx):
    """circuit/mod.rs:746-769: inputs a, b, c = 3, 2, 4 -> [1, b, x, temp, a, c] = [1, 2, 34, 6, 3, 4] (no reduction
    happens on these values, so the Z251 vector is also the BN254 one)."""
    rep = circuit.try_parse(FR, SIMPLE)
    qap = zk.QAP.from_root_representation(ctx, rep)
    plan = zk.WitnessPlan(ctx, qap, circuit.input_wires(FR, SIMPLE))
    assert zk.weights(ctx, plan, [3, 2, 4]) == [1, 2, 34, 6, 3, 4]
    assert plan.info() == {"n_gates": 2, "n_levels": 2, "max_width": 1, "n_launches": 3}


@pytest.mark.parametrize("reindex", [False, True])
@pytest.mark.parametrize("name,text,n_in", [("simple", SIMPLE, 3), ("quad", QUAD, 4), ("mixed", MIXED, 3),
                                            ("deg_15", synthetic.horner_program_text(16), 17),
                                            ("deg_299", synthetic.horner_program_text(300), 301)])
def test_weights_match_oracle_on_parser_circuits(ctx, name, text, n_in, reindex):
    """The reference's circuits (simple.zk of lib.rs:156-190, the quad share of fr.rs:273-302, deg_15 of fr.rs:361-416)
    as its parser emits them (roots 1..=n) and re-indexed onto the roots of unity (padding gates are skipped)."""
    rng = random.Random(len(text))
    rep = circuit.try_parse(FR, text)
    qap = zk.QAP.from_root_representation(ctx, rep, reindex=reindex)
    plan = zk.WitnessPlan(ctx, qap, circuit.input_wires(FR, text))
    assert plan.info()["n_gates"] == len(rep.roots)
    for _ in range(3):
        vals = [rng.randrange(P) for _ in range(n_in)]
        assert zk.weights(ctx, plan, vals) == circuit.weights(FR, text, vals)
    assert zk.weights(ctx, plan, [0] * n_in) == circuit.weights(FR, text, [0] * n_in)
    assert zk.weights(ctx, plan, [P - 1] * n_in) == circuit.weights(FR, text, [P - 1] * n_in)


def test_generated_witness_proves_and_verifies(ctx):
    """weights -> prove -> verify as in the reference's end-to-end tests (fr.rs:361-416), the witness never leaving
    the device between generation and proof."""
    n = 16
    text = synthetic.horner_program_text(n)
    rng = random.Random(5)
    rep = circuit.try_parse(FR, text)
    qap = zk.QAP.from_root_representation(ctx, rep)
    plan = zk.WitnessPlan(ctx, qap, circuit.input_wires(FR, text))
    vals = [rng.randrange(P) for _ in range(n + 1)]
    want = circuit.weights(FR, text, vals)
    toxic = tuple(rng.randrange(1, P) for _ in range(5))
    r, s = rng.randrange(1, P), rng.randrange(1, P)
    crs = zk.setup(ctx, qap, toxic)
    d_vals, d_wit = ctx.dev_alloc(32 * len(vals)), ctx.dev_alloc(32 * qap.m)
    ctx.h2d(d_vals, zg.fr_limbs(vals))
    zk.witness_generate_dev(ctx, plan, d_vals, len(vals), d_wit)
    got = zg.prove_dev(ctx, qap, crs, d_wit, r, s)
    ref = zk.prove(ctx, qap, crs, want, r, s)
    assert (got.a, got.b, got.c) == (ref.a, ref.b, ref.c)
    assert zk.verify(ctx, crs, want[1:rep.input + 1], got)
    back = np.empty((qap.m, 4), dtype=np.uint64)
    ctx.d2h(back, d_wit)
    assert zg.limbs_to_ints(back) == want
    ctx.dev_free(d_vals)
    ctx.dev_free(d_wit)


@pytest.mark.parametrize("width,depth,fan_in", [(1, 40, 1), (8, 5, 3), (32, 20, 2), (33, 20, 2), (64, 64, 2), (513, 9, 2), (600, 3, 4),
                                                (1024, 64, 2), (4096, 5, 2), (4097, 5, 2), (8192, 4, 2)])
def test_layered_circuit_matches_oracle(ctx, width, depth, fan_in):
    """Wide synthetic circuits against the gate-by-gate walk on the CPU; the plan's levels are the circuit's layers.  Levels
    above 512 gates are one launch each (chained by programmatic dependent launch), runs of narrower levels one block."""
    n, m, n_input, rows, free = zg.layered_qap_rows(width, depth, fan_in=fan_in, seed=width + depth)
    n2 = max(2, 1 << (n - 1).bit_length())
    qap = zk.QAP(ctx, n2, m, n_input, rows)
    plan = zk.WitnessPlan(ctx, qap, free)
    info = plan.info()
    assert (info["n_gates"], info["n_levels"], info["max_width"]) == (n, depth, width)
    assert info["n_launches"] == (depth + 2 if width > 512 else 3)
    rng = random.Random(n)
    vals = [rng.randrange(P) for _ in free]
    got = zk.weights(ctx, plan, vals)
    assert got == circuit.weights_from_rows(P, n, m, circuit.csr_by_gate(n, m, rows), free, vals)


def test_unit_coefficient_entries(ctx):
    """Entries with coefficient 1 skip their product (flagged at plan time); mixed with weighted entries in one gate, on the
    chain kernel (prefetched and non-prefetched entries: fan-in 6) and on wide levels."""
    for width, depth, fan_in in ((8, 30, 6), (700, 3, 6)):
        n, m, n_input, rows, free = zg.layered_qap_rows(width, depth, fan_in=fan_in, seed=99, unit_coeffs=True)
        rng = random.Random(width)
        for t in (0, 1):  # re-weight every third entry
            rows[t][2][::3, 0] = np.asarray([rng.randrange(2, 1000) for _ in range(len(rows[t][2][::3]))], dtype=np.uint64)
        qap = zk.QAP(ctx, max(2, 1 << (n - 1).bit_length()), m, n_input, rows)
        plan = zk.WitnessPlan(ctx, qap, free)
        vals = [rng.randrange(P) for _ in free]
        assert zk.weights(ctx, plan, vals) == circuit.weights_from_rows(P, n, m, circuit.csr_by_gate(n, m, rows), free, vals)


def _hand_rows(m, n, entries):
    """entries[t] = [(wire, gate, coeff)] -> CSR triples by wire"""
    rows = []
    for t in range(3):
        per = [[] for _ in range(m)]
        for w, g, c in entries[t]:
            per[w].append((g, c))
        rows.append(zg._csr(per, m))
    return rows


def test_w_coefficient_and_mixed_widths(ctx):
    """A gate whose w row carries a coefficient c != 1 (c * out = U * V: not producible by the parser, legal in a
    DummyRep) is solved with 1/c; gates listed out of order are accepted without the program-order flag."""
    rng = random.Random(9)
    c3, c5 = rng.randrange(2, P), rng.randrange(2, P)
    # wires: 0 one, 1 x, 2 y, 3 t ; gate 1: c3 * t = (x + 2) * (x + 3*1) ; gate 0: c5 * y = t * (t + x)
    entries = [[(1, 1, 1), (0, 1, 2), (3, 0, 1)], [(1, 1, 1), (0, 1, 3), (3, 0, 1), (1, 0, 1)], [(3, 1, c3), (2, 0, c5)]]
    qap = zk.QAP(ctx, 2, 4, 2, _hand_rows(4, 2, entries))
    with pytest.raises(zk.ZkbError, match="Under constrained expression"):
        zk.WitnessPlan(ctx, qap, [1], program_order=True)
    plan = zk.WitnessPlan(ctx, qap, [1], program_order=False)
    x = rng.randrange(P)
    t = (x + 2) * (x + 3) * pow(c3, -1, P) % P
    y = t * (t + x) * pow(c5, -1, P) % P
    assert zk.weights(ctx, plan, [x]) == [1, x, y, t]
    assert plan.info()["n_levels"] == 2
    entries[2] = [(3, 1, 0), (2, 0, c5)]  # 0 * t = ...: nothing to solve for
    with pytest.raises(zk.ZkbError, match="coefficient 0"):
        zk.WitnessPlan(ctx, zk.QAP(ctx, 2, 4, 2, _hand_rows(4, 2, entries)), [1], program_order=False)


def test_error_behaviour(ctx):
    """The reference's error cases (circuit/mod.rs:553-558, 601-616, 630) surface as ZKB_ERR_ARG with its messages."""
    rep, bg, fw, names = swapped_two_gates()
    m, n = len(rep.u), 2
    per = [[[] for _ in range(m)] for _ in range(3)]
    for t in range(3):
        for k in range(n):
            for w, c in bg[t][k]:
                per[t][w].append((k, c))
    qap = zk.QAP(ctx, n, m, rep.input, [zg._csr(per[t], m) for t in range(3)])
    with pytest.raises(zk.ZkbError, match="Under constrained expression"):
        zk.WitnessPlan(ctx, qap, fw)  # gate 0 reads `late`, which gate 1 assigns
    plan = zk.WitnessPlan(ctx, qap, fw, program_order=False)
    assert zk.weights(ctx, plan, [3, 5]) == circuit.weights(FR, TWO_GATES, [3, 5])
    with pytest.raises(zk.ZkbError, match="Wrong number of values supplied"):
        zk.weights(ctx, plan, [3])
    with pytest.raises(zk.ZkbError, match="already assigned variable"):
        zk.WitnessPlan(ctx, qap, fw + [names.index("y")], program_order=False)
    with pytest.raises(zk.ZkbError, match="listed twice"):
        zk.WitnessPlan(ctx, qap, fw + fw[:1], program_order=False)
    with pytest.raises(zk.ZkbError, match="Under constrained expression"):
        zk.WitnessPlan(ctx, qap, fw[:1], program_order=False)  # `a` has no value and no gate produces it
    with pytest.raises(zk.ZkbError, match="out of range"):
        zk.WitnessPlan(ctx, qap, [0], program_order=False)
    # a wire nothing reads or assigns
    qap2 = zk.QAP(ctx, n, m + 1, rep.input, [zg._csr(per[t] + [[]], m + 1) for t in range(3)])
    with pytest.raises(zk.ZkbError, match="Every variable should have an assignment"):
        zk.WitnessPlan(ctx, qap2, fw, program_order=False)
    # two gates depending on each other
    cyc = [[(1, 0, 1), (2, 1, 1)], [(0, 0, 1), (0, 1, 1)], [(2, 0, 1), (1, 1, 1)]]
    qap3 = zk.QAP(ctx, 2, 3, 1, _hand_rows(3, 2, cyc))
    with pytest.raises(zk.ZkbError, match="depend on each other"):
        zk.WitnessPlan(ctx, qap3, [], program_order=False)
    # two wires in one gate's w row
    two = [[(1, 0, 1)], [(1, 0, 1)], [(2, 0, 1), (3, 0, 1)]]
    qap4 = zk.QAP(ctx, 2, 4, 1, _hand_rows(4, 2, two))
    with pytest.raises(zk.ZkbError, match="exactly one output"):
        zk.WitnessPlan(ctx, qap4, [1])


def test_full_size_layered_circuit(ctx):
    """2^20 gates (4096 wide, 256 deep): every gate's constraint checked on the host with numpy-free Python on a
    sample, the whole vector through the device's own h(x): a satisfying assignment makes u*v - w divisible by t,
    which prove + verify confirms end to end."""
    width, depth = 4096, 256
    n, m, n_input, rows, free = zg.layered_qap_rows(width, depth, seed=11)
    qap = zk.QAP(ctx, n, m, n_input, rows)
    plan = zk.WitnessPlan(ctx, qap, free)
    rng = random.Random(12)
    vals = [rng.randrange(P) for _ in free]
    t0 = time.perf_counter()
    raw = zk.witness_generate_raw(ctx, plan, vals)
    dt = time.perf_counter() - t0
    a = zg.limbs_to_ints(raw)
    assert a[0] == 1 and a[1:width + 1] == vals
    # by-gate view of a sample of gates from the by-wire CSR
    sample = sorted(rng.sample(range(n), 2000) + [0, width - 1, width, n - 1])
    want = set(sample)
    terms = [{k: [] for k in sample} for _ in range(2)]
    for t in range(2):
        ptr, gate, coeff = rows[t]
        wire_of = np.repeat(np.arange(m), np.diff(ptr.astype(np.int64)))
        hit = np.nonzero(np.isin(gate, np.asarray(sample, dtype=np.uint32)))[0]
        for e in hit:
            terms[t][int(gate[e])].append((int(wire_of[e]), int(coeff[e][0])))
    for k in sample:
        su = sum(c * a[w] for w, c in terms[0][k]) % P
        sv = sum(c * a[w] for w, c in terms[1][k]) % P
        assert a[1 + width + k] == su * sv % P, f"gate {k}"
    toxic = tuple(rng.randrange(1, P) for _ in range(5))
    crs = zk.setup(ctx, qap, toxic)
    proof = zk.prove(ctx, qap, crs, raw, rng.randrange(1, P), rng.randrange(1, P))
    assert zk.verify(ctx, crs, a[1:n_input + 1], proof)
    bad = raw.copy()
    bad[1 + width + n // 2, 0] ^= 1
    assert not zk.verify(ctx, crs, a[1:n_input + 1], zk.prove(ctx, qap, crs, bad, 5, 7))
    print(f"witness generation 2^20 gates ({depth} levels x {width}): {dt * 1e3:.2f} ms incl. H2D/D2H, {plan.info()}")
