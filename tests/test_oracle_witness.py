"""CPU tests of the witness-generation oracle (oracle/circuit.py: weights_from_rows -- the gate-by-gate walk of
`weights()` / `evaluate()`, /root/reference/src/groth16/circuit/mod.rs:598-656, restated on the DummyRep rows).  It is
pinned to the literal restatement `circuit.weights` (itself pinned to the reference's `weights_test` golden vector,
circuit/mod.rs:746-769) and then serves as the checker for the device path on circuits that have no program text
(tests/test_gpu_witness.py)."""

import importlib
import random

import pytest

from oracle import circuit, synthetic
from oracle.fields import FR, Z251
from test_oracle_kats import QUAD, SIMPLE

zg = importlib.import_module("zksnark-rs_b200.groth16")

MIXED = """(in x a b)
(out y z)
(verify x y z)

(program
    (= t1 (* x x))
    (= t2 (* (+ t1 a) (+ x b 7)))
    (= y (* 1 (+ t2 t1 3)))
    (= z (* t2 (+ y x))))"""

# two gates; with their order swapped, gate 0 reads `late` before gate 1 assigns it: the sequential walk of weights()
# rejects that ("Under constrained expression"), a topological evaluation (the builder's evaluate) does not
TWO_GATES = """(in x a)
(out y)
(verify x y)

(program
    (= late (* x a))
    (= y (* late (+ x 1))))"""


def swapped_two_gates():
    rep = circuit.try_parse(FR, TWO_GATES)
    bg = tuple(list(reversed(g)) for g in circuit.rep_by_gate(rep))
    names = ["1"] + circuit.variable_order(circuit.try_to_list(FR, TWO_GATES))
    return rep, bg, circuit.input_wires(FR, TWO_GATES), names


def test_reference_weights_vector():  # circuit/mod.rs:746-769
    rep = circuit.try_parse(Z251, SIMPLE)
    got = circuit.weights_from_rows(251, len(rep.roots), len(rep.u), circuit.rep_by_gate(rep), circuit.input_wires(Z251, SIMPLE), [3, 2, 4])
    assert got == [1, 2, 34, 6, 3, 4] == circuit.weights(Z251, SIMPLE, [3, 2, 4])


@pytest.mark.parametrize("F", [Z251, FR])
@pytest.mark.parametrize("text,n_in", [(SIMPLE, 3), (QUAD, 4), (MIXED, 3), (synthetic.horner_program_text(16), 17),
                                       (synthetic.horner_program_text(100), 101)])
def test_rows_walk_equals_literal_weights(F, text, n_in):
    rng = random.Random(n_in)
    rep = circuit.try_parse(F, text)
    vals = [rng.randrange(F.p) for _ in range(n_in)]
    want = circuit.weights(F, text, vals)
    got = circuit.weights_from_rows(F.p, len(rep.roots), len(rep.u), circuit.rep_by_gate(rep), circuit.input_wires(F, text), vals)
    assert got == want


def test_rows_walk_errors():
    rep, bg, fw, names = swapped_two_gates()
    n, m = len(rep.roots), len(rep.u)
    with pytest.raises(circuit.ParseErr, match="Under constrained"):
        circuit.weights_from_rows(FR.p, n, m, bg, fw, [3, 5])
    a = circuit.weights_from_rows(FR.p, n, m, bg, fw, [3, 5], program_order=False)
    assert a == circuit.weights(FR, TWO_GATES, [3, 5])
    env = dict(zip(names, a))
    assert env["late"] == 15 and env["y"] == 15 * 4
    with pytest.raises(circuit.ParseErr, match="Wrong number"):
        circuit.weights_from_rows(FR.p, n, m, bg, fw, [3])
    with pytest.raises(circuit.ParseErr, match="already assigned"):
        circuit.weights_from_rows(FR.p, n, m, bg, fw + [names.index("y")], [3, 5, 1])
    with pytest.raises(circuit.ParseErr, match="Under constrained"):
        circuit.weights_from_rows(FR.p, n, m, bg, fw[:1], [3], program_order=False)


def test_layered_rows_shape_and_walk():
    n, m, n_input, rows, free = zg.layered_qap_rows(8, 5, fan_in=3, seed=7)
    assert (n, m, free) == (40, 49, list(range(1, 9)))
    bg = circuit.csr_by_gate(n, m, rows)
    for k in range(n):
        lo = 1 if k < 8 else 1 + 8 + (k // 8 - 1) * 8
        for t in (0, 1):
            assert 1 <= len(bg[t][k]) <= 3 and all(lo <= w < lo + 8 for w, _ in bg[t][k])
            assert 3 <= sum(c for _, c in bg[t][k]) <= 21  # three literals in 1..7, duplicates merged
        assert bg[2][k] == [(9 + k, 1)]
    rng = random.Random(3)
    a = circuit.weights_from_rows(FR.p, n, m, bg, free, [rng.randrange(FR.p) for _ in free])
    for k in range(n):  # every gate's constraint holds
        su = sum(c * a[w] for w, c in bg[0][k]) % FR.p
        sv = sum(c * a[w] for w, c in bg[1][k]) % FR.p
        assert a[9 + k] == su * sv % FR.p


# ---- the planner's host logic (libzkb200's zkb_witness_levels: no device needed) ------------------------------------------
def _levels_ref(n, m, by_gate, free):
    """level(gate) = 1 + the deepest producer among its u / v inputs; free wires and the unity wire are level 0."""
    gu, gv, gw = by_gate
    wl = {0: 0, **{w: 0 for w in free}}
    prod = {gw[k][0][0]: k for k in range(n) if len(gw[k]) == 1}
    lev = [0] * n

    def level(k):
        if lev[k]:
            return lev[k]
        lv = 0
        for w, _ in gu[k] + gv[k]:
            lv = max(lv, wl[w] if w in wl else level(prod[w]))
        lev[k] = lv + 1
        wl[gw[k][0][0]] = lv + 1
        return lev[k]

    for k in range(n):
        if len(gw[k]) == 1:
            level(k)
    return lev


def _rows_of(by_gate, m):
    per = [[[] for _ in range(m)] for _ in range(3)]
    for t in range(3):
        for k, row in enumerate(by_gate[t]):
            for w, c in row:
                per[t][w].append((k, c))
    return [zg._csr(per[t], m) for t in range(3)]


@pytest.mark.parametrize("text,n_in", [(SIMPLE, 3), (QUAD, 4), (MIXED, 3), (synthetic.horner_program_text(16), 17)])
def test_planner_levels_on_parser_circuits(text, n_in):
    rep = circuit.try_parse(FR, text)
    bg, fw = circuit.rep_by_gate(rep), circuit.input_wires(FR, text)
    n, m = len(rep.roots), len(rep.u)
    got = zg.witness_levels(n, m, rep.input, _rows_of(bg, m), fw)
    assert got == _levels_ref(n, m, bg, fw) and min(got) >= 1
    assert got == zg.witness_levels(n, m, rep.input, _rows_of(bg, m), fw, program_order=False)


def test_planner_levels_layered_and_padding():
    n, m, n_input, rows, free = zg.layered_qap_rows(16, 7, fan_in=3, seed=2)
    got = zg.witness_levels(32 * 4, m, n_input, rows, free)  # 112 gates on a 128-gate domain: the padding gates assign nothing
    assert got[:n] == [1 + k // 16 for k in range(n)] and got[n:] == [0] * (128 - n)


def test_planner_error_messages():
    rep, bg, fw, names = swapped_two_gates()
    m = len(rep.u)
    rows = _rows_of(bg, m)
    with pytest.raises(zg.ZkbError, match="Under constrained expression.*later gate"):
        zg.witness_levels(2, m, rep.input, rows, fw)
    assert zg.witness_levels(2, m, rep.input, rows, fw, program_order=False) == [2, 1]
    with pytest.raises(zg.ZkbError, match="already assigned variable"):
        zg.witness_levels(2, m, rep.input, rows, fw + [names.index("y")], program_order=False)
    with pytest.raises(zg.ZkbError, match="listed twice"):
        zg.witness_levels(2, m, rep.input, rows, fw + fw[:1])
    with pytest.raises(zg.ZkbError, match="Under constrained expression"):
        zg.witness_levels(2, m, rep.input, rows, fw[:1], program_order=False)
    with pytest.raises(zg.ZkbError, match="out of range"):
        zg.witness_levels(2, m, rep.input, rows, [m])
    rows2 = _rows_of(tuple(g + [] for g in bg), m + 1)  # one more wire that nothing reads or assigns
    with pytest.raises(zg.ZkbError, match="Every variable should have an assignment"):
        zg.witness_levels(2, m + 1, rep.input, rows2, fw, program_order=False)
    cyc = ([[(1, 1)], [(2, 1)]], [[(0, 1)], [(0, 1)]], [[(2, 1)], [(1, 1)]])  # b = a * 1, a = b * 1
    with pytest.raises(zg.ZkbError, match="depend on each other"):
        zg.witness_levels(2, 3, 1, _rows_of(cyc, 3), [], program_order=False)
    two = ([[(1, 1)], []], [[(1, 1)], []], [[(2, 1), (3, 1)], []])
    with pytest.raises(zg.ZkbError, match="exactly one output"):
        zg.witness_levels(2, 4, 1, _rows_of(two, 4), [1])
