"""Oracle A -- .zk front end (test infrastructure only, see oracle/__init__.py).

Restates just enough of the reference's circuit front end to reproduce the
row / witness ORDER the hot path consumes:

* tokeniser + expression tree  src/groth16/circuit/ast.rs:62-370
* ``ASTParser::try_parse``     src/groth16/circuit/mod.rs:230-526  -> ``DummyRep``
* ``weights()``                src/groth16/circuit/mod.rs:529-637
* legacy line format           src/groth16/circuit/dummy_rep.rs:55-142
"""

from __future__ import annotations

from dataclasses import dataclass, field as _dc_field

from .fields import Field


class ParseErr(Exception):
    pass


@dataclass
class DummyRep:
    """src/groth16/circuit/dummy_rep.rs:7-13.  Rows are lists of (root, value)."""

    u: list
    v: list
    w: list
    roots: list
    input: int


_KEYWORDS = {"in": "In", "out": "Out", "verify": "Verify", "program": "Program",
             "=": "Equal", "*": "Mul", "+": "Add"}


def try_to_list(F: Field, code: str) -> list:
    """ast.rs:263-287 + parse_token :300-370.  Tokens are tuples (kind, payload)."""
    tokens = []
    for lineno, line in enumerate(code.splitlines(), 1):
        for sub in line.split():
            if sub.startswith("("):
                tokens.append(("open", None))
                sub = sub[1:]
                opened = True
            else:
                opened = False
            if len(sub) == 0:
                raise ParseErr(f"SyntaxErr({lineno}, found whitespace after '(')")
            if sub in _KEYWORDS:
                tokens.append(("kw", _KEYWORDS[sub]))
                continue
            if "(" in sub:
                raise ParseErr(f"SyntaxErr({lineno}, unexpected '(')")
            if any(ch in sub for ch in "*+="):
                raise ParseErr(f"SyntaxErr({lineno}, unexpected operator)")
            k = sub.find(")")
            start, end = (sub, "") if k < 0 else (sub[:k], sub[k:])
            if opened and end:
                raise ParseErr(f"SyntaxErr({lineno}, unexpected ')')")
            if start[:1].isnumeric():
                try:
                    tokens.append(("lit", F.from_str(start)))
                except Exception as e:  # noqa: BLE001
                    raise ParseErr(f"SyntaxErr({lineno}, could not parse literal)") from e
            else:
                tokens.append(("var", start))
            for ch in end:
                if ch != ")":
                    raise ParseErr(f"SyntaxErr({lineno}, expected ')')")
                tokens.append(("close", None))
    return tokens


def variable_order(tokens: list) -> list:
    """ast.rs:62-83: first appearance order of Vars from the 'verify' keyword onward."""
    seen, out, started = set(), [], False
    for kind, val in tokens:
        if not started:
            if kind == "kw" and val == "Verify":
                started = True
            else:
                continue
        if kind == "var" and val not in seen:
            seen.add(val)
            out.append(val)
    return out


def _next_group(it) -> list:
    """ast.rs:230-261."""
    tok = next(it, None)
    if tok is None:
        return []
    if tok[0] == "open":
        depth, out = 1, []
        for t in it:
            if t[0] == "open":
                depth += 1
            elif t[0] == "close":
                depth -= 1
            if depth == 0:
                break
            out.append(t)
        return out
    if tok[0] in ("var", "lit"):
        return [tok]
    raise ParseErr("Cannot parse malformed group")


def _parse_expression(tokens: list):
    """ast.rs:106-228.  Expression = (tag, ...) tuples."""
    it = iter(tokens)
    tok = next(it, None)
    if tok is None:
        raise ParseErr("Malformed expression")
    kind, val = tok
    if kind == "var":
        return ("Var", val)
    if kind == "lit":
        return ("Literal", val)
    if kind != "kw":
        raise ParseErr("Malformed expression")
    if val in ("In", "Out", "Verify"):
        vs = []
        for t in it:
            if t[0] != "var":
                raise ParseErr(f"Non variable found in '{val.lower()}' expression")
            vs.append(("Var", t[1]))
        return (val, vs)
    if val in ("Program", "Add"):
        items = []
        while True:
            g = _next_group(it)
            if not g:
                break
            items.append(_parse_expression(g))
        return (val, items)
    if val == "Equal":
        left = _next_group(it)
        if len(left) != 1 or left[0][0] != "var":
            raise ParseErr("Can only assign to a variable")
        right = _parse_expression(_next_group(it))
        return ("Assign", ("Var", left[0][1]), right)
    if val == "Mul":
        left = _parse_expression(_next_group(it))
        right = _parse_expression(_next_group(it))
        return ("Mul", left, right)
    raise ParseErr("Malformed expression")


def expressions(F: Field, code: str) -> list:
    """ast.rs:85-104."""
    it = iter(try_to_list(F, code))
    out = []
    while True:
        g = _next_group(it)
        if not g:
            break
        out.append(_parse_expression(g))
    return out


def try_parse(F: Field, code: str) -> DummyRep:
    """``ASTParser::try_parse``, circuit/mod.rs:230-526."""
    exps = expressions(F, code)
    if len(exps) != 4:
        raise ParseErr("Expected exactly one each of 'in', 'out', 'verify' and 'program'")
    if exps[0][0] != "In":
        raise ParseErr("Expected first expression to be 'in'")
    if exps[1][0] != "Out":
        raise ParseErr("Expected second expression to be 'out'")
    if exps[2][0] != "Verify":
        raise ParseErr("Expected third expression to be 'verify'")
    if exps[3][0] != "Program":
        raise ParseErr("Expected fourth expression to be 'program'")

    variables: dict = {}
    u, v, w = [[]], [[]], [[]]
    n_input = 0
    one = F.from_usize(1)

    for _, name in exps[2][1]:
        variables[name] = len(u)  # duplicates overwrite, exactly like HashMap::insert
        u.append([]); v.append([]); w.append([])
        n_input += 1

    gate = 0

    def side(rows_self, which, name, coeff, g):
        """Add (gate, coeff) to variable ``name`` on side ``which`` ('u' or 'v')."""
        if name not in variables:
            variables[name] = len(rows_self)
            u.append([(g, coeff)] if which == "u" else [])
            v.append([(g, coeff)] if which == "v" else [])
            w.append([])
        else:
            rows_self[variables[name]].append((g, coeff))

    def handle(exp, rows, which, g):
        tag = exp[0]
        if tag == "Literal":
            rows[0].append((g, exp[1]))
        elif tag == "Var":
            side(rows, which, exp[1], one, g)
        elif tag == "Add":
            for e in exp[1]:
                if e[0] == "Literal":
                    rows[0].append((g, e[1]))
                elif e[0] == "Var":
                    side(rows, which, e[1], one, g)
                elif e[0] == "Mul":
                    if e[1][0] != "Literal":
                        raise ParseErr("LHS of a '*' expression in a '+' expression must be a literal")
                    if e[2][0] != "Var":
                        raise ParseErr("RHS of a '*' expression in a '+' expression must be a variable")
                    side(rows, which, e[2][1], e[1][1], g)
                else:
                    raise ParseErr("Invalid expression found in '+' expression")
        else:
            raise ParseErr("Invalid expression found in '*' expression")

    for assignment in exps[3][1]:
        gate += 1
        g = F.from_usize(gate)
        if assignment[0] != "Assign":
            raise ParseErr("Program expression must be a list of '=' expressions")
        name = assignment[1][1]
        if name not in variables:
            variables[name] = len(u)
            u.append([]); v.append([]); w.append([(g, one)])
        else:
            idx = variables[name]
            if idx <= n_input:
                if len(w[idx]) != 0:
                    raise ParseErr("Varify variable cannot be the output of two different gates")
                w[idx].append((g, one))
            else:
                raise ParseErr("Already declared variable cannot be the output wire of a gate")
        right = assignment[2]
        if right[0] == "Mul":  # anything else is silently ignored (mod.rs:339 `if let`)
            handle(right[1], u, "u", g)
            handle(right[2], v, "v", g)

    roots = [F.from_usize(k) for k in range(1, gate + 1)]
    return DummyRep(u=u, v=v, w=w, roots=roots, input=n_input)


def _evaluate(F: Field, exp, env: dict):
    """circuit/mod.rs:639-656."""
    tag = exp[0]
    if tag == "Literal":
        return exp[1]
    if tag == "Var":
        return env.get(exp[1])
    if tag == "Mul":
        l = _evaluate(F, exp[1], env)
        if l is None:
            return None
        r = _evaluate(F, exp[2], env)
        return None if r is None else F.mul(l, r)
    if tag == "Add":
        acc = F.zero()
        for e in exp[1]:
            val = _evaluate(F, e, env)
            if val is None:
                return None
            acc = F.add(acc, val)
        return acc
    return None


def weights(F: Field, code: str, values: list) -> list:
    """``groth16::weights``, circuit/mod.rs:529-637: [1, verify vars..., others by first appearance]."""
    exps = expressions(F, code)
    order = variable_order(try_to_list(F, code))
    if not exps or exps[0][0] != "In":
        raise ParseErr("Expected first expression to be 'in'")
    inputs = exps[0][1]
    if len(inputs) != len(values):
        raise ParseErr("Wrong number of values supplied")
    env = {name: val for (_, name), val in zip(inputs, values)}
    if len(exps) < 2 or exps[1][0] != "Out":
        raise ParseErr("Expected second expression to be 'out'")
    if len(exps) < 3 or exps[2][0] != "Verify":
        raise ParseErr("Expected third expression to be 'verify'")
    if len(exps) < 4 or exps[3][0] != "Program":
        raise ParseErr("Expected fourth expression to be 'program'")
    for assignment in exps[3][1]:
        if assignment[0] != "Assign":
            raise ParseErr("Program expression must be a list of '=' expressions")
        name = assignment[1][1]
        if name in env:
            raise ParseErr("Attempted to assign to an already assigned variable")
        val = _evaluate(F, assignment[2], env)
        if val is None:
            raise ParseErr("Under constrained expression")
        env[name] = val
    out = [F.one()]
    for name in order:
        if name not in env:
            raise ParseErr("Every variable should have an assignment")
        out.append(env.pop(name))
    return out


def dummy_rep_from_legacy(F: Field, code: str) -> DummyRep:
    """``From<&str> for DummyRep<Z251>``, dummy_rep.rs:55-142 (quad_share.zk / cubic_share.zk)."""
    lines = code.split("\n")
    lines = [ln.rstrip("\r") for ln in lines]
    if lines and lines[-1] == "":
        lines = lines[:-1]  # str::lines() drops a trailing empty line
    inputs = lines[0].split(" ")
    witness = lines[1].split(" ")
    temps = lines[2].split(" ")
    names = inputs + witness + temps
    num_vars = len(names) + 1
    u = [[] for _ in range(num_vars)]
    v = [[] for _ in range(num_vars)]
    w = [[] for _ in range(num_vars)]
    one = F.from_usize(1)
    count = 0
    for n, line in enumerate(lines[4:]):
        count += 1
        g = F.from_usize(n + 1)
        syms = iter(line.split(" "))
        first = next(syms)
        w[names.index(first) + 1].append((g, one))
        next(syms)  # "("
        for tok in syms:
            if tok == ")":
                break
            if tok == "1":
                u[0].append((g, one))
            else:
                u[names.index(tok) + 1].append((g, one))
        next(syms, None)  # "("
        for tok in syms:
            if tok == ")":
                break
            v[names.index(tok) + 1].append((g, one))
    roots = [F.from_usize(k) for k in range(1, count + 1)]
    return DummyRep(u=u, v=v, w=w, roots=roots, input=len(inputs))


def input_wires(F: Field, code: str) -> list:
    """Indices (into the weight vector) of the program's `(in ...)` variables, in their order of declaration:
    weights()[i] belongs to variable_order()[i - 1] (circuit/mod.rs:625-636), so `in` variable v sits at
    1 + variable_order().index(v)."""
    exps = expressions(F, code)
    order = variable_order(try_to_list(F, code))
    return [1 + order.index(name) for (_, name) in exps[0][1]]


def weights_from_rows(p: int, n: int, m: int, by_gate, free_wires: list, values: list, program_order: bool = True) -> list:
    """The same walk as ``weights()`` / ``evaluate()`` (circuit/mod.rs:598-656), restated on the DummyRep rows instead
    of the expression tree: gate k (= the k-th assignment `(= var (* lhs rhs))`) computes
        a[out_k] = (sum of coeff * a[wire] over gate k's u entries) * (same over its v entries) / (w coefficient),
    out_k the single wire of gate k's w row.  ``by_gate``: for u, v, w a list (per gate) of [(wire, coeff)].  Errors as
    the reference's: assigning twice, reading a wire without a value ("Under constrained expression"), a wire nothing
    assigns, wrong number of values.  ``program_order=False``: gates may appear in any order (the builder's demand-driven
    ``evaluate``, builder/mod.rs:556-580); evaluated by repeated sweeps."""
    if len(free_wires) != len(values):
        raise ParseErr("Wrong number of values supplied")
    a = [None] * m
    a[0] = 1
    for w, val in zip(free_wires, values):
        if a[w] is not None:
            raise ParseErr("Attempted to assign to an already assigned variable")
        a[w] = val % p
    gu, gv, gw = by_gate
    outs = set()
    for k in range(n):
        if len(gw[k]) == 1:
            w = gw[k][0][0]
            if a[w] is not None or w in outs:
                raise ParseErr("Attempted to assign to an already assigned variable")
            outs.add(w)
        elif len(gw[k]) > 1:
            raise ParseErr("more than one output wire")
    pending = [k for k in range(n) if len(gw[k]) == 1]
    while pending:
        later = []
        for k in pending:
            terms = [a[w] for w, _ in gu[k]] + [a[w] for w, _ in gv[k]]
            if any(t is None for t in terms):
                if program_order:
                    raise ParseErr("Under constrained expression")
                later.append(k)
                continue
            su = sum(c * a[w] for w, c in gu[k]) % p
            sv = sum(c * a[w] for w, c in gv[k]) % p
            w, c = gw[k][0]
            a[w] = su * sv * pow(c, -1, p) % p
        if len(later) == len(pending):
            raise ParseErr("Under constrained expression")
        pending = later
    if any(x is None for x in a):
        raise ParseErr("Every variable should have an assignment")
    return a


def rep_by_gate(rep: DummyRep) -> tuple:
    """DummyRep rows (per wire: [(root, value)]) -> per gate [(wire, value)], gate k <-> rep.roots[k]."""
    index = {r: k for k, r in enumerate(rep.roots)}
    out = []
    for mat in (rep.u, rep.v, rep.w):
        g = [[] for _ in rep.roots]
        for wire, row in enumerate(mat):
            for r, c in row:
                g[index[r]].append((wire, c))
        out.append(g)
    return tuple(out)


def csr_by_gate(n: int, m: int, rows) -> tuple:
    """CSR triples by wire (row_ptr, gate, coeff limbs) x (u, v, w) -> per gate [(wire, coeff int)]."""
    out = []
    for ptr, gate, coeff in rows:
        g = [[] for _ in range(n)]
        for wire in range(m):
            for e in range(int(ptr[wire]), int(ptr[wire + 1])):
                c = sum(int(coeff[e][j]) << (64 * j) for j in range(4))
                g[int(gate[e])].append((wire, c))
        out.append(g)
    return tuple(out)
