"""Oracle A -- synthetic workload family (test infrastructure only).

The Horner circuit ``deg_{n-1}`` generalises /root/reference/test_programs/deg_15.zk
(SURVEY.md section 8d):  gate 1: t1 = x*c1 ; gate k: t_k = x*(t_{k-1}+c_k) ;
gate n: y = 1*(t_{n-1}+c_n).  Row order is the one ``ASTParser::try_parse``
(src/groth16/circuit/mod.rs:230-526) produces for that program text:
0 unity, 1 x, 2 y, then t_k -> 2k+1, c_k -> 2k+2 (k<n), c_n -> 2n+1.
"""

from __future__ import annotations

from .circuit import DummyRep
from .fields import FR, Field

TWO_ADICITY = 28
OMEGA_2_28 = pow(5, (FR.p - 1) >> TWO_ADICITY, FR.p)


def omega(log_n: int) -> int:
    """Primitive 2^log_n-th root of unity of Fr: 5^((r-1)/2^28) squared down."""
    assert 0 <= log_n <= TWO_ADICITY
    return pow(OMEGA_2_28, 1 << (TWO_ADICITY - log_n), FR.p)


def horner_program_text(n: int) -> str:
    """The .zk text of the n-gate Horner circuit (n=16 reproduces deg_15.zk's structure)."""
    ins = " ".join(["x"] + [f"c{k}" for k in range(1, n + 1)])
    lines = [f"(in {ins})", "(out y)", "(verify x y)", "", "(program"]
    if n == 1:
        lines.append("    (= y (* x c1))")
    else:
        lines.append("    (= t1 (* x c1))")
        for k in range(2, n):
            lines.append(f"    (= t{k} (* x (+ t{k-1} c{k})))")
        lines.append(f"    (= y (* 1 (+ t{n-1} c{n})))")
    lines[-1] += ")"
    return "\n".join(lines)


def horner_rep(F: Field, n: int, roots: list) -> DummyRep:
    """Sparse rows of the n-gate Horner circuit (n >= 2) on an arbitrary root list."""
    assert n >= 2 and len(roots) == n
    M = 2 * n + 2
    one = F.from_usize(1)
    u = [[] for _ in range(M)]
    v = [[] for _ in range(M)]
    w = [[] for _ in range(M)]
    t = lambda k: 2 * k + 1
    c = lambda k: 2 * k + 2 if k < n else 2 * n + 1
    for k in range(1, n + 1):
        g = roots[k - 1]
        if k < n:
            u[1].append((g, one))
            w[t(k)].append((g, one))
        else:
            u[0].append((g, one))
            w[2].append((g, one))
        if k > 1:
            v[t(k - 1)].append((g, one))
        v[c(k)].append((g, one))
    return DummyRep(u=u, v=v, w=w, roots=list(roots), input=2)


def horner_witness(F: Field, n: int, x: int, cs: list) -> list:
    """Weights in row order for inputs x, c_1..c_n."""
    M = 2 * n + 2
    a = [0] * M
    a[0], a[1] = F.one(), x
    acc = F.mul(x, cs[0])
    a[3], a[4] = acc, cs[0]
    for k in range(2, n):
        acc = F.mul(x, F.add(acc, cs[k - 1]))
        a[2 * k + 1], a[2 * k + 2] = acc, cs[k - 1]
    a[2 * n + 1] = cs[n - 1]
    a[2] = F.add(acc, cs[n - 1])
    return a
