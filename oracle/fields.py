"""Oracle A -- field layer (test infrastructure only, see oracle/__init__.py).

Elements are plain Python ints; a ``Field`` object carries the arithmetic.  Two
instances mirror the two ``Field`` impls of the reference:

* ``Z251``  -- src/field/z251.rs:1-97 (toy field in a u8, the reference's test field)
* ``FR``    -- src/groth16/fr.rs:9-99 (``FrLocal`` newtype over ``bn::Fr``; BN254 scalar field)
"""

from __future__ import annotations


class FieldPanic(Exception):
    """Stands in for a Rust panic (``expect``/``assert!``) on the reference side."""


class Field:
    """Generic prime field with the reference's trait surface.

    Mirrors ``FieldIdentity`` (src/field/mod.rs:62-65: zero/one) and ``Field``
    (src/field/mod.rs:77-93: + - * / neg mul_inv).
    """

    name = "F"

    def __init__(self, p: int):
        self.p = p

    # FieldIdentity
    def zero(self) -> int:
        return 0

    def one(self) -> int:
        return 1

    def add(self, a: int, b: int) -> int:
        return (a + b) % self.p

    def neg(self, a: int) -> int:
        return (-a) % self.p

    def sub(self, a: int, b: int) -> int:
        return (a - b) % self.p

    def mul(self, a: int, b: int) -> int:
        return (a * b) % self.p

    def mul_inv(self, a: int) -> int:
        if a % self.p == 0:
            raise FieldPanic("Tried to get mul inv of zero")
        return pow(a, -1, self.p)

    def div(self, a: int, b: int) -> int:
        if b % self.p == 0:
            raise FieldPanic("Tried to divide by zero")
        return (a * pow(b, -1, self.p)) % self.p

    def from_usize(self, n: int) -> int:
        return n % self.p

    def from_str(self, s: str) -> int:
        return int(s) % self.p

    def eq(self, a: int, b: int) -> bool:
        return a == b

    def is_zero(self, a: int) -> bool:
        return a == self.zero()


class _Z251(Field):
    """src/field/z251.rs.  Quirks kept on purpose:

    * ``neg`` is ``251 - inner`` (z251.rs:21-29): ``-0`` is the NON-canonical 251,
      and derived ``PartialEq`` compares the raw byte, so ``-0 != 0``.
    * ``div`` goes through ``ext_euc_alg`` (z251.rs:50-61, field/mod.rs:360-385);
      dividing by zero yields 0 instead of panicking.
    * ``From<usize>`` asserts ``n < 251`` (z251.rs:78-83).
    """

    name = "Z251"

    def __init__(self):
        super().__init__(251)

    def add(self, a, b):
        return (a + b) % 251  # u16 sum % 251, z251.rs:11-17

    def neg(self, a):
        return 251 - a  # z251.rs:24-28 (not reduced)

    def sub(self, a, b):
        return self.add(a, self.neg(b))  # z251.rs:34-36

    def mul(self, a, b):
        return (a * b) % 251

    def div(self, a, b):
        # ext_euc_alg(rhs, 251) on isize with truncating division
        r0, r1, s0, s1 = b, 251, 1, 0
        while r1 != 0:
            q = int(r0 / r1)  # isize division truncates toward zero
            r0, r1 = r1, r0 - q * r1
            s0, s1 = s1, s0 - q * s1
        inv = s0
        while inv < 0:
            inv += 251
        return self.mul(a, inv % 256)

    def mul_inv(self, a):
        return self.div(1, a)  # z251.rs:72-76

    def from_usize(self, n):
        if not n < 251:
            raise FieldPanic("assertion failed: n < 251")
        return n

    def from_str(self, s):
        return self.from_usize(int(s))


class _Fr(Field):
    """src/groth16/fr.rs:18-99 over bn::Fr (BN254 scalar field, canonical residues).

    ``Div``/``mul_inv`` panic on zero (fr.rs:54,69); ``From<usize>`` parses the
    decimal string (fr.rs:73-77).
    """

    name = "Fr"

    def __init__(self):
        super().__init__(
            21888242871839275222246405745257275088548364400416034343698204186575808495617
        )


Z251 = _Z251()
FR = _Fr()

# BN254 base field modulus (crate bn, Fq)
Q_MODULUS = 21888242871839275222246405745257275088696311157297823662689037894645226208583
FQ = Field(Q_MODULUS)
FQ.name = "Fq"
