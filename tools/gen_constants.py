#!/usr/bin/env python3
"""Generate zksnark-rs_b200/csrc/constants.h (BN254 field/curve constants, Montgomery form, R = 2^256).

Every value is derived here from the two primes and the published generators, and self-checked
(primality via pow, on-curve, order) before it is written, so host code, device code and tests share
one source of truth.
"""
import os, sys

R_MOD = 21888242871839275222246405745257275088548364400416034343698204186575808495617
Q_MOD = 21888242871839275222246405745257275088696311157297823662689037894645226208583
MONT = 1 << 256

def limbs32(x): return [(x >> (32 * i)) & 0xffffffff for i in range(8)]
def arr(x): return "{" + ", ".join("0x%08xu" % l for l in limbs32(x)) + "}"
def mont(x, p): return x * MONT % p

def params(name, p, extra=""):
    inv32 = (-pow(p, -1, 1 << 32)) % (1 << 32)
    L = limbs32(p)
    s = [f"struct {name} {{"]
    for i, l in enumerate(L):
        s.append(f"  static constexpr uint32_t P{i} = 0x{l:08x}u;")
    s.append(f"  static constexpr uint32_t INV = 0x{inv32:08x}u;  // -p^-1 mod 2^32")
    for nm, val in (("ONE", MONT % p), ("R2", MONT * MONT % p), ("R3", MONT ** 3 % p)):
        for i, l in enumerate(limbs32(val)):
            s.append(f"  static constexpr uint32_t {nm}{i} = 0x{l:08x}u;")
    # p^2 as 16 limbs: added to a difference of two wide products to keep it non-negative before REDC
    for i in range(16):
        s.append(f"  static constexpr uint32_t PSQ{i} = 0x{((p * p) >> (32 * i)) & 0xffffffff:08x}u;")
    s.append(extra)
    s.append("};")
    return "\n".join(s)

def g1_add(P, Q_):
    p = Q_MOD
    if P is None: return Q_
    if Q_ is None: return P
    (x1, y1), (x2, y2) = P, Q_
    if x1 == x2:
        if (y1 + y2) % p == 0: return None
        lam = 3 * x1 * x1 * pow(2 * y1, -1, p) % p
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, p) % p
    x3 = (lam * lam - x1 - x2) % p
    return (x3, (lam * (x1 - x3) - y1) % p)

def f2mul(a, b): return ((a[0]*b[0] - a[1]*b[1]) % Q_MOD, (a[0]*b[1] + a[1]*b[0]) % Q_MOD)
def f2inv(a):
    n = pow(a[0]*a[0] + a[1]*a[1], -1, Q_MOD); return (a[0]*n % Q_MOD, -a[1]*n % Q_MOD)
def f2sub(a, b): return ((a[0]-b[0]) % Q_MOD, (a[1]-b[1]) % Q_MOD)
def g2_add(P, Q_):
    if P is None: return Q_
    if Q_ is None: return P
    (x1, y1), (x2, y2) = P, Q_
    if x1 == x2:
        if ((y1[0]+y2[0]) % Q_MOD, (y1[1]+y2[1]) % Q_MOD) == (0, 0): return None
        xx = f2mul(x1, x1)
        lam = f2mul((3*xx[0] % Q_MOD, 3*xx[1] % Q_MOD), f2inv((2*y1[0] % Q_MOD, 2*y1[1] % Q_MOD)))
    else:
        lam = f2mul(f2sub(y2, y1), f2inv(f2sub(x2, x1)))
    x3 = f2sub(f2sub(f2mul(lam, lam), x1), x2)
    return (x3, f2sub(f2mul(lam, f2sub(x1, x3)), y1))
def smul(add, P, k):
    acc = None
    while k:
        if k & 1: acc = add(acc, P)
        P = add(P, P); k >>= 1
    return acc

def main(out):
    for p in (R_MOD, Q_MOD):
        assert pow(2, p - 1, p) == 1 and pow(3, p - 1, p) == 1 and p.bit_length() == 254
    assert (R_MOD - 1) % (1 << 28) == 0 and pow(5, (R_MOD - 1) // 2, R_MOD) == R_MOD - 1
    omega = pow(5, (R_MOD - 1) >> 28, R_MOD)
    assert pow(omega, 1 << 28, R_MOD) == 1 and pow(omega, 1 << 27, R_MOD) == R_MOD - 1
    G1 = (1, 2)
    G2 = ((10857046999023057135944570762232829481370756359578518086990519993285655852781,
           11559732032986387107991004021392285783925812861821192530917403151452391805634),
          (8495653923123431417604973247489272438418190587263600148770280649306958101930,
           4082367875863433681332203403145435568316851327593401208105741076214120093531))
    assert smul(g1_add, G1, R_MOD) is None and smul(g2_add, G2, R_MOD) is None
    B1 = smul(g1_add, G1, 69)   # reference base point, src/groth16/fr.rs:106-109
    B2 = smul(g2_add, G2, 96)   # src/groth16/fr.rs:110-113
    b2 = f2mul((3, 0), f2inv((9, 1)))
    # pairing constants: Frobenius twists xi^((q-1)/3), xi^((q-1)/2), xi^((q^2-1)/3) (xi = 9 + u), the final
    # exponent (q^12 - 1)/r and the optimal-ate loop count 6u + 2 (u = 4965661367192848881)
    def f2pow(a, e):
        r = (1, 0)
        while e:
            if e & 1: r = f2mul(r, a)
            a = f2mul(a, a); e >>= 1
        return r
    fg2, fg3 = f2pow((9, 1), (Q_MOD - 1) // 3), f2pow((9, 1), (Q_MOD - 1) // 2)
    fg2s = f2pow((9, 1), (Q_MOD * Q_MOD - 1) // 3)
    assert fg2s[1] == 0 and f2pow((9, 1), (Q_MOD * Q_MOD - 1) // 2) == (Q_MOD - 1, 0)
    # final exponentiation = easy part (q^6 - 1)(q^2 + 1) by conjugation / inversion / Frobenius^2, then the hard
    # part (q^4 - q^2 + 1)/r as a plain exponent; Frobenius^2 on Fq12 = Fq2[w]/(w^6 - xi) multiplies the w^i
    # coefficient by xi^(i (q^2-1)/6), which lies in Fq
    fexp = (Q_MOD ** 4 - Q_MOD ** 2 + 1) // R_MOD
    assert (Q_MOD ** 4 - Q_MOD ** 2 + 1) % R_MOD == 0
    assert (Q_MOD ** 12 - 1) // R_MOD == (Q_MOD ** 6 - 1) * (Q_MOD ** 2 + 1) * fexp
    fr1 = [f2pow((9, 1), i * (Q_MOD - 1) // 6) for i in range(6)]  # Frobenius on Fq12: conj(a_i) * fr1[i]
    assert fr1[2] == fg2 and fr1[3] == fg3
    fr2 = [f2pow((9, 1), i * (Q_MOD * Q_MOD - 1) // 6) for i in range(6)]
    assert all(c[1] == 0 for c in fr2) and fr2[2] == fg2s
    fwords = [(fexp >> (32 * i)) & 0xffffffff for i in range((fexp.bit_length() + 31) // 32)]
    ate = 6 * 4965661367192848881 + 2
    txt = ["// GENERATED by tools/gen_constants.py -- do not edit.",
           "// BN254 (alt_bn128) constants; Montgomery form with R = 2^256, 8 x 32-bit little-endian limbs.",
           "#pragma once", "#include <stdint.h>", "namespace zkb {",
           params("FrParams", R_MOD, "  static constexpr int TWO_ADICITY = 28;"),
           params("FqParams", Q_MOD),
           f"// omega = 5^((r-1)/2^28): primitive 2^28-th root of unity of Fr, and its inverse (Montgomery form)",
           f"#define ZKB_FR_OMEGA28 {arr(mont(omega, R_MOD))}",
           f"#define ZKB_FR_OMEGA28_INV {arr(mont(pow(omega, -1, R_MOD), R_MOD))}",
           f"// inverse of 2 in Fr (Montgomery form)",
           f"#define ZKB_FR_INV2 {arr(mont(pow(2, -1, R_MOD), R_MOD))}",
           f"// reference base points (Montgomery form): 69*G1::one() and 96*G2::one()",
           f"#define ZKB_G1_BASE_X {arr(mont(B1[0], Q_MOD))}",
           f"#define ZKB_G1_BASE_Y {arr(mont(B1[1], Q_MOD))}",
           f"#define ZKB_G2_BASE_X0 {arr(mont(B2[0][0], Q_MOD))}",
           f"#define ZKB_G2_BASE_X1 {arr(mont(B2[0][1], Q_MOD))}",
           f"#define ZKB_G2_BASE_Y0 {arr(mont(B2[1][0], Q_MOD))}",
           f"#define ZKB_G2_BASE_Y1 {arr(mont(B2[1][1], Q_MOD))}",
           f"// curve coefficients b (G1: 3, G2 twist: 3/(9+u)), Montgomery form",
           f"#define ZKB_G1_B {arr(mont(3, Q_MOD))}",
           f"#define ZKB_G2_B0 {arr(mont(b2[0], Q_MOD))}",
           f"#define ZKB_G2_B1 {arr(mont(b2[1], Q_MOD))}",
           f"// pairing: Frobenius constants (Montgomery form), hard part of the final exponent (q^4-q^2+1)/r ({fexp.bit_length()} bits), ate loop 6u+2 ({ate.bit_length()} bits)",
           "#define ZKB_FROB1_W {" + ", ".join(arr(mont(c[k], Q_MOD)) for c in fr1 for k in (0, 1)) + "}",
           f"#define ZKB_FQ_INV2 {arr(mont(pow(2, -1, Q_MOD), Q_MOD))}",
           f"#define ZKB_BN_U 4965661367192848881ull",
           "#define ZKB_FROB2_W {" + ", ".join(arr(mont(c[0], Q_MOD)) for c in fr2) + "}",
           f"#define ZKB_FINAL_EXP_BITS {fexp.bit_length()}",
           f"#define ZKB_FINAL_EXP_WORDS {len(fwords)}",
           "#define ZKB_FINAL_EXP {" + ", ".join("0x%08xu" % w for w in fwords) + "}",
           f"#define ZKB_ATE_LOOP_BITS {ate.bit_length()}",
           "#define ZKB_ATE_LOOP {" + ", ".join("0x%08xu" % ((ate >> (32 * i)) & 0xffffffff) for i in range(3)) + "}",
           "}  // namespace zkb", ""]
    with open(out, "w") as f:
        f.write("\n".join(txt))

if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(here, "..", "zksnark-rs_b200", "csrc", "constants.h"))
