"""Bit-accurate emulation of the wide (512-bit) product, wide square and stand-alone Montgomery
reduction used by ff.cuh for lazy reduction (Fq2 Karatsuba, squarings, sums of products).

Every PTX instruction is emulated with an explicit carry flag and every "no carry out" assumption
is asserted, so the carry-chain bookkeeping can be checked on the CPU before a kernel runs.
Run: python tools/emul_wide.py
"""
import random

M32 = 0xffffffff


class CC:
    cf = 0


def add_cc(a, b, cin=0):
    t = a + b + cin
    CC.cf = t >> 32
    return t & M32


def mad_lo_cc(a, b, c, cin=0):
    t = ((a * b) & M32) + c + cin
    CC.cf = t >> 32
    return t & M32


def mad_hi_cc(a, b, c, cin=0):
    t = ((a * b) >> 32) + c + cin
    CC.cf = t >> 32
    return t & M32


def limbs(x, n=8):
    return [(x >> (32 * i)) & M32 for i in range(n)]


def unlimbs(v):
    return sum(x << (32 * i) for i, x in enumerate(v))


def chain(acc, off, src, idxs, m):
    """acc[off + 2k, off + 2k + 1] += src[idxs[k]] * m  as one carry chain; the carry out is added into
    acc[off + 2 len] when that limb exists (asserted not to overflow), else asserted zero."""
    cin = 0
    for k, j in enumerate(idxs):
        acc[off + 2 * k] = mad_lo_cc(src[j], m, acc[off + 2 * k], cin)
        acc[off + 2 * k + 1] = mad_hi_cc(src[j], m, acc[off + 2 * k + 1], CC.cf)
        cin = CC.cf
    top = off + 2 * len(idxs)
    if top < len(acc):
        t = acc[top] + cin
        assert t <= M32, "carry-absorbing limb overflowed"
        acc[top] = t
    else:
        assert cin == 0, "carry out of the top"


def merge(E, O):
    """E + (O << 32) over len(E) limbs; O's top limb must be zero."""
    assert O[-1] == 0
    r = [E[0]]
    cf = 0
    for i in range(1, len(E)):
        t = E[i] + O[i - 1] + cf
        r.append(t & M32)
        cf = t >> 32
    assert cf == 0
    return r


def mul_wide(a, b):
    A, B = limbs(a), limbs(b)
    E, O = [0] * 16, [0] * 16
    for i in range(8):
        if i % 2 == 0:
            chain(E, i, A, (0, 2, 4, 6), B[i])
            chain(O, i, A, (1, 3, 5, 7), B[i])
        else:
            chain(O, i - 1, A, (0, 2, 4, 6), B[i])
            chain(E, i + 1, A, (1, 3, 5, 7), B[i])
    return merge(E, O)


def mul_half(A, B):
    """4 x 4 limbs -> 8 limbs with the same even / odd carry chains (the three products of mul_wide_k)."""
    E, O = [0] * 9, [0] * 8
    chain(E, 0, A, (0, 2), B[0])
    chain(O, 0, A, (1, 3), B[0])
    chain(O, 0, A, (0, 2), B[1])
    chain(E, 2, A, (1, 3), B[1])
    chain(E, 2, A, (0, 2), B[2])
    chain(O, 2, A, (1, 3), B[2])
    chain(O, 2, A, (0, 2), B[3])
    chain(E, 4, A, (1, 3), B[3])
    assert E[8] == 0 and O[7] == 0
    return merge(E[:8], O)


def add_n(x, y, cin=0):
    r, cf = [], cin
    for a, b in zip(x, y):
        t = a + b + cf
        r.append(t & M32)
        cf = t >> 32
    return r, cf


def sub_n(x, y, bin_=0):
    r, bf = [], bin_
    for a, b in zip(x, y):
        t = a - b - bf
        r.append(t & M32)
        bf = 1 if t < 0 else 0
    return r, bf


def mul_wide_k(a, b):
    """One level of Karatsuba: 48 limb products instead of 64.  a = a0 + a1 B, b = b0 + b1 B (B = 2^128):
    z0 = a0 b0, z2 = a1 b1, zm = (a0 + a1)(b0 + b1) - z0 - z2 (9 limbs), r = z0 + zm B + z2 B^2."""
    A, Bv = limbs(a), limbs(b)
    z0 = mul_half(A[:4], Bv[:4])
    z2 = mul_half(A[4:], Bv[4:])
    sa, ca = add_n(A[:4], A[4:])
    sb, cb = add_n(Bv[:4], Bv[4:])
    zm = mul_half(sa, sb) + [0]
    # + ca * sb * B + cb * sa * B + ca * cb * B^2
    t = [x & (M32 if ca else 0) for x in sb]
    u = [x & (M32 if cb else 0) for x in sa]
    hi, c = add_n(zm[4:8], t)
    top = zm[8] + c
    hi, c = add_n(hi, u)
    top += c + (ca & cb)
    assert top <= 7
    zm = zm[:4] + hi + [top]
    zm, bf = sub_n(zm, z0 + [0])
    assert bf == 0
    zm, bf = sub_n(zm, z2 + [0])
    assert bf == 0 and zm[8] <= 1
    r = z0[:4]
    mid, c1 = add_n(z0[4:], zm[:4])
    hi2, c2 = add_n(z2[:4], zm[4:8], c1)
    top2, c3 = add_n(z2[4:], [zm[8], 0, 0, 0], c2)
    assert c3 == 0
    return r + mid + hi2 + top2


def sqr_wide(a):
    A = limbs(a)
    E, O = [0] * 16, [0] * 16
    for i in range(7):
        odd = list(range(i + 1, 8, 2))   # i + j odd  -> O at index i + j - 1 = 2 i, ...
        even = list(range(i + 2, 8, 2))  # i + j even -> E at index i + j = 2 i + 2, ...
        chain(O, 2 * i, A, odd, A[i])
        if even:
            chain(E, 2 * i + 2, A, even, A[i])
    S = merge(E, O)
    # double (funnel shifts), then add the diagonal a_i^2 at limbs (2i, 2i+1)
    assert S[15] >> 31 == 0
    D = [((S[i] << 1) | (S[i - 1] >> 31 if i else 0)) & M32 for i in range(16)]
    r, cf = [], 0
    for i in range(8):
        sq = A[i] * A[i]
        t = D[2 * i] + (sq & M32) + cf
        r.append(t & M32)
        cf = t >> 32
        t = D[2 * i + 1] + (sq >> 32) + cf
        r.append(t & M32)
        cf = t >> 32
    assert cf == 0
    return r


def redc(T, p, inv):
    """T: 16 limbs, value < p * 2^256.  Returns (T / 2^256) mod p, computed as in ff.cuh: state
    t = E + (O << 32); round: m = E[0] * inv, O += m * p_odd, E += m * p_even (carry -> O[7]);
    shift: E' = O (+ E[1] at limb 0), O' = E >> 64 with T[8 + i] entering at limb 7 (= O'[6])."""
    P = limbs(p)
    E, O = list(T[:8]), [0] * 8
    for i in range(8):
        m = (E[0] * inv) & M32
        # O += m * p_odd   (no carry out)
        cin = 0
        for k, j in enumerate((1, 3, 5, 7)):
            O[2 * k] = mad_lo_cc(P[j], m, O[2 * k], cin)
            O[2 * k + 1] = mad_hi_cc(P[j], m, O[2 * k + 1], CC.cf)
            cin = CC.cf
        assert cin == 0
        # E += m * p_even  (carry out -> O[7])
        cin = 0
        for k, j in enumerate((0, 2, 4, 6)):
            E[2 * k] = mad_lo_cc(P[j], m, E[2 * k], cin)
            E[2 * k + 1] = mad_hi_cc(P[j], m, E[2 * k + 1], CC.cf)
            cin = CC.cf
        t = O[7] + cin
        assert t <= M32
        O[7] = t
        assert E[0] == 0
        if i == 7:
            break
        # shift by one limb and bring in T[8 + i] at limb 7
        nE = list(O)
        nE[0] = add_cc(O[0], E[1])
        # the carry of limb 0 goes to limb 1 = nO[0]
        nO = [0] * 8
        cf = CC.cf
        for k in range(6):
            nO[k] = add_cc(E[k + 2], 0, cf)
            cf = CC.cf
        nO[6] = add_cc(T[8 + i], 0, cf)
        nO[7] = CC.cf
        E, O = nE, nO
    # result = (E >> 32) + O + (T[15] << 224)
    r, cf = [], 0
    for k in range(7):
        t = E[k + 1] + O[k] + cf
        r.append(t & M32)
        cf = t >> 32
    t = O[7] + T[15] + cf
    assert t <= M32
    r.append(t)
    v = unlimbs(r)
    assert v < 2 * p
    return v - p if v >= p else v


if __name__ == "__main__":
    r_mod = 21888242871839275222246405745257275088548364400416034343698204186575808495617
    q_mod = 21888242871839275222246405745257275088696311157297823662689037894645226208583
    rng = random.Random(2)
    R = 1 << 256
    for p in (r_mod, q_mod):
        inv = (-pow(p, -1, 1 << 32)) % (1 << 32)
        Rinv = pow(R, -1, p)
        edge = [0, 1, p - 1, p - 2, (1 << 255) - 1, R - 1, (1 << 254), M32, R - (1 << 224), (1 << 128) - 1, ((1 << 128) - 1) << 128,
                (1 << 128), R - (1 << 128) - 1, ((1 << 127) << 128) | (1 << 127)]
        cases = [(a, b) for a in edge for b in edge] + [(rng.randrange(R), rng.randrange(R)) for _ in range(3000)]
        for a, b in cases:
            assert unlimbs(mul_wide(a, b)) == a * b, (a, b)
            assert unlimbs(mul_wide_k(a, b)) == a * b, (a, b)
            assert unlimbs(sqr_wide(a)) == a * a, a
        red = [0, 1, p - 1, p * R - 1, p * R - p, (p - 1) * (p - 1), 2 * (p - 1) * (p - 1), p * p + (p - 1) ** 2]
        red += [rng.randrange(p * R) for _ in range(20000)]
        for t in red:
            assert t < p * R
            assert redc(limbs(t, 16), p, inv) == t * Rinv % p, t
    print("wide mul / sqr / redc emulation OK")
