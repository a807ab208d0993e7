"""Bit-accurate emulation of the even/odd 32-bit-limb Montgomery multiply used in ff.cuh.

Emulates every PTX instruction (mad.lo.cc / madc.hi.cc / addc ...) with an explicit carry flag so
the carry-chain bookkeeping (and the no-overflow assumptions for 254-bit moduli) can be checked on
the CPU before the kernel ever runs on a GPU.
"""
import random, sys
M32 = 0xffffffff

class CC:
    cf = 0

def mad_lo_cc(a, b, c, cin=0):
    t = ((a * b) & M32) + c + cin
    CC.cf = t >> 32
    return t & M32

def mad_hi_cc(a, b, c, cin=0):
    t = ((a * b) >> 32) + c + cin
    CC.cf = t >> 32
    return t & M32

def limbs(x): return [(x >> (32 * i)) & M32 for i in range(8)]
def unlimbs(v): return sum(x << (32 * i) for i, x in enumerate(v))

def montmul(a, b, p, inv, check=True):
    A, B, P = limbs(a), limbs(b), limbs(p)
    X = [0] * 8  # "even"
    Y = [0] * 8  # "odd"
    def cmad(acc, src, off, m):
        # acc += sum_{j in off,off+2,..} src[j]*m * 2^(32*(j-off)); returns carry-out
        acc[0] = mad_lo_cc(src[off], m, acc[0], 0)
        acc[1] = mad_hi_cc(src[off], m, acc[1], CC.cf)
        for j in (2, 4, 6):
            acc[j] = mad_lo_cc(src[off + j], m, acc[j], CC.cf)
            acc[j + 1] = mad_hi_cc(src[off + j], m, acc[j + 1], CC.cf)
        return CC.cf
    for i in range(8):
        bi = B[i]
        if i == 0:
            for j in (0, 2, 4, 6):
                pr = A[j] * bi; X[j], X[j + 1] = pr & M32, pr >> 32
                pr = A[j + 1] * bi; Y[j], Y[j + 1] = pr & M32, pr >> 32
        else:
            # roles: X is even-aligned, Y odd-aligned *after* the swap done at the loop bottom
            t = X[0] + Y[1]; X[0] = t & M32; CC.cf = t >> 32
            for j in (0, 2, 4):
                Y[j] = mad_lo_cc(A[j + 1], bi, Y[j + 2], CC.cf)
                Y[j + 1] = mad_hi_cc(A[j + 1], bi, Y[j + 3], CC.cf)
            Y[6] = mad_lo_cc(A[7], bi, 0, CC.cf)
            Y[7] = mad_hi_cc(A[7], bi, 0, CC.cf)
            if check: assert CC.cf == 0
            c = cmad(X, A, 0, bi)
            t = Y[7] + c
            if check: assert t <= M32
            Y[7] = t & M32
        mi = (X[0] * inv) & M32
        c = cmad(Y, P, 1, mi)
        if check: assert c == 0
        c = cmad(X, P, 0, mi)
        t = Y[7] + c
        if check: assert t <= M32
        Y[7] = t & M32
        assert X[0] == 0
        X, Y = Y, X  # shift by one limb: old odd becomes even-aligned
    # after the last swap: Y is the array whose limb0 == 0 (old X); result = X + Y[1..7]
    t = X[0] + Y[1]; R = [t & M32]; cf = t >> 32
    for i in range(1, 7):
        t = X[i] + Y[i + 1] + cf; R.append(t & M32); cf = t >> 32
    t = X[7] + cf
    if check: assert t <= M32
    R.append(t & M32)
    r = unlimbs(R)
    if check: assert r < 2 * p
    return r - p if r >= p else r

if __name__ == "__main__":
    r_mod = 21888242871839275222246405745257275088548364400416034343698204186575808495617
    q_mod = 21888242871839275222246405745257275088696311157297823662689037894645226208583
    rng = random.Random(1)
    for p in (r_mod, q_mod):
        inv = (-pow(p, -1, 1 << 32)) % (1 << 32)
        Rinv = pow(1 << 256, -1, p)
        cases = [(0, 0), (p - 1, p - 1), (1, p - 1), (p - 1, 1), ((1 << 256) % p, (1 << 256) % p)]
        cases += [(rng.randrange(p), rng.randrange(p)) for _ in range(20000)]
        for a, b in cases:
            assert montmul(a, b, p, inv) == a * b * Rinv % p, (a, b)
    print("montmul emulation OK")
