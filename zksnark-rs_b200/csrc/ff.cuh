// 256-bit prime-field arithmetic for BN254 Fr / Fq on sm_100a.
//
// Replaces the arithmetic of crate `bn` 0.4.3 (Fr, Fq, Fq2) that the reference reaches through
// /root/reference/src/groth16/fr.rs:18-71 (FrLocal + - * / neg) and, via G1/G2, fr.rs:101-123.
//
// Representation: 8 x 32-bit little-endian limbs in Montgomery form (R = 2^256), fully reduced
// (< p) between operations.  The device multiplier is an interleaved (CIOS) Montgomery product on
// two staggered accumulators ("even" / "odd" 64-bit columns) so that every partial product is one
// 32x32+64 wide multiply-add (IMAD.WIDE.U32 with carry) -- no tensor cores, this is integer work.
// The carry-chain bookkeeping is validated bit-for-bit on the CPU by tools/emul_montmul.py.
//
// The same header compiles for the host (portable unsigned __int128 path) so the host side of the
// library (constant setup, twiddle generation) shares one implementation; the host path is never
// used to produce a result the GPU path is supposed to produce.
#pragma once
#include <stdint.h>
#include "constants.h"

#if defined(__CUDACC__)
#define ZKB_HD __host__ __device__ __forceinline__
#define ZKB_D __device__ __forceinline__
#else
#define ZKB_HD inline
#define ZKB_D inline
#endif

namespace zkb {

template <class P>
struct alignas(16) Fp {
  uint32_t v[8];

  ZKB_HD static Fp zero() {
    Fp r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = 0;
    return r;
  }
  ZKB_HD static Fp one() {
    Fp r;
    r.v[0] = P::ONE0; r.v[1] = P::ONE1; r.v[2] = P::ONE2; r.v[3] = P::ONE3;
    r.v[4] = P::ONE4; r.v[5] = P::ONE5; r.v[6] = P::ONE6; r.v[7] = P::ONE7;
    return r;
  }
  ZKB_HD static Fp r2() {
    Fp r;
    r.v[0] = P::R20; r.v[1] = P::R21; r.v[2] = P::R22; r.v[3] = P::R23;
    r.v[4] = P::R24; r.v[5] = P::R25; r.v[6] = P::R26; r.v[7] = P::R27;
    return r;
  }
  ZKB_HD static Fp modulus() {
    Fp r;
    r.v[0] = P::P0; r.v[1] = P::P1; r.v[2] = P::P2; r.v[3] = P::P3;
    r.v[4] = P::P4; r.v[5] = P::P5; r.v[6] = P::P6; r.v[7] = P::P7;
    return r;
  }
  ZKB_HD bool is_zero() const {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) o |= v[i];
    return o == 0;
  }
  ZKB_HD bool operator==(const Fp& b) const {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) o |= v[i] ^ b.v[i];
    return o == 0;
  }
  ZKB_HD bool operator!=(const Fp& b) const { return !(*this == b); }
};

// ------------------------------------------------------------------------------------------------
// host (portable) primitives
// ------------------------------------------------------------------------------------------------
namespace host_impl {
typedef unsigned __int128 u128;

template <class P>
inline void load64(const Fp<P>& a, uint64_t o[4]) {
  for (int i = 0; i < 4; i++) o[i] = (uint64_t)a.v[2 * i] | ((uint64_t)a.v[2 * i + 1] << 32);
}
template <class P>
inline Fp<P> store64(const uint64_t o[4]) {
  Fp<P> r;
  for (int i = 0; i < 4; i++) { r.v[2 * i] = (uint32_t)o[i]; r.v[2 * i + 1] = (uint32_t)(o[i] >> 32); }
  return r;
}
template <class P>
inline void mod64(uint64_t p[4]) { Fp<P> m = Fp<P>::modulus(); load64(m, p); }

// returns borrow of a - b
inline uint64_t sub4(uint64_t r[4], const uint64_t a[4], const uint64_t b[4]) {
  u128 br = 0;
  for (int i = 0; i < 4; i++) {
    u128 t = (u128)a[i] - b[i] - br;
    r[i] = (uint64_t)t;
    br = (t >> 64) & 1;
  }
  return (uint64_t)br;
}
inline uint64_t add4(uint64_t r[4], const uint64_t a[4], const uint64_t b[4]) {
  u128 c = 0;
  for (int i = 0; i < 4; i++) {
    u128 t = (u128)a[i] + b[i] + c;
    r[i] = (uint64_t)t;
    c = t >> 64;
  }
  return (uint64_t)c;
}
template <class P>
inline Fp<P> add(const Fp<P>& x, const Fp<P>& y) {
  uint64_t a[4], b[4], p[4], s[4], d[4];
  load64(x, a); load64(y, b); mod64<P>(p);
  add4(s, a, b);  // no overflow: a,b < p < 2^254
  uint64_t br = sub4(d, s, p);
  return store64<P>(br ? s : d);
}
template <class P>
inline Fp<P> sub(const Fp<P>& x, const Fp<P>& y) {
  uint64_t a[4], b[4], p[4], d[4], e[4];
  load64(x, a); load64(y, b); mod64<P>(p);
  uint64_t br = sub4(d, a, b);
  add4(e, d, p);
  return store64<P>(br ? e : d);
}
template <class P>
inline Fp<P> mul(const Fp<P>& x, const Fp<P>& y) {
  uint64_t a[4], b[4], p[4];
  load64(x, a); load64(y, b); mod64<P>(p);
  // -p^-1 mod 2^64 from the 32-bit constant by one Newton step
  uint64_t inv = P::INV;            // -p^-1 mod 2^32
  uint64_t pinv = (uint64_t)0 - inv;  // p^-1 mod 2^32 (as 64-bit: valid mod 2^32)
  pinv = pinv * (2 - p[0] * pinv);    // now valid mod 2^64
  uint64_t ninv = (uint64_t)0 - pinv;
  uint64_t t[6] = {0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 4; i++) {
    u128 c = 0;
    for (int j = 0; j < 4; j++) {
      u128 s = (u128)a[j] * b[i] + t[j] + c;
      t[j] = (uint64_t)s; c = s >> 64;
    }
    u128 s = (u128)t[4] + c; t[4] = (uint64_t)s; t[5] = (uint64_t)(s >> 64);
    uint64_t m = t[0] * ninv;
    c = ((u128)m * p[0] + t[0]) >> 64;
    for (int j = 1; j < 4; j++) {
      u128 s2 = (u128)m * p[j] + t[j] + c;
      t[j - 1] = (uint64_t)s2; c = s2 >> 64;
    }
    s = (u128)t[4] + c; t[3] = (uint64_t)s; t[4] = t[5] + (uint64_t)(s >> 64);
  }
  uint64_t d[4];
  uint64_t br = sub4(d, t, p);
  return store64<P>((t[4] == 0 && br) ? t : d);
}
}  // namespace host_impl

// ------------------------------------------------------------------------------------------------
// device primitives (PTX carry chains)
// ------------------------------------------------------------------------------------------------
#if defined(__CUDA_ARCH__)
namespace dev_impl {

// acc[0..7] += {s0,s2,s4,s6} * m on 64-bit columns (0,1),(2,3),(4,5),(6,7); carry-out added to *top
#define ZKB_CMAD_TOP(acc, top, s0, s2, s4, s6, m)                                                  \
  asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\t"                                                         \
      "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"                                                        \
      "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"                                                       \
      "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"                                                       \
      "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"                                                       \
      "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"                                                       \
      "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"                                                       \
      "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"                                                       \
      "addc.u32 %8, %8, 0;"                                                                        \
      : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]),        \
        "+r"(acc[6]), "+r"(acc[7]), "+r"(top)                                                      \
      : "r"(s0), "r"(s2), "r"(s4), "r"(s6), "r"(m))

// same without a carry-out (the caller guarantees none: top limb has >= 2 spare bits)
#define ZKB_CMAD(acc, s0, s2, s4, s6, m)                                                           \
  asm("mad.lo.cc.u32 %0, %8, %12, %0;\n\t"                                                         \
      "madc.hi.cc.u32 %1, %8, %12, %1;\n\t"                                                        \
      "madc.lo.cc.u32 %2, %9, %12, %2;\n\t"                                                        \
      "madc.hi.cc.u32 %3, %9, %12, %3;\n\t"                                                        \
      "madc.lo.cc.u32 %4, %10, %12, %4;\n\t"                                                       \
      "madc.hi.cc.u32 %5, %10, %12, %5;\n\t"                                                       \
      "madc.lo.cc.u32 %6, %11, %12, %6;\n\t"                                                       \
      "madc.hi.u32 %7, %11, %12, %7;"                                                              \
      : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]),        \
        "+r"(acc[6]), "+r"(acc[7])                                                                 \
      : "r"(s0), "r"(s2), "r"(s4), "r"(s6), "r"(m))

// X[0] += Y[1] (carry c); Y = (Y >> 64) + {s1,s3,s5,s7} * m + c
#define ZKB_SHIFT_MAD(X0, Y, s1, s3, s5, s7, m)                                                    \
  asm("add.cc.u32 %0, %0, %2;\n\t"                                                                 \
      "madc.lo.cc.u32 %1, %9, %13, %3;\n\t"                                                        \
      "madc.hi.cc.u32 %2, %9, %13, %4;\n\t"                                                        \
      "madc.lo.cc.u32 %3, %10, %13, %5;\n\t"                                                       \
      "madc.hi.cc.u32 %4, %10, %13, %6;\n\t"                                                       \
      "madc.lo.cc.u32 %5, %11, %13, %7;\n\t"                                                       \
      "madc.hi.cc.u32 %6, %11, %13, %8;\n\t"                                                       \
      "madc.lo.cc.u32 %7, %12, %13, 0;\n\t"                                                        \
      "madc.hi.u32 %8, %12, %13, 0;"                                                               \
      : "+r"(X0), "+r"(Y[0]), "+r"(Y[1]), "+r"(Y[2]), "+r"(Y[3]), "+r"(Y[4]), "+r"(Y[5]),          \
        "+r"(Y[6]), "+r"(Y[7])                                                                     \
      : "r"(s1), "r"(s3), "r"(s5), "r"(s7), "r"(m))

template <class P>
__device__ __forceinline__ void mont_round(uint32_t (&X)[8], uint32_t (&Y)[8], const uint32_t* a,
                                           uint32_t bi, bool first) {
  if (first) {
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
      asm("mul.lo.u32 %0, %2, %3;\n\tmul.hi.u32 %1, %2, %3;" : "=&r"(X[j]), "=&r"(X[j + 1]) : "r"(a[j]), "r"(bi));
      asm("mul.lo.u32 %0, %2, %3;\n\tmul.hi.u32 %1, %2, %3;" : "=&r"(Y[j]), "=&r"(Y[j + 1]) : "r"(a[j + 1]), "r"(bi));
    }
  } else {
    ZKB_SHIFT_MAD(X[0], Y, a[1], a[3], a[5], a[7], bi);
    ZKB_CMAD_TOP(X, Y[7], a[0], a[2], a[4], a[6], bi);
  }
  uint32_t mi = X[0] * P::INV;
  const uint32_t p0 = P::P0, p1 = P::P1, p2 = P::P2, p3 = P::P3, p4 = P::P4, p5 = P::P5, p6 = P::P6, p7 = P::P7;
  ZKB_CMAD(Y, p1, p3, p5, p7, mi);
  ZKB_CMAD_TOP(X, Y[7], p0, p2, p4, p6, mi);
}

// r = r - p if r >= p   (r < 2p)
template <class P>
__device__ __forceinline__ void final_sub(uint32_t (&r)[8]) {
  uint32_t t[8], br;
  asm("sub.cc.u32 %0, %9, %17;\n\t"
      "subc.cc.u32 %1, %10, %18;\n\t"
      "subc.cc.u32 %2, %11, %19;\n\t"
      "subc.cc.u32 %3, %12, %20;\n\t"
      "subc.cc.u32 %4, %13, %21;\n\t"
      "subc.cc.u32 %5, %14, %22;\n\t"
      "subc.cc.u32 %6, %15, %23;\n\t"
      "subc.cc.u32 %7, %16, %24;\n\t"
      "subc.u32 %8, 0, 0;"
      : "=&r"(t[0]), "=&r"(t[1]), "=&r"(t[2]), "=&r"(t[3]), "=&r"(t[4]), "=&r"(t[5]), "=&r"(t[6]), "=&r"(t[7]), "=&r"(br)
      : "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(P::P0), "r"(P::P1), "r"(P::P2), "r"(P::P3), "r"(P::P4), "r"(P::P5), "r"(P::P6), "r"(P::P7));
  // br == 0xffffffff when r < p (keep r), 0 when r >= p (take t)
#pragma unroll
  for (int i = 0; i < 8; i++) r[i] = br ? r[i] : t[i];
}

template <class P>
__device__ __forceinline__ Fp<P> mul(const Fp<P>& a, const Fp<P>& b) {
  uint32_t X[8], Y[8];
  mont_round<P>(X, Y, a.v, b.v[0], true);
  mont_round<P>(Y, X, a.v, b.v[1], false);
  mont_round<P>(X, Y, a.v, b.v[2], false);
  mont_round<P>(Y, X, a.v, b.v[3], false);
  mont_round<P>(X, Y, a.v, b.v[4], false);
  mont_round<P>(Y, X, a.v, b.v[5], false);
  mont_round<P>(X, Y, a.v, b.v[6], false);
  mont_round<P>(Y, X, a.v, b.v[7], false);
  // after the last round Y[0] == 0 and the value is X + (Y >> 32)
  asm("add.cc.u32 %0, %0, %8;\n\t"
      "addc.cc.u32 %1, %1, %9;\n\t"
      "addc.cc.u32 %2, %2, %10;\n\t"
      "addc.cc.u32 %3, %3, %11;\n\t"
      "addc.cc.u32 %4, %4, %12;\n\t"
      "addc.cc.u32 %5, %5, %13;\n\t"
      "addc.cc.u32 %6, %6, %14;\n\t"
      "addc.u32 %7, %7, 0;"
      : "+r"(X[0]), "+r"(X[1]), "+r"(X[2]), "+r"(X[3]), "+r"(X[4]), "+r"(X[5]), "+r"(X[6]), "+r"(X[7])
      : "r"(Y[1]), "r"(Y[2]), "r"(Y[3]), "r"(Y[4]), "r"(Y[5]), "r"(Y[6]), "r"(Y[7]));
  final_sub<P>(X);
  Fp<P> r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = X[i];
  return r;
}

// ------------------------------------------------------------------------------------------------
// Wide (512-bit) products and a stand-alone Montgomery reduction, for LAZY reduction: several
// products are summed as 16-limb integers and reduced once (Fq2 Karatsuba: 3 products + 2
// reductions instead of 3 + 3; a*b - c*d: 2 + 1 instead of 2 + 2), and squarings use the 36
// distinct limb products instead of 64.  Same even/odd accumulator idea as mul(): E collects the
// limb products a_i b_j with i+j even, O (worth 2^32 more) those with i+j odd, so every product is
// one IMAD.WIDE.U32.X in a carry chain; each chain's carry out is absorbed by the next limb up.
// Bookkeeping validated instruction by instruction in tools/emul_wide.py.

// acc[0..2n-1] += {s...} * m (n = 1..3 pairs), carry out added to `top`
#define ZKB_WCHAIN1(acc, top, s0, m)                                                               \
  asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\t"                                                         \
      "madc.hi.cc.u32 %1, %3, %4, %1;\n\t"                                                        \
      "addc.u32 %2, %2, 0;"                                                                        \
      : "+r"(acc[0]), "+r"(acc[1]), "+r"(top)                                                      \
      : "r"(s0), "r"(m))
#define ZKB_WCHAIN2(acc, top, s0, s1, m)                                                           \
  asm("mad.lo.cc.u32 %0, %5, %7, %0;\n\t"                                                         \
      "madc.hi.cc.u32 %1, %5, %7, %1;\n\t"                                                        \
      "madc.lo.cc.u32 %2, %6, %7, %2;\n\t"                                                        \
      "madc.hi.cc.u32 %3, %6, %7, %3;\n\t"                                                        \
      "addc.u32 %4, %4, 0;"                                                                        \
      : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(top)                          \
      : "r"(s0), "r"(s1), "r"(m))
#define ZKB_WCHAIN3(acc, top, s0, s1, s2, m)                                                       \
  asm("mad.lo.cc.u32 %0, %7, %10, %0;\n\t"                                                        \
      "madc.hi.cc.u32 %1, %7, %10, %1;\n\t"                                                       \
      "madc.lo.cc.u32 %2, %8, %10, %2;\n\t"                                                       \
      "madc.hi.cc.u32 %3, %8, %10, %3;\n\t"                                                       \
      "madc.lo.cc.u32 %4, %9, %10, %4;\n\t"                                                       \
      "madc.hi.cc.u32 %5, %9, %10, %5;\n\t"                                                       \
      "addc.u32 %6, %6, 0;"                                                                        \
      : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(top) \
      : "r"(s0), "r"(s1), "r"(s2), "r"(m))

// r[0..7] = x[0..7] + y[0..7] + cin, cout = carry out   (cin, cout are 0 / 1 registers)
#define ZKB_ADD8(r, x, y, cin, cout)                                                               \
  asm("{\n\t.reg .u32 t;\n\t"                                                                    \
      "add.cc.u32 t, %25, 0xffffffff;\n\t"                                                        \
      "addc.cc.u32 %0, %9, %17;\n\t"                                                              \
      "addc.cc.u32 %1, %10, %18;\n\t"                                                             \
      "addc.cc.u32 %2, %11, %19;\n\t"                                                             \
      "addc.cc.u32 %3, %12, %20;\n\t"                                                             \
      "addc.cc.u32 %4, %13, %21;\n\t"                                                             \
      "addc.cc.u32 %5, %14, %22;\n\t"                                                             \
      "addc.cc.u32 %6, %15, %23;\n\t"                                                             \
      "addc.cc.u32 %7, %16, %24;\n\t"                                                             \
      "addc.u32 %8, 0, 0;\n\t}"                                                                   \
      : "=&r"(r[0]), "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3]), "=&r"(r[4]), "=&r"(r[5]), "=&r"(r[6]), "=&r"(r[7]), "=&r"(cout) \
      : "r"(x[0]), "r"(x[1]), "r"(x[2]), "r"(x[3]), "r"(x[4]), "r"(x[5]), "r"(x[6]), "r"(x[7]),    \
        "r"(y[0]), "r"(y[1]), "r"(y[2]), "r"(y[3]), "r"(y[4]), "r"(y[5]), "r"(y[6]), "r"(y[7]), "r"(cin))
// r = x - y - bin, bout = borrow out
#define ZKB_SUB8(r, x, y, bin, bout)                                                               \
  asm("{\n\t.reg .u32 t;\n\t"                                                                    \
      "sub.cc.u32 t, 0, %25;\n\t"                                                                 \
      "subc.cc.u32 %0, %9, %17;\n\t"                                                              \
      "subc.cc.u32 %1, %10, %18;\n\t"                                                             \
      "subc.cc.u32 %2, %11, %19;\n\t"                                                             \
      "subc.cc.u32 %3, %12, %20;\n\t"                                                             \
      "subc.cc.u32 %4, %13, %21;\n\t"                                                             \
      "subc.cc.u32 %5, %14, %22;\n\t"                                                             \
      "subc.cc.u32 %6, %15, %23;\n\t"                                                             \
      "subc.cc.u32 %7, %16, %24;\n\t"                                                             \
      "subc.u32 %8, 0, 0;\n\t"                                                                    \
      "and.b32 %8, %8, 1;\n\t}"                                                                   \
      : "=&r"(r[0]), "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3]), "=&r"(r[4]), "=&r"(r[5]), "=&r"(r[6]), "=&r"(r[7]), "=&r"(bout) \
      : "r"(x[0]), "r"(x[1]), "r"(x[2]), "r"(x[3]), "r"(x[4]), "r"(x[5]), "r"(x[6]), "r"(x[7]),    \
        "r"(y[0]), "r"(y[1]), "r"(y[2]), "r"(y[3]), "r"(y[4]), "r"(y[5]), "r"(y[6]), "r"(y[7]), "r"(bin))

// 16-limb add / subtract (callers guarantee no carry / borrow out of the top)
__device__ __forceinline__ void wide_add(uint32_t (&r)[16], const uint32_t (&x)[16], const uint32_t (&y)[16]) {
  uint32_t c, c2;
  const uint32_t zero = 0;
  ZKB_ADD8(r, x, y, zero, c);
  ZKB_ADD8((r + 8), (x + 8), (y + 8), c, c2);
}
__device__ __forceinline__ void wide_sub(uint32_t (&r)[16], const uint32_t (&x)[16], const uint32_t (&y)[16]) {
  uint32_t b, b2;
  const uint32_t zero = 0;
  ZKB_SUB8(r, x, y, zero, b);
  ZKB_SUB8((r + 8), (x + 8), (y + 8), b, b2);
}
// 8-limb add without reduction (a + b < 2^256 guaranteed by the caller: both < p < 2^254)
__device__ __forceinline__ void add_nored(uint32_t (&r)[8], const uint32_t (&x)[8], const uint32_t (&y)[8]) {
  uint32_t c;
  const uint32_t zero = 0;
  ZKB_ADD8(r, x, y, zero, c);
}

// r = E + (O << 32)
__device__ __forceinline__ void wide_merge(uint32_t (&r)[16], const uint32_t* E, const uint32_t* O) {
  uint32_t c, c2;
  const uint32_t zero = 0;
  uint32_t lo[8] = {0, O[0], O[1], O[2], O[3], O[4], O[5], O[6]};
  uint32_t hi[8] = {O[7], O[8], O[9], O[10], O[11], O[12], O[13], O[14]};
  ZKB_ADD8(r, E, lo, zero, c);
  ZKB_ADD8((r + 8), (E + 8), hi, c, c2);
}

// r = a * b (both < 2^256)
__device__ __forceinline__ void mul_wide(uint32_t (&r)[16], const uint32_t (&a)[8], const uint32_t (&b)[8]) {
  uint32_t E[17], O[16];
#pragma unroll
  for (int i = 0; i < 17; i++) E[i] = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) O[i] = 0;
#pragma unroll
  for (int i = 0; i < 8; i += 2) {
    ZKB_CMAD_TOP((E + i), E[i + 8], a[0], a[2], a[4], a[6], b[i]);
    ZKB_CMAD_TOP((O + i), O[i + 8], a[1], a[3], a[5], a[7], b[i]);
    ZKB_CMAD_TOP((O + i), O[i + 8], a[0], a[2], a[4], a[6], b[i + 1]);
    ZKB_CMAD_TOP((E + i + 2), E[i + 10], a[1], a[3], a[5], a[7], b[i + 1]);  // E[16] stays 0 (product < 2^512)
  }
  wide_merge(r, E, O);
}

// ---- one level of Karatsuba on the wide product: 48 limb products instead of 64 ------------------------------------
// a = a0 + a1 B, b = b0 + b1 B (B = 2^128):  z0 = a0 b0,  z2 = a1 b1,  zm = (a0 + a1)(b0 + b1) - z0 - z2 = a0 b1 + a1 b0
// (9 limbs),  a b = z0 + zm B + z2 B^2.  The idea: the accumulation kernels are bound by the multiplier pipe at ~40 % issue
// utilisation, so 16 fewer IMAD.WIDE for ~70 more integer adds should pay.  MEASURED ON B200 (zkb_bench_modmul fields
// 1 / 2 / 3, profiles/r02_notes.md): it does not -- CIOS 67.0 G modmul/s, wide product + reduction 56.5, Karatsuba wide
// product + reduction 54.5-55.0: ptxas schedules a large share of the extra adds as IMAD forms on the same pipe (SASS per
// product: 120 IMAD.WIDE + 16 IMAD + 82 other for CIOS; 102 + 58 + 202 here).  Kept as a validated primitive (field-op
// hook 7, bench field 3; -DZKB_KARATSUBA routes the Fq2 products through it); not used by any kernel.  Bookkeeping
// validated in tools/emul_wide.py (mul_wide_k).
// r[0..7] = a[0..3] * b[0..3]
__device__ __forceinline__ void mul_half(uint32_t (&r)[8], const uint32_t* a, const uint32_t* b) {
  uint32_t E[9], O[8];
#pragma unroll
  for (int i = 0; i < 9; i++) E[i] = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) O[i] = 0;
  ZKB_WCHAIN2((E + 0), E[4], a[0], a[2], b[0]);
  ZKB_WCHAIN2((O + 0), O[4], a[1], a[3], b[0]);
  ZKB_WCHAIN2((O + 0), O[4], a[0], a[2], b[1]);
  ZKB_WCHAIN2((E + 2), E[6], a[1], a[3], b[1]);
  ZKB_WCHAIN2((E + 2), E[6], a[0], a[2], b[2]);
  ZKB_WCHAIN2((O + 2), O[6], a[1], a[3], b[2]);
  ZKB_WCHAIN2((O + 2), O[6], a[0], a[2], b[3]);
  ZKB_WCHAIN2((E + 4), E[8], a[1], a[3], b[3]);  // E[8] and O[7] stay 0 (product < 2^256)
  r[0] = E[0];
  asm("add.cc.u32 %0, %7, %14;\n\t"
      "addc.cc.u32 %1, %8, %15;\n\t"
      "addc.cc.u32 %2, %9, %16;\n\t"
      "addc.cc.u32 %3, %10, %17;\n\t"
      "addc.cc.u32 %4, %11, %18;\n\t"
      "addc.cc.u32 %5, %12, %19;\n\t"
      "addc.u32 %6, %13, %20;"
      : "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3]), "=&r"(r[4]), "=&r"(r[5]), "=&r"(r[6]), "=&r"(r[7])
      : "r"(E[1]), "r"(E[2]), "r"(E[3]), "r"(E[4]), "r"(E[5]), "r"(E[6]), "r"(E[7]),
        "r"(O[0]), "r"(O[1]), "r"(O[2]), "r"(O[3]), "r"(O[4]), "r"(O[5]), "r"(O[6]));
}
// r[0..3] = x + y (+ cin), cout
#define ZKB_ADD4(r, x, y, cin, cout)                                                               \
  asm("{\n\t.reg .u32 t;\n\t"                                                                    \
      "add.cc.u32 t, %13, 0xffffffff;\n\t"                                                        \
      "addc.cc.u32 %0, %5, %9;\n\t"                                                               \
      "addc.cc.u32 %1, %6, %10;\n\t"                                                              \
      "addc.cc.u32 %2, %7, %11;\n\t"                                                              \
      "addc.cc.u32 %3, %8, %12;\n\t"                                                              \
      "addc.u32 %4, 0, 0;\n\t}"                                                                   \
      : "=&r"(r[0]), "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3]), "=&r"(cout)                            \
      : "r"(x[0]), "r"(x[1]), "r"(x[2]), "r"(x[3]), "r"(y[0]), "r"(y[1]), "r"(y[2]), "r"(y[3]), "r"(cin))
__device__ __forceinline__ void mul_wide_k(uint32_t (&r)[16], const uint32_t (&a)[8], const uint32_t (&b)[8]) {
  const uint32_t zero = 0;
  uint32_t z0[8], z2[8], zm[8], sa[4], sb[4], ca, cb, top, c;
  mul_half(z0, a, b);
  mul_half(z2, a + 4, b + 4);
  ZKB_ADD4(sa, a, (a + 4), zero, ca);
  ZKB_ADD4(sb, b, (b + 4), zero, cb);
  mul_half(zm, sa, sb);
  // + ca sb B + cb sa B + ca cb B^2  (zm grows to 9 limbs: `top`)
  uint32_t t[4], u[4], hi[4];
  const uint32_t ma = 0u - ca, mb = 0u - cb;
#pragma unroll
  for (int i = 0; i < 4; i++) { t[i] = sb[i] & ma; u[i] = sa[i] & mb; }
  ZKB_ADD4(hi, (zm + 4), t, zero, c);
  top = c + (ca & cb);
  ZKB_ADD4((zm + 4), hi, u, zero, c);
  top += c;
  // zm -= z0, zm -= z2 (no borrow out of the 9 limbs)
  uint32_t bw;
  ZKB_SUB8(zm, zm, z0, zero, bw);
  top -= bw;
  ZKB_SUB8(zm, zm, z2, zero, bw);
  top -= bw;
  // r = z0 + zm B + z2 B^2
#pragma unroll
  for (int i = 0; i < 4; i++) r[i] = z0[i];
  uint32_t X[8] = {z0[4], z0[5], z0[6], z0[7], z2[0], z2[1], z2[2], z2[3]};
  ZKB_ADD8((r + 4), X, zm, zero, c);
  uint32_t Y[4] = {top, 0, 0, 0};
  uint32_t c2;
  ZKB_ADD4((r + 12), (z2 + 4), Y, c, c2);
}

// r = a * a: 28 off-diagonal products, doubled, plus the 8 diagonal squares
__device__ __forceinline__ void sqr_wide(uint32_t (&r)[16], const uint32_t (&a)[8]) {
  uint32_t E[17], O[16];
#pragma unroll
  for (int i = 0; i < 17; i++) E[i] = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) O[i] = 0;
  // row i: a_i * a_j, j > i;  i + j odd -> O at 2i, 2i+2, ...;  i + j even -> E at 2i+2, 2i+4, ...
  ZKB_CMAD_TOP((O + 0), O[8], a[1], a[3], a[5], a[7], a[0]);
  ZKB_WCHAIN3((E + 2), E[8], a[2], a[4], a[6], a[0]);
  ZKB_WCHAIN3((O + 2), O[8], a[2], a[4], a[6], a[1]);
  ZKB_WCHAIN3((E + 4), E[10], a[3], a[5], a[7], a[1]);
  ZKB_WCHAIN3((O + 4), O[10], a[3], a[5], a[7], a[2]);
  ZKB_WCHAIN2((E + 6), E[10], a[4], a[6], a[2]);
  ZKB_WCHAIN2((O + 6), O[10], a[4], a[6], a[3]);
  ZKB_WCHAIN2((E + 8), E[12], a[5], a[7], a[3]);
  ZKB_WCHAIN2((O + 8), O[12], a[5], a[7], a[4]);
  ZKB_WCHAIN1((E + 10), E[12], a[6], a[4]);
  ZKB_WCHAIN1((O + 10), O[12], a[6], a[5]);
  ZKB_WCHAIN1((E + 12), E[14], a[7], a[5]);
  ZKB_WCHAIN1((O + 12), O[14], a[7], a[6]);
  uint32_t S[16];
  wide_merge(S, E, O);
  uint32_t D[16];
  D[0] = S[0] << 1;
#pragma unroll
  for (int i = 1; i < 16; i++) D[i] = __funnelshift_l(S[i - 1], S[i], 1);
  // r = D + sum_i a_i^2 2^(64 i): the diagonal products ride the carry chain as multiply-adds
  asm("mad.lo.cc.u32 %0, %16, %16, %0;\n\t"
      "madc.hi.cc.u32 %1, %16, %16, %1;\n\t"
      "madc.lo.cc.u32 %2, %17, %17, %2;\n\t"
      "madc.hi.cc.u32 %3, %17, %17, %3;\n\t"
      "madc.lo.cc.u32 %4, %18, %18, %4;\n\t"
      "madc.hi.cc.u32 %5, %18, %18, %5;\n\t"
      "madc.lo.cc.u32 %6, %19, %19, %6;\n\t"
      "madc.hi.cc.u32 %7, %19, %19, %7;\n\t"
      "madc.lo.cc.u32 %8, %20, %20, %8;\n\t"
      "madc.hi.cc.u32 %9, %20, %20, %9;\n\t"
      "madc.lo.cc.u32 %10, %21, %21, %10;\n\t"
      "madc.hi.cc.u32 %11, %21, %21, %11;\n\t"
      "madc.lo.cc.u32 %12, %22, %22, %12;\n\t"
      "madc.hi.cc.u32 %13, %22, %22, %13;\n\t"
      "madc.lo.cc.u32 %14, %23, %23, %14;\n\t"
      "madc.hi.cc.u32 %15, %23, %23, %15;"
      : "+r"(D[0]), "+r"(D[1]), "+r"(D[2]), "+r"(D[3]), "+r"(D[4]), "+r"(D[5]), "+r"(D[6]), "+r"(D[7]), "+r"(D[8]),
        "+r"(D[9]), "+r"(D[10]), "+r"(D[11]), "+r"(D[12]), "+r"(D[13]), "+r"(D[14]), "+r"(D[15])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]));
#pragma unroll
  for (int i = 0; i < 16; i++) r[i] = D[i];
}

// r = T / 2^256 mod p for T < p * 2^256 (fully reduced result).  State t = E + (O << 32); a round
// clears the low limb (m = E[0] * inv; O += m * p_odd; E += m * p_even), the shift renames O to E
// and (E >> 64) to O, adds the stray limb E[1] at limb 0 and brings T[8 + i] in at limb 7.
template <class P>
__device__ __forceinline__ void redc(uint32_t (&r)[8], const uint32_t (&T)[16]) {
  const uint32_t p0 = P::P0, p1 = P::P1, p2 = P::P2, p3 = P::P3, p4 = P::P4, p5 = P::P5, p6 = P::P6, p7 = P::P7;
  uint32_t E[8], O[8];
#pragma unroll
  for (int k = 0; k < 8; k++) { E[k] = T[k]; O[k] = 0; }
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const uint32_t mi = E[0] * P::INV;
    ZKB_CMAD(O, p1, p3, p5, p7, mi);
    ZKB_CMAD_TOP(E, O[7], p0, p2, p4, p6, mi);
    if (i == 7) break;
    uint32_t nO[8];
    asm("add.cc.u32 %0, %0, %9;\n\t"
        "addc.cc.u32 %1, %10, 0;\n\t"
        "addc.cc.u32 %2, %11, 0;\n\t"
        "addc.cc.u32 %3, %12, 0;\n\t"
        "addc.cc.u32 %4, %13, 0;\n\t"
        "addc.cc.u32 %5, %14, 0;\n\t"
        "addc.cc.u32 %6, %15, 0;\n\t"
        "addc.cc.u32 %7, %16, 0;\n\t"
        "addc.u32 %8, 0, 0;"
        : "+r"(O[0]), "=&r"(nO[0]), "=&r"(nO[1]), "=&r"(nO[2]), "=&r"(nO[3]), "=&r"(nO[4]), "=&r"(nO[5]), "=&r"(nO[6]), "=&r"(nO[7])
        : "r"(E[1]), "r"(E[2]), "r"(E[3]), "r"(E[4]), "r"(E[5]), "r"(E[6]), "r"(E[7]), "r"(T[8 + i]));
#pragma unroll
    for (int k = 0; k < 8; k++) { E[k] = O[k]; O[k] = nO[k]; }
  }
  // (E >> 32) + O + (T[15] << 224)
  asm("add.cc.u32 %0, %8, %15;\n\t"
      "addc.cc.u32 %1, %9, %16;\n\t"
      "addc.cc.u32 %2, %10, %17;\n\t"
      "addc.cc.u32 %3, %11, %18;\n\t"
      "addc.cc.u32 %4, %12, %19;\n\t"
      "addc.cc.u32 %5, %13, %20;\n\t"
      "addc.cc.u32 %6, %14, %21;\n\t"
      "addc.u32 %7, %22, %23;"
      : "=&r"(r[0]), "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3]), "=&r"(r[4]), "=&r"(r[5]), "=&r"(r[6]), "=&r"(r[7])
      : "r"(E[1]), "r"(E[2]), "r"(E[3]), "r"(E[4]), "r"(E[5]), "r"(E[6]), "r"(E[7]),
        "r"(O[0]), "r"(O[1]), "r"(O[2]), "r"(O[3]), "r"(O[4]), "r"(O[5]), "r"(O[6]), "r"(O[7]), "r"(T[15]));
  final_sub<P>(r);
}

template <class P>
__device__ __forceinline__ void psq(uint32_t (&q2)[16]) {
  q2[0] = P::PSQ0; q2[1] = P::PSQ1; q2[2] = P::PSQ2; q2[3] = P::PSQ3; q2[4] = P::PSQ4; q2[5] = P::PSQ5;
  q2[6] = P::PSQ6; q2[7] = P::PSQ7; q2[8] = P::PSQ8; q2[9] = P::PSQ9; q2[10] = P::PSQ10; q2[11] = P::PSQ11;
  q2[12] = P::PSQ12; q2[13] = P::PSQ13; q2[14] = P::PSQ14; q2[15] = P::PSQ15;
}

// a^2 (Montgomery)
template <class P>
__device__ __forceinline__ Fp<P> sqr(const Fp<P>& a) {
  uint32_t T[16];
  sqr_wide(T, a.v);
  Fp<P> r;
  redc<P>(r.v, T);
  return r;
}

// a * b - c * d (Montgomery), one reduction
template <class P>
__device__ __forceinline__ Fp<P> mul_sub_mul(const Fp<P>& a, const Fp<P>& b, const Fp<P>& c, const Fp<P>& d) {
  uint32_t T[16], U[16], Q2[16];
  mul_wide(T, a.v, b.v);
  mul_wide(U, c.v, d.v);
  psq<P>(Q2);
  wide_add(T, T, Q2);  // < 2 p^2 < p * 2^256
  wide_sub(T, T, U);
  Fp<P> r;
  redc<P>(r.v, T);
  return r;
}

template <class P>
__device__ __forceinline__ Fp<P> add(const Fp<P>& a, const Fp<P>& b) {
  uint32_t s[8];
  asm("add.cc.u32 %0, %8, %16;\n\t"
      "addc.cc.u32 %1, %9, %17;\n\t"
      "addc.cc.u32 %2, %10, %18;\n\t"
      "addc.cc.u32 %3, %11, %19;\n\t"
      "addc.cc.u32 %4, %12, %20;\n\t"
      "addc.cc.u32 %5, %13, %21;\n\t"
      "addc.cc.u32 %6, %14, %22;\n\t"
      "addc.u32 %7, %15, %23;"
      : "=&r"(s[0]), "=&r"(s[1]), "=&r"(s[2]), "=&r"(s[3]), "=&r"(s[4]), "=&r"(s[5]), "=&r"(s[6]), "=&r"(s[7])
      : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
        "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
  final_sub<P>(s);
  Fp<P> r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = s[i];
  return r;
}

template <class P>
__device__ __forceinline__ Fp<P> sub(const Fp<P>& a, const Fp<P>& b) {
  uint32_t d[8], br;
  asm("sub.cc.u32 %0, %9, %17;\n\t"
      "subc.cc.u32 %1, %10, %18;\n\t"
      "subc.cc.u32 %2, %11, %19;\n\t"
      "subc.cc.u32 %3, %12, %20;\n\t"
      "subc.cc.u32 %4, %13, %21;\n\t"
      "subc.cc.u32 %5, %14, %22;\n\t"
      "subc.cc.u32 %6, %15, %23;\n\t"
      "subc.cc.u32 %7, %16, %24;\n\t"
      "subc.u32 %8, 0, 0;"
      : "=&r"(d[0]), "=&r"(d[1]), "=&r"(d[2]), "=&r"(d[3]), "=&r"(d[4]), "=&r"(d[5]), "=&r"(d[6]), "=&r"(d[7]), "=&r"(br)
      : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
        "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
  // br = 0xffffffff on borrow: add p back
  asm("add.cc.u32 %0, %0, %8;\n\t"
      "addc.cc.u32 %1, %1, %9;\n\t"
      "addc.cc.u32 %2, %2, %10;\n\t"
      "addc.cc.u32 %3, %3, %11;\n\t"
      "addc.cc.u32 %4, %4, %12;\n\t"
      "addc.cc.u32 %5, %5, %13;\n\t"
      "addc.cc.u32 %6, %6, %14;\n\t"
      "addc.u32 %7, %7, %15;"
      : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3]), "+r"(d[4]), "+r"(d[5]), "+r"(d[6]), "+r"(d[7])
      : "r"(P::P0 & br), "r"(P::P1 & br), "r"(P::P2 & br), "r"(P::P3 & br), "r"(P::P4 & br), "r"(P::P5 & br),
        "r"(P::P6 & br), "r"(P::P7 & br));
  Fp<P> r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = d[i];
  return r;
}
}  // namespace dev_impl
#endif  // __CUDA_ARCH__

// ------------------------------------------------------------------------------------------------
// dispatch
// ------------------------------------------------------------------------------------------------
// A translation unit may define ZKB_FP_OOL (ZKB_FQ2_OOL for the Fq2 layer) before including this
// header: products are then calls to one shared out-of-line copy per operation (operands by value,
// i.e. in registers) instead of being inlined at every use.  A fully inlined mixed addition is
// 35 KB (G1) / 100 KB (G2) of code; whether the instruction cache or the call overhead costs more is
// measured per kernel (profiles/).  ZKB_NO_LAZY selects the plain CIOS product everywhere.
#if defined(__CUDA_ARCH__)
template <class P> static __device__ __noinline__ Fp<P> fp_mul_ool(Fp<P> a, Fp<P> b) { return dev_impl::mul(a, b); }
template <class P> static __device__ __noinline__ Fp<P> fp_sqr_ool(Fp<P> a) { return dev_impl::sqr(a); }
template <class P> static __device__ __noinline__ Fp<P> fp_msm_ool(Fp<P> a, Fp<P> b, Fp<P> c, Fp<P> d) {
  return dev_impl::mul_sub_mul(a, b, c, d);
}
#endif
template <class P> ZKB_HD Fp<P> operator*(const Fp<P>& a, const Fp<P>& b) {
#if defined(__CUDA_ARCH__) && (defined(ZKB_FP_OOL) || defined(ZKB_FP_MUL_OOL))
  return fp_mul_ool<P>(a, b);
#elif defined(__CUDA_ARCH__)
  return dev_impl::mul(a, b);
#else
  return host_impl::mul(a, b);
#endif
}
template <class P> ZKB_HD Fp<P> operator+(const Fp<P>& a, const Fp<P>& b) {
#if defined(__CUDA_ARCH__)
  return dev_impl::add(a, b);
#else
  return host_impl::add(a, b);
#endif
}
template <class P> ZKB_HD Fp<P> operator-(const Fp<P>& a, const Fp<P>& b) {
#if defined(__CUDA_ARCH__)
  return dev_impl::sub(a, b);
#else
  return host_impl::sub(a, b);
#endif
}
template <class P> ZKB_HD Fp<P> neg(const Fp<P>& a) { return Fp<P>::zero() - a; }
template <class P> ZKB_HD Fp<P> sqr(const Fp<P>& a) {
#if defined(__CUDA_ARCH__) && !defined(ZKB_NO_LAZY) && defined(ZKB_FP_OOL)
  return fp_sqr_ool<P>(a);
#elif defined(__CUDA_ARCH__) && !defined(ZKB_NO_LAZY)
  return dev_impl::sqr(a);
#else
  return a * a;
#endif
}
// a * b - c * d
template <class P> ZKB_HD Fp<P> mul_sub_mul(const Fp<P>& a, const Fp<P>& b, const Fp<P>& c, const Fp<P>& d) {
#if defined(__CUDA_ARCH__) && !defined(ZKB_NO_LAZY) && defined(ZKB_FP_OOL)
  return fp_msm_ool<P>(a, b, c, d);
#elif defined(__CUDA_ARCH__) && !defined(ZKB_NO_LAZY)
  return dev_impl::mul_sub_mul(a, b, c, d);
#else
  return a * b - c * d;
#endif
}
template <class P> ZKB_HD Fp<P> dbl(const Fp<P>& a) { return a + a; }

// canonical residue (4x u64 LE at the C ABI == 8x u32 LE) <-> Montgomery form
template <class P> ZKB_HD Fp<P> to_mont(const Fp<P>& canon) { return canon * Fp<P>::r2(); }
template <class P> ZKB_HD Fp<P> from_mont(const Fp<P>& m) {
  Fp<P> one_raw = Fp<P>::zero();
  one_raw.v[0] = 1;
  return m * one_raw;
}

// a^e for a 256-bit exponent given as 8 u32 limbs (plain integer, not Montgomery)
template <class P> ZKB_HD Fp<P> pow_limbs(const Fp<P>& a, const uint32_t e[8]) {
  Fp<P> acc = Fp<P>::one();
  bool started = false;
  for (int i = 7; i >= 0; i--) {
    for (int b = 31; b >= 0; b--) {
      if (started) acc = sqr(acc);
      if ((e[i] >> b) & 1u) { acc = started ? acc * a : a; started = true; }
    }
  }
  return acc;
}
template <class P> ZKB_HD Fp<P> pow_u64(const Fp<P>& a, uint64_t e) {
  uint32_t l[8] = {(uint32_t)e, (uint32_t)(e >> 32), 0, 0, 0, 0, 0, 0};
  return pow_limbs(a, l);
}
// Fermat inverse a^(p-2) (254 squarings + ~127 products); inverse(0) = 0.  Kept as the definition the fast
// inverse below is checked against.
template <class P> ZKB_HD Fp<P> inverse_fermat(const Fp<P>& a) {
  uint32_t e[8] = {P::P0 - 2u, P::P1, P::P2, P::P3, P::P4, P::P5, P::P6, P::P7};  // P0 >= 2 for both primes
  return pow_limbs(a, e);
}

// ------------------------------------------------------------------------------------------------
// Fast modular inverse: Bernstein-Yang "safegcd" division steps on signed 30-bit limbs (the layout
// libsecp256k1's modinv32 made standard): 20 rounds of 30 branch-free division steps on the low limbs
// plus two 2x2-matrix updates of the 9-limb values (90 32x32->64 multiplies per round).  About 1800
// wide multiplies and ~15 k simple integer instructions instead of the ~50 k wide multiplies of the
// Fermat exponentiation: an inversion costs ~15 Montgomery products of the multiplier pipe instead of
// ~380.  No data-dependent branches: the 32 lanes of a warp stay converged.  inverse(0) = 0.
// ------------------------------------------------------------------------------------------------
namespace modinv {
struct S30 { int32_t v[9]; };
struct T2 { int32_t u, v, q, r; };
static constexpr int32_t M30 = (int32_t)(0xffffffffu >> 2);

// 30 division steps on the low 30 bits of (f, g); t = 2^30 times the transition matrix; zeta = -(delta + 1/2)
ZKB_HD int32_t divsteps_30(int32_t zeta, uint32_t f0, uint32_t g0, T2& t) {
  uint32_t u = 1, v = 0, q = 0, r = 1, f = f0, g = g0;
  for (int i = 0; i < 30; i++) {
    uint32_t mask1 = (uint32_t)(zeta >> 31);  // zeta < 0
    uint32_t mask2 = (uint32_t)0 - (g & 1u);  // g odd
    uint32_t x = (f ^ mask1) - mask1, y = (u ^ mask1) - mask1, z = (v ^ mask1) - mask1;
    g += x & mask2; q += y & mask2; r += z & mask2;
    mask1 &= mask2;
    zeta = (int32_t)((uint32_t)zeta ^ mask1) - 1;
    f += g & mask1; u += q & mask1; v += r & mask1;
    g >>= 1; u <<= 1; v <<= 1;
  }
  t.u = (int32_t)u; t.v = (int32_t)v; t.q = (int32_t)q; t.r = (int32_t)r;
  return zeta;
}
// (d, e) <- t (d, e) / 2^30 mod p, kept in (-2p, p)
ZKB_HD void update_de(S30& d, S30& e, const T2& t, const S30& mod, uint32_t mod_inv30) {
  const int32_t u = t.u, v = t.v, q = t.q, r = t.r;
  const int32_t sd = d.v[8] >> 31, se = e.v[8] >> 31;
  int32_t md = (u & sd) + (v & se), me = (q & sd) + (r & se);
  int32_t di = d.v[0], ei = e.v[0];
  int64_t cd = (int64_t)u * di + (int64_t)v * ei;
  int64_t ce = (int64_t)q * di + (int64_t)r * ei;
  md -= (int32_t)((mod_inv30 * (uint32_t)cd + (uint32_t)md) & (uint32_t)M30);
  me -= (int32_t)((mod_inv30 * (uint32_t)ce + (uint32_t)me) & (uint32_t)M30);
  cd += (int64_t)mod.v[0] * md;
  ce += (int64_t)mod.v[0] * me;
  cd >>= 30; ce >>= 30;
  for (int i = 1; i < 9; i++) {
    di = d.v[i]; ei = e.v[i];
    cd += (int64_t)u * di + (int64_t)v * ei;
    ce += (int64_t)q * di + (int64_t)r * ei;
    cd += (int64_t)mod.v[i] * md;
    ce += (int64_t)mod.v[i] * me;
    d.v[i - 1] = (int32_t)cd & M30; cd >>= 30;
    e.v[i - 1] = (int32_t)ce & M30; ce >>= 30;
  }
  d.v[8] = (int32_t)cd;
  e.v[8] = (int32_t)ce;
}
// (f, g) <- t (f, g) / 2^30 (exact)
ZKB_HD void update_fg(S30& f, S30& g, const T2& t) {
  const int32_t u = t.u, v = t.v, q = t.q, r = t.r;
  int32_t fi = f.v[0], gi = g.v[0];
  int64_t cf = (int64_t)u * fi + (int64_t)v * gi;
  int64_t cg = (int64_t)q * fi + (int64_t)r * gi;
  cf >>= 30; cg >>= 30;
  for (int i = 1; i < 9; i++) {
    fi = f.v[i]; gi = g.v[i];
    cf += (int64_t)u * fi + (int64_t)v * gi;
    cg += (int64_t)q * fi + (int64_t)r * gi;
    f.v[i - 1] = (int32_t)cf & M30; cf >>= 30;
    g.v[i - 1] = (int32_t)cg & M30; cg >>= 30;
  }
  f.v[8] = (int32_t)cf;
  g.v[8] = (int32_t)cg;
}
// r in (-2p, p), negated if sign < 0, brought to [0, p) with limbs in [0, 2^30)
ZKB_HD void normalize(S30& r, int32_t sign, const S30& mod) {
  int32_t cond_add = r.v[8] >> 31;
  for (int i = 0; i < 9; i++) r.v[i] += mod.v[i] & cond_add;
  const int32_t cond_negate = sign >> 31;
  for (int i = 0; i < 9; i++) r.v[i] = (r.v[i] ^ cond_negate) - cond_negate;
  for (int i = 0; i < 8; i++) { r.v[i + 1] += r.v[i] >> 30; r.v[i] &= M30; }
  cond_add = r.v[8] >> 31;
  for (int i = 0; i < 9; i++) r.v[i] += mod.v[i] & cond_add;
  for (int i = 0; i < 8; i++) { r.v[i + 1] += r.v[i] >> 30; r.v[i] &= M30; }
}
ZKB_HD void to_s30(S30& o, const uint32_t x[8]) {
  for (int i = 0; i < 9; i++) {
    const int bit = 30 * i, w = bit >> 5, sh = bit & 31;
    uint64_t two = (uint64_t)x[w] | (w + 1 < 8 ? (uint64_t)x[w + 1] << 32 : 0);
    o.v[i] = (int32_t)((uint32_t)(two >> sh) & (uint32_t)M30);
  }
}
ZKB_HD void from_s30(uint32_t x[8], const S30& a) {
  for (int w = 0; w < 8; w++) {
    const int bit = 32 * w, i = bit / 30, sh = bit - 30 * i;  // word w = bits [32w, 32w+32): limbs i (from bit sh) and i+1
    uint64_t two = (uint64_t)(uint32_t)a.v[i] | (i + 1 < 9 ? (uint64_t)(uint32_t)a.v[i + 1] << 30 : 0);
    x[w] = (uint32_t)(two >> sh);
  }
}
}  // namespace modinv

// x^-1 mod p of the stored 256-bit residue (no Montgomery factor applied or removed); 0 -> 0
template <class P> ZKB_HD Fp<P> inverse_plain(const Fp<P>& x) {
  using namespace modinv;
  const uint32_t pm[8] = {P::P0, P::P1, P::P2, P::P3, P::P4, P::P5, P::P6, P::P7};
  S30 mod, d, e, f, g;
  to_s30(mod, pm);
  to_s30(g, x.v);
  for (int i = 0; i < 9; i++) { d.v[i] = 0; e.v[i] = 0; }
  e.v[0] = 1;
  f = mod;
  const uint32_t mod_inv30 = ((uint32_t)0 - P::INV) & (uint32_t)M30;  // p^-1 mod 2^30 from -p^-1 mod 2^32
  int32_t zeta = -1;
  for (int it = 0; it < 20; it++) {  // 600 >= 590 division steps: enough for any 256-bit input
    T2 t;
    zeta = divsteps_30(zeta, (uint32_t)f.v[0], (uint32_t)g.v[0], t);
    update_de(d, e, t, mod, mod_inv30);
    update_fg(f, g, t);
  }
  normalize(d, f.v[8], mod);
  Fp<P> r;
  from_s30(r.v, d);
  return r;
}
// Montgomery-form inverse: stored aR -> stored a^-1 R.  (aR)^-1 = a^-1 R^-1; times R^3 under the
// Montgomery product gives a^-1 R.  inverse(0) = 0 (callers check for zero where the reference would panic)
template <class P> ZKB_HD Fp<P> inverse(const Fp<P>& a) {
  Fp<P> r3;
  r3.v[0] = P::R30; r3.v[1] = P::R31; r3.v[2] = P::R32; r3.v[3] = P::R33;
  r3.v[4] = P::R34; r3.v[5] = P::R35; r3.v[6] = P::R36; r3.v[7] = P::R37;
  return inverse_plain(a) * r3;
}

typedef Fp<FrParams> Fr;
typedef Fp<FqParams> Fq;

// ------------------------------------------------------------------------------------------------
// Fq2 = Fq[u]/(u^2 + 1)   (crate bn: Fq2, non-residue -1)
// ------------------------------------------------------------------------------------------------
struct alignas(16) Fq2 {
  Fq c0, c1;
  ZKB_HD static Fq2 zero() { Fq2 r; r.c0 = Fq::zero(); r.c1 = Fq::zero(); return r; }
  ZKB_HD static Fq2 one() { Fq2 r; r.c0 = Fq::one(); r.c1 = Fq::zero(); return r; }
  ZKB_HD bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
  ZKB_HD bool operator==(const Fq2& b) const { return c0 == b.c0 && c1 == b.c1; }
  ZKB_HD bool operator!=(const Fq2& b) const { return !(*this == b); }
};
ZKB_HD Fq2 operator+(const Fq2& a, const Fq2& b) { Fq2 r; r.c0 = a.c0 + b.c0; r.c1 = a.c1 + b.c1; return r; }
ZKB_HD Fq2 operator-(const Fq2& a, const Fq2& b) { Fq2 r; r.c0 = a.c0 - b.c0; r.c1 = a.c1 - b.c1; return r; }
#if defined(__CUDA_ARCH__)
namespace dev_impl {
// Karatsuba on unreduced 512-bit products: 3 wide products, 2 Montgomery reductions
#if defined(ZKB_KARATSUBA)
#define ZKB_MULW mul_wide_k
#else
#define ZKB_MULW mul_wide
#endif
__device__ __forceinline__ Fq2 fq2_mul_lazy(const Fq2& a, const Fq2& b) {
  uint32_t v0[16], v1[16], v2[16], sa[8], sb[8], q2[16];
  ZKB_MULW(v0, a.c0.v, b.c0.v);
  ZKB_MULW(v1, a.c1.v, b.c1.v);
  add_nored(sa, a.c0.v, a.c1.v);
  add_nored(sb, b.c0.v, b.c1.v);
  ZKB_MULW(v2, sa, sb);
  wide_sub(v2, v2, v0);
  wide_sub(v2, v2, v1);  // a0 b1 + a1 b0 < 2 q^2
  psq<FqParams>(q2);
  wide_add(v0, v0, q2);
  wide_sub(v0, v0, v1);  // a0 b0 - a1 b1 + q^2 in (0, 2 q^2)
  Fq2 r;
  redc<FqParams>(r.c0.v, v0);
  redc<FqParams>(r.c1.v, v2);
  return r;
}
__device__ __forceinline__ Fq2 fq2_mul_plain(const Fq2& a, const Fq2& b) {
  Fq v0 = mul(a.c0, b.c0), v1 = mul(a.c1, b.c1);
  Fq s = mul(add(a.c0, a.c1), add(b.c0, b.c1));
  Fq2 r;
  r.c0 = sub(v0, v1);
  r.c1 = sub(sub(s, v0), v1);
  return r;
}
__device__ __forceinline__ Fq2 fq2_sqr(const Fq2& a) {
  Fq t = mul(a.c0, a.c1);
  Fq2 r;
  r.c0 = mul(add(a.c0, a.c1), sub(a.c0, a.c1));
  r.c1 = add(t, t);
  return r;
}
// a * b - c * d over Fq2 with TWO reductions (two Karatsuba products unreduced, combined in 512 bits): the last step
// of the mixed addition, rr (q - x3) - y1 ppp.  Ranges (q < 2^254, so 4 q^2 < q 2^256, the bound of redc):
//   re = a0 b0 - a1 b1 - c0 d0 + c1 d1 + 2 q^2   in (0, 4 q^2)
//   im = (a0 b1 + a1 b0) - (c0 d1 + c1 d0) + 2 q^2 in (0, 4 q^2); formed as [a0 b1 + a1 b0 + c0 d0 + c1 d1 + 2 q^2]
//        (< 6 q^2 < 2^512) minus (c0 + c1)(d0 + d1), so no intermediate goes negative.
__device__ __forceinline__ Fq2 fq2_msm_lazy(const Fq2& a, const Fq2& b, const Fq2& c, const Fq2& d) {
  uint32_t re[16], im[16], t[16], s1[8], s2[8], q2[16];
  psq<FqParams>(q2);
  mul_wide(re, a.c0.v, b.c0.v);
  mul_wide(t, a.c1.v, b.c1.v);
  add_nored(s1, a.c0.v, a.c1.v);
  add_nored(s2, b.c0.v, b.c1.v);
  mul_wide(im, s1, s2);
  wide_sub(im, im, re);
  wide_sub(im, im, t);   // a0 b1 + a1 b0
  wide_add(re, re, q2);
  wide_sub(re, re, t);   // a0 b0 - a1 b1 + q^2
  mul_wide(t, c.c0.v, d.c0.v);
  wide_add(re, re, q2);
  wide_sub(re, re, t);   // ... - c0 d0 + q^2
  wide_add(im, im, t);
  mul_wide(t, c.c1.v, d.c1.v);
  wide_add(re, re, t);   // ... + c1 d1
  wide_add(im, im, t);
  add_nored(s1, c.c0.v, c.c1.v);
  add_nored(s2, d.c0.v, d.c1.v);
  mul_wide(t, s1, s2);
  wide_add(im, im, q2);
  wide_add(im, im, q2);
  wide_sub(im, im, t);
  Fq2 r;
  redc<FqParams>(r.c0.v, re);
  redc<FqParams>(r.c1.v, im);
  return r;
}
}  // namespace dev_impl
#if defined(ZKB_NO_LAZY)
#define ZKB_FQ2_MUL_IMPL dev_impl::fq2_mul_plain
#else
#define ZKB_FQ2_MUL_IMPL dev_impl::fq2_mul_lazy
#endif
static __device__ __noinline__ Fq2 fq2_mul_ool(Fq2 a, Fq2 b) { return ZKB_FQ2_MUL_IMPL(a, b); }
static __device__ __noinline__ Fq2 fq2_sqr_ool(Fq2 a) { return dev_impl::fq2_sqr(a); }
static __device__ __noinline__ Fq2 fq2_msm_ool(Fq2 a, Fq2 b, Fq2 c, Fq2 d) { return dev_impl::fq2_msm_lazy(a, b, c, d); }
#endif
ZKB_HD Fq2 operator*(const Fq2& a, const Fq2& b) {
#if defined(__CUDA_ARCH__) && defined(ZKB_FQ2_OOL)
  return fq2_mul_ool(a, b);
#elif defined(__CUDA_ARCH__)
  return ZKB_FQ2_MUL_IMPL(a, b);
#else
  // Karatsuba: 3 base multiplications
  Fq v0 = a.c0 * b.c0, v1 = a.c1 * b.c1;
  Fq s = (a.c0 + a.c1) * (b.c0 + b.c1);
  Fq2 r;
  r.c0 = v0 - v1;
  r.c1 = s - v0 - v1;
  return r;
#endif
}
// a * b - c * d over Fq2
// a * b as a wide product followed by a stand-alone reduction; KARA: the wide product by one level of Karatsuba
// (micro-benchmark and validation hooks; host: the plain product)
template <bool KARA, class P>
ZKB_HD Fp<P> mul_wide_redc(const Fp<P>& a, const Fp<P>& b) {
#if defined(__CUDA_ARCH__)
  uint32_t T[16];
  if (KARA) dev_impl::mul_wide_k(T, a.v, b.v); else dev_impl::mul_wide(T, a.v, b.v);
  Fp<P> r;
  dev_impl::redc<P>(r.v, T);
  return r;
#else
  return a * b;
#endif
}
// the two-reduction form by name (validation hook zkb_field_op; host: the plain definition)
ZKB_HD Fq2 mul_sub_mul_lazy(const Fq2& a, const Fq2& b, const Fq2& c, const Fq2& d) {
#if defined(__CUDA_ARCH__)
  return dev_impl::fq2_msm_lazy(a, b, c, d);
#else
  return a * b - c * d;
#endif
}
// ZKB_FQ2_MSM_LAZY: the two-reduction form (device, lazy build); 2 = out of line (the G2 accumulation kernel)
ZKB_HD Fq2 mul_sub_mul(const Fq2& a, const Fq2& b, const Fq2& c, const Fq2& d) {
#if defined(__CUDA_ARCH__) && !defined(ZKB_NO_LAZY) && defined(ZKB_FQ2_MSM_LAZY) && ZKB_FQ2_MSM_LAZY == 2
  return fq2_msm_ool(a, b, c, d);
#elif defined(__CUDA_ARCH__) && !defined(ZKB_NO_LAZY) && defined(ZKB_FQ2_MSM_LAZY)
  return dev_impl::fq2_msm_lazy(a, b, c, d);
#else
  return a * b - c * d;
#endif
}
ZKB_HD Fq2 sqr(const Fq2& a) {
#if defined(__CUDA_ARCH__) && defined(ZKB_FQ2_OOL)
  return fq2_sqr_ool(a);
#elif defined(__CUDA_ARCH__)
  return dev_impl::fq2_sqr(a);
#else
  Fq t = a.c0 * a.c1;
  Fq2 r;
  r.c0 = (a.c0 + a.c1) * (a.c0 - a.c1);
  r.c1 = t + t;
  return r;
#endif
}
ZKB_HD Fq2 neg(const Fq2& a) { Fq2 r; r.c0 = neg(a.c0); r.c1 = neg(a.c1); return r; }
ZKB_HD Fq2 dbl(const Fq2& a) { return a + a; }
ZKB_HD Fq2 inverse(const Fq2& a) {
  Fq n = inverse(sqr(a.c0) + sqr(a.c1));
  Fq2 r;
  r.c0 = a.c0 * n;
  r.c1 = neg(a.c1 * n);
  return r;
}

}  // namespace zkb
