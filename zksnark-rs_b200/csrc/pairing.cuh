// Optimal-ate pairing on BN254 for groth16::verify (/root/reference/src/groth16/mod.rs:299-320):
// replaces crate `bn`'s `pairing(G1, G2) -> Gt` (fr.rs:120-122) and the Gt product the reference
// writes as "+" (fr.rs:225-231).
//
// Tower: Fq2 = Fq[u]/(u^2+1); xi = 9 + u; Fq6 = Fq2[v]/(v^3 - xi); Fq12 = Fq6[w]/(w^2 - v), i.e.
// Fq12 = Fq2[w]/(w^6 - xi) with the w^i coefficient a_i stored at c[i & 1].c[i >> 1].  G2 lives on the
// D-type twist y^2 = x^3 + 3/xi; the untwist is (x, y) -> (x w^2, y w^3).
//
// The value is f_{6u+2,Q}(P) l_{[6u+2]Q,pi(Q)}(P) l_{..,-pi^2(Q)}(P) raised to (q^12 - 1)/r.  Lines
// are the affine chord/tangent l(P) = -yP + (lambda xP) w + (y1 - lambda x1) w^3 (lambda in Fq2), so
// the Miller value equals the textbook one up to factors in proper subfields and the GT element is
// THE reduced pairing: it can be compared coefficient by coefficient with any other correct
// implementation (tests compare with the oracle's dense-polynomial Fq12).
//
// Everything here is __host__ __device__: tests/hostcheck compiles it with g++ and checks it against
// the oracle on a machine without a GPU; the product only ever runs the device instantiation.
#pragma once
#include "ec.cuh"

// the tower products are shared out-of-line bodies on the device (a fully inlined Fq12 product is
// 18 Fq2 products plus ~60 additions; the Miller loop and the final exponentiation use dozens)
#if defined(__CUDACC__)
#define ZKB_OOL static __host__ __device__ __noinline__
#else
#define ZKB_OOL static inline
#endif

namespace zkb {

ZKB_HD Fq2 mul_xi(const Fq2& a) {  // (a0 + a1 u)(9 + u) = (9 a0 - a1) + (9 a1 + a0) u
  Fq t0 = dbl(dbl(dbl(a.c0))) + a.c0, t1 = dbl(dbl(dbl(a.c1))) + a.c1;
  Fq2 r; r.c0 = t0 - a.c1; r.c1 = t1 + a.c0;
  return r;
}
ZKB_HD Fq2 conj(const Fq2& a) { Fq2 r; r.c0 = a.c0; r.c1 = neg(a.c1); return r; }
ZKB_HD Fq2 mul_fq(const Fq2& a, const Fq& k) { Fq2 r; r.c0 = a.c0 * k; r.c1 = a.c1 * k; return r; }

struct Fq6 {
  Fq2 c[3];
  ZKB_HD static Fq6 zero() { Fq6 r; r.c[0] = r.c[1] = r.c[2] = Fq2::zero(); return r; }
  ZKB_HD static Fq6 one() { Fq6 r = zero(); r.c[0] = Fq2::one(); return r; }
};
ZKB_HD Fq6 operator+(const Fq6& a, const Fq6& b) { Fq6 r; for (int i = 0; i < 3; i++) r.c[i] = a.c[i] + b.c[i]; return r; }
ZKB_HD Fq6 operator-(const Fq6& a, const Fq6& b) { Fq6 r; for (int i = 0; i < 3; i++) r.c[i] = a.c[i] - b.c[i]; return r; }
ZKB_HD Fq6 neg(const Fq6& a) { Fq6 r; for (int i = 0; i < 3; i++) r.c[i] = neg(a.c[i]); return r; }
ZKB_HD Fq6 mul_v(const Fq6& a) { Fq6 r; r.c[0] = mul_xi(a.c[2]); r.c[1] = a.c[0]; r.c[2] = a.c[1]; return r; }
// Karatsuba over Fq2: 6 products
ZKB_OOL Fq6 operator*(const Fq6& a, const Fq6& b) {
  Fq2 v0 = a.c[0] * b.c[0], v1 = a.c[1] * b.c[1], v2 = a.c[2] * b.c[2];
  Fq6 r;
  r.c[0] = v0 + mul_xi((a.c[1] + a.c[2]) * (b.c[1] + b.c[2]) - v1 - v2);
  r.c[1] = (a.c[0] + a.c[1]) * (b.c[0] + b.c[1]) - v0 - v1 + mul_xi(v2);
  r.c[2] = (a.c[0] + a.c[2]) * (b.c[0] + b.c[2]) - v0 - v2 + v1;
  return r;
}
ZKB_OOL Fq6 inverse(const Fq6& a) {
  Fq2 t0 = sqr(a.c[0]) - mul_xi(a.c[1] * a.c[2]);
  Fq2 t1 = mul_xi(sqr(a.c[2])) - a.c[0] * a.c[1];
  Fq2 t2 = sqr(a.c[1]) - a.c[0] * a.c[2];
  Fq2 d = inverse(a.c[0] * t0 + mul_xi(a.c[2] * t1 + a.c[1] * t2));
  Fq6 r; r.c[0] = t0 * d; r.c[1] = t1 * d; r.c[2] = t2 * d;
  return r;
}

struct Fq12 {
  Fq6 c[2];
  ZKB_HD static Fq12 one() { Fq12 r; r.c[0] = Fq6::one(); r.c[1] = Fq6::zero(); return r; }
  // coefficient of w^i, i < 6
  ZKB_HD Fq2& w(int i) { return c[i & 1].c[i >> 1]; }
  ZKB_HD const Fq2& w(int i) const { return c[i & 1].c[i >> 1]; }
  ZKB_HD bool operator==(const Fq12& b) const {
    bool e = true;
    for (int i = 0; i < 6; i++) e = e && (w(i) == b.w(i));
    return e;
  }
};
ZKB_OOL Fq12 operator*(const Fq12& a, const Fq12& b) {
  Fq6 v0 = a.c[0] * b.c[0], v1 = a.c[1] * b.c[1];
  Fq12 r;
  r.c[1] = (a.c[0] + a.c[1]) * (b.c[0] + b.c[1]) - v0 - v1;
  r.c[0] = v0 + mul_v(v1);
  return r;
}
ZKB_OOL Fq12 sqr(const Fq12& a) {  // complex squaring: 2 Fq6 products
  Fq6 t = a.c[0] * a.c[1];
  Fq12 r;
  r.c[0] = (a.c[0] + a.c[1]) * (a.c[0] + mul_v(a.c[1])) - t - mul_v(t);
  r.c[1] = t + t;
  return r;
}
ZKB_HD Fq12 conj(const Fq12& a) { Fq12 r; r.c[0] = a.c[0]; r.c[1] = neg(a.c[1]); return r; }  // a^(q^6)
ZKB_OOL Fq12 inverse(const Fq12& a) {
  Fq6 d = inverse(a.c[0] * a.c[0] - mul_v(a.c[1] * a.c[1]));
  Fq12 r; r.c[0] = a.c[0] * d; r.c[1] = neg(a.c[1] * d);
  return r;
}
// a^(q^2): the Fq2 coefficients are fixed, w^i picks up xi^(i (q^2 - 1)/6) in Fq
ZKB_OOL Fq12 frobenius2(const Fq12& a) {
  const uint32_t g[6][8] = ZKB_FROB2_W;
  Fq12 r;
  r.w(0) = a.w(0);
  for (int i = 1; i < 6; i++) {
    Fq k;
    for (int j = 0; j < 8; j++) k.v[j] = g[i][j];
    r.w(i) = mul_fq(a.w(i), k);
  }
  return r;
}

// l(P) = -yP + (lambda xP) w + (y1 - lambda x1) w^3 as a full Fq12 element
ZKB_OOL Fq12 line_value(const Fq2& lambda, const Fq2& x1, const Fq2& y1, const Fq& xP, const Fq& yP) {
  Fq12 l; l.c[0] = Fq6::zero(); l.c[1] = Fq6::zero();
  l.w(0).c0 = neg(yP);
  l.w(1) = mul_fq(lambda, xP);
  l.w(3) = y1 - lambda * x1;
  return l;
}

ZKB_HD Fq2 fq2_const(const uint32_t a[8], const uint32_t b[8]) {
  Fq2 r;
  for (int j = 0; j < 8; j++) { r.c0.v[j] = a[j]; r.c1.v[j] = b[j]; }
  return r;
}

// One step R <- R + S on the twist (S == R: tangent), multiplying the line through them, evaluated at
// P, into f.  R and S are finite and R != -S on every step of the loop for points of order r.
ZKB_OOL void miller_step(Fq12& f, Affine<Fq2>& R, const Affine<Fq2>& S, bool tangent, const Fq& xP, const Fq& yP) {
  Fq2 lambda;
  if (tangent) { Fq2 xx = sqr(R.x); lambda = (dbl(xx) + xx) * inverse(dbl(R.y)); }
  else lambda = (S.y - R.y) * inverse(S.x - R.x);
  f = f * line_value(lambda, R.x, R.y, xP, yP);
  Fq2 x3 = sqr(lambda) - R.x - S.x;
  R.y = lambda * (R.x - x3) - R.y;
  R.x = x3;
}

// Miller loop of the optimal-ate pairing; identity in either slot gives 1 (as `bn` does)
ZKB_HD Fq12 miller_loop(const Affine<Fq>& P, const Affine<Fq2>& Q) {
  Fq12 f = Fq12::one();
  if (P.is_inf() || Q.is_inf()) return f;
  const uint32_t loop[3] = ZKB_ATE_LOOP;
  Affine<Fq2> R = Q;
  for (int i = ZKB_ATE_LOOP_BITS - 2; i >= 0; i--) {
    f = sqr(f);
    miller_step(f, R, R, true, P.x, P.y);
    if ((loop[i >> 5] >> (i & 31)) & 1u) miller_step(f, R, Q, false, P.x, P.y);
  }
  // Q1 = pi(Q), nQ2 = -pi^2(Q) in twist coordinates
  const uint32_t g2a[8] = ZKB_FROB_G2_C0, g2b[8] = ZKB_FROB_G2_C1, g3a[8] = ZKB_FROB_G3_C0, g3b[8] = ZKB_FROB_G3_C1;
  const uint32_t g2s[8] = ZKB_FROB_G2SQ;
  Affine<Fq2> Q1, nQ2;
  Q1.x = conj(Q.x) * fq2_const(g2a, g2b);
  Q1.y = conj(Q.y) * fq2_const(g3a, g3b);
  Fq k; for (int j = 0; j < 8; j++) k.v[j] = g2s[j];
  nQ2.x = mul_fq(Q.x, k);
  nQ2.y = Q.y;
  miller_step(f, R, Q1, false, P.x, P.y);
  // the last line only contributes its value (the point sum is not needed)
  Fq2 lambda = (nQ2.y - R.y) * inverse(nQ2.x - R.x);
  f = f * line_value(lambda, R.x, R.y, P.x, P.y);
  return f;
}

// f^((q^12 - 1)/r) = ((f^(q^6 - 1))^(q^2 + 1))^((q^4 - q^2 + 1)/r)
ZKB_HD Fq12 final_exponentiation(const Fq12& f) {
  Fq12 t = conj(f) * inverse(f);
  t = frobenius2(t) * t;
  const uint32_t e[ZKB_FINAL_EXP_WORDS] = ZKB_FINAL_EXP;
  Fq12 acc = t;  // the top bit of the exponent
  for (int i = ZKB_FINAL_EXP_BITS - 2; i >= 0; i--) {
    acc = sqr(acc);
    if ((e[i >> 5] >> (i & 31)) & 1u) acc = acc * t;
  }
  return acc;
}

}  // namespace zkb
