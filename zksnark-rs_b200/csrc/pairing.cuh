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

// a^q: conjugate every Fq2 coefficient, w^i picks up xi^(i (q - 1)/6) in Fq2
ZKB_OOL Fq12 frobenius(const Fq12& a) {
  const uint32_t g[12][8] = ZKB_FROB1_W;
  Fq12 r;
  r.w(0) = conj(a.w(0));
  for (int i = 1; i < 6; i++) r.w(i) = conj(a.w(i)) * fq2_const(g[2 * i], g[2 * i + 1]);
  return r;
}

// f * (l0 + l1 w + l3 w^3): the sparse product with a line (13 Fq2 products instead of 18).
// With f = A + B w (A, B in Fq6) and the line = a + b w, a = (l0, 0, 0), b = (l1, l3, 0):
//   f * line = (A a + v B b) + ((A + B)(a + b) - A a - B b) w
ZKB_HD Fq6 mul_by_01(const Fq6& x, const Fq2& b0, const Fq2& b1) {  // x * (b0 + b1 v)
  Fq2 v0 = x.c[0] * b0, v1 = x.c[1] * b1;
  Fq6 r;
  r.c[0] = v0 + mul_xi((x.c[1] + x.c[2]) * b1 - v1);
  r.c[1] = (x.c[0] + x.c[1]) * (b0 + b1) - v0 - v1;
  r.c[2] = (x.c[0] + x.c[2]) * b0 - v0 + v1;
  return r;
}
ZKB_OOL void mul_by_line(Fq12& f, const Fq2& l0, const Fq2& l1, const Fq2& l3) {
  Fq6 aa; aa.c[0] = f.c[0].c[0] * l0; aa.c[1] = f.c[0].c[1] * l0; aa.c[2] = f.c[0].c[2] * l0;
  Fq6 bb = mul_by_01(f.c[1], l1, l3);
  Fq6 e = mul_by_01(f.c[0] + f.c[1], l0 + l1, l3);
  f.c[1] = e - aa - bb;
  f.c[0] = aa + mul_v(bb);
}

// Homogeneous projective point on the twist and the two Miller steps (Costello-Lange-Naehrig
// formulas for a = 0, D-type twist).  Each returns the line through the points evaluated at P up to
// a factor in Fq2 (killed by the final exponentiation): l0 = c0 yP, l1 = c1 xP, l3 = c2.
struct G2Proj { Fq2 x, y, z; };
ZKB_OOL void miller_double(Fq12& f, G2Proj& r, const Fq& xP, const Fq& yP) {
  const uint32_t h2[8] = ZKB_FQ_INV2, tb0[8] = ZKB_G2_B0, tb1[8] = ZKB_G2_B1;
  Fq half; for (int j = 0; j < 8; j++) half.v[j] = h2[j];
  Fq2 a = mul_fq(r.x * r.y, half);
  Fq2 b = sqr(r.y), c = sqr(r.z);
  Fq2 e = fq2_const(tb0, tb1) * (dbl(c) + c);
  Fq2 f3 = dbl(e) + e;
  Fq2 g = mul_fq(b + f3, half);
  Fq2 h = sqr(r.y + r.z) - (b + c);
  Fq2 i = e - b;
  Fq2 j = sqr(r.x);
  Fq2 ee = sqr(e);
  r.x = a * (b - f3);
  r.y = sqr(g) - (dbl(ee) + ee);
  r.z = b * h;
  mul_by_line(f, mul_fq(neg(h), yP), mul_fq(dbl(j) + j, xP), i);
}
ZKB_OOL void miller_add(Fq12& f, G2Proj& r, const Affine<Fq2>& q, const Fq& xP, const Fq& yP) {
  Fq2 theta = r.y - q.y * r.z;
  Fq2 lambda = r.x - q.x * r.z;
  Fq2 c = sqr(theta), d = sqr(lambda);
  Fq2 e = lambda * d;
  Fq2 ff = r.z * c;
  Fq2 g = r.x * d;
  Fq2 h = e + ff - dbl(g);
  r.x = lambda * h;
  r.y = theta * (g - h) - e * r.y;
  r.z = r.z * e;
  Fq2 j = theta * q.x - lambda * q.y;
  mul_by_line(f, mul_fq(lambda, yP), mul_fq(neg(theta), xP), j);
}

// Miller loop of the optimal-ate pairing; identity in either slot gives 1 (as `bn` does)
ZKB_HD Fq12 miller_loop(const Affine<Fq>& P, const Affine<Fq2>& Q) {
  Fq12 f = Fq12::one();
  if (P.is_inf() || Q.is_inf()) return f;
  const uint32_t loop[3] = ZKB_ATE_LOOP;
  G2Proj R; R.x = Q.x; R.y = Q.y; R.z = Fq2::one();
  for (int i = ZKB_ATE_LOOP_BITS - 2; i >= 0; i--) {
    f = sqr(f);
    miller_double(f, R, P.x, P.y);
    if ((loop[i >> 5] >> (i & 31)) & 1u) miller_add(f, R, Q, P.x, P.y);
  }
  // Q1 = pi(Q), nQ2 = -pi^2(Q) in twist coordinates
  const uint32_t g1[12][8] = ZKB_FROB1_W;
  const uint32_t g2[6][8] = ZKB_FROB2_W;
  Affine<Fq2> Q1, nQ2;
  Q1.x = conj(Q.x) * fq2_const(g1[4], g1[5]);   // xi^((q-1)/3)
  Q1.y = conj(Q.y) * fq2_const(g1[6], g1[7]);   // xi^((q-1)/2)
  Fq k; for (int j = 0; j < 8; j++) k.v[j] = g2[2][j];  // xi^((q^2-1)/3)
  nQ2.x = mul_fq(Q.x, k);
  nQ2.y = Q.y;
  miller_add(f, R, Q1, P.x, P.y);
  miller_add(f, R, nQ2, P.x, P.y);
  return f;
}

// a^2 for a in the cyclotomic subgroup (a^(q^6+1) = 1, true after the easy part of the final exponentiation):
// Granger-Scott squaring, three Fq4 squarings = 6 Fq2 products instead of 12
ZKB_HD void fq4_sqr(const Fq2& a, const Fq2& b, Fq2& t0, Fq2& t1) {  // (a + b y)^2, y^2 = xi
  Fq2 tmp = a * b;
  t0 = (a + b) * (mul_xi(b) + a) - tmp - mul_xi(tmp);
  t1 = dbl(tmp);
}
ZKB_OOL Fq12 cyclotomic_sqr(const Fq12& a) {
  const Fq2 &z0 = a.c[0].c[0], &z4 = a.c[0].c[1], &z3 = a.c[0].c[2], &z2 = a.c[1].c[0], &z1 = a.c[1].c[1], &z5 = a.c[1].c[2];
  Fq2 t0, t1, t2, t3, t4, t5;
  fq4_sqr(z0, z1, t0, t1);
  fq4_sqr(z2, z3, t2, t3);
  fq4_sqr(z4, z5, t4, t5);
  Fq2 t5x = mul_xi(t5);
  Fq12 r;
  r.c[0].c[0] = dbl(t0 - z0) + t0;   // 3 t0 - 2 z0
  r.c[1].c[1] = dbl(t1 + z1) + t1;   // 3 t1 + 2 z1
  r.c[1].c[0] = dbl(t5x + z2) + t5x; // 3 xi t5 + 2 z2
  r.c[0].c[2] = dbl(t4 - z3) + t4;   // 3 t4 - 2 z3
  r.c[0].c[1] = dbl(t2 - z4) + t2;   // 3 t2 - 2 z4
  r.c[1].c[2] = dbl(t3 + z5) + t3;   // 3 t3 + 2 z5
  return r;
}

ZKB_OOL Fq12 pow_u(const Fq12& a) {  // a^u, u = 4965661367192848881 (63 bits), a in the cyclotomic subgroup
  const uint64_t u = ZKB_BN_U;
  Fq12 acc = a;
  for (int i = 61; i >= 0; i--) {
    acc = cyclotomic_sqr(acc);
    if ((u >> i) & 1ull) acc = acc * a;
  }
  return acc;
}

// f^((q^12 - 1)/r) = ((f^(q^6 - 1))^(q^2 + 1))^((q^4 - q^2 + 1)/r); the hard part by the
// lambda_0 + lambda_1 q + lambda_2 q^2 + lambda_3 q^3 decomposition with three powers of u
// (Scott et al.; exact, not a multiple: checked against the plain exponent in tests/test_host_pairing.py)
ZKB_HD Fq12 final_exponentiation(const Fq12& f) {
  Fq12 t = conj(f) * inverse(f);
  t = frobenius2(t) * t;
  Fq12 fp = frobenius(t), fp2 = frobenius2(t);
  Fq12 fp3 = frobenius(fp2);
  Fq12 fu = pow_u(t);
  Fq12 fu2 = pow_u(fu);
  Fq12 fu3 = pow_u(fu2);
  Fq12 y0 = fp * fp2 * fp3;
  Fq12 y1 = conj(t);
  Fq12 y2 = frobenius2(fu2);
  Fq12 y3 = conj(frobenius(fu));
  Fq12 y4 = conj(fu * frobenius(fu2));
  Fq12 y5 = conj(fu2);
  Fq12 y6 = conj(fu3 * frobenius(fu3));
  Fq12 t0 = sqr(y6) * y4 * y5;
  Fq12 t1 = y3 * y5 * t0;
  t0 = t0 * y2;
  t1 = sqr(sqr(t1) * t0);
  t0 = t1 * y1;
  t1 = t1 * y0;
  return sqr(t0) * t1;
}

// ---- plain versions (affine Miller steps with one inversion each; the hard part as a 761-bit
// exponent): the definition the fast versions above are checked against
ZKB_OOL void miller_step_affine(Fq12& f, Affine<Fq2>& R, const Affine<Fq2>& S, bool tangent, const Fq& xP, const Fq& yP) {
  Fq2 lambda;
  if (tangent) { Fq2 xx = sqr(R.x); lambda = (dbl(xx) + xx) * inverse(dbl(R.y)); }
  else lambda = (S.y - R.y) * inverse(S.x - R.x);
  f = f * line_value(lambda, R.x, R.y, xP, yP);
  Fq2 x3 = sqr(lambda) - R.x - S.x;
  R.y = lambda * (R.x - x3) - R.y;
  R.x = x3;
}
ZKB_HD Fq12 miller_loop_affine(const Affine<Fq>& P, const Affine<Fq2>& Q) {
  Fq12 f = Fq12::one();
  if (P.is_inf() || Q.is_inf()) return f;
  const uint32_t loop[3] = ZKB_ATE_LOOP;
  Affine<Fq2> R = Q;
  for (int i = ZKB_ATE_LOOP_BITS - 2; i >= 0; i--) {
    f = sqr(f);
    miller_step_affine(f, R, R, true, P.x, P.y);
    if ((loop[i >> 5] >> (i & 31)) & 1u) miller_step_affine(f, R, Q, false, P.x, P.y);
  }
  const uint32_t g1[12][8] = ZKB_FROB1_W;
  const uint32_t g2[6][8] = ZKB_FROB2_W;
  Affine<Fq2> Q1, nQ2;
  Q1.x = conj(Q.x) * fq2_const(g1[4], g1[5]);
  Q1.y = conj(Q.y) * fq2_const(g1[6], g1[7]);
  Fq k; for (int j = 0; j < 8; j++) k.v[j] = g2[2][j];
  nQ2.x = mul_fq(Q.x, k);
  nQ2.y = Q.y;
  miller_step_affine(f, R, Q1, false, P.x, P.y);
  Fq2 lambda = (nQ2.y - R.y) * inverse(nQ2.x - R.x);
  return f * line_value(lambda, R.x, R.y, P.x, P.y);
}
ZKB_HD Fq12 final_exponentiation_plain(const Fq12& f) {
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
