/* Oracle B -- C restatement of the reference's groth16::prove() path.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): used by tests/ as a second checker and by
 * bench.py's cpu_baseline / --impl reference legs as the timed CPU implementation.  Never linked
 * into, loaded by, or called from the product (libzkb200.so / zksnark-rs_b200/).
 *
 * What it follows (all paths under /root/reference/src):
 *   groth16/mod.rs:213-296        prove(): dense weighted sums, 5 MSMs as per-term scalar
 *                                 multiplications folded sequentially, h = (u*v - w)/t, assembly
 *   groth16/coefficient_poly.rs   Add :24-49 (length = max), Neg :51-62, Sub :64-73, Sum :75-91 (seed
 *                                 [0]), Mul :93-130 (schoolbook, strips leading zeros first),
 *                                 Mul<T> :132-146, Div :148-157
 *   field/mod.rs:428-469          polynomial_division (long division, one inversion per step,
 *                                 re-scan for the degree every step), :291-297 degree, :344-355
 *                                 remove_leading_zeros
 *   groth16/fr.rs:101-123,175-223 exp_encrypted_g1/g2 = one scalar multiplication each; Sum = fold
 * Arithmetic of crate `bn` 0.4.3 (not vendored in the reference tree, Cargo.toml:15) is restated from
 * the published BN254 definition: 4x64-bit Montgomery fields, Jacobian points (identity z = 0),
 * MSB-first double-and-add.  Every constant is DERIVED here from the two primes at start-up
 * (independently of zksnark-rs_b200/csrc/constants.h), so agreement with the CUDA path is a real
 * cross-check.  Pinning: "parity unpinned" at the bn boundary (no fixed-value test exists in the
 * reference); this file is pinned against Oracle A (Python), which reproduces all of the
 * reference's Z251 golden vectors -- tests/test_oracle_b.py.
 *
 * All values cross this file's ABI as 4 x uint64 little-endian canonical limbs; points affine,
 * identity = all-zero.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef unsigned __int128 u128;
typedef uint64_t u64;
typedef struct { u64 l[4]; } fe;
typedef struct { u64 p[4]; u64 ninv; fe one, r2; } field;

static field FR, FQ;
static int g_init = 0;

/* ---- multi-precision helpers ---- */
static int ge4(const u64* a, const u64* b) {
  for (int i = 3; i >= 0; i--) { if (a[i] != b[i]) return a[i] > b[i]; }
  return 1;
}
#if defined(__x86_64__)
#include <immintrin.h>
static u64 add4(u64* r, const u64* a, const u64* b) {
  unsigned long long t; unsigned char c = 0;
  for (int i = 0; i < 4; i++) { c = _addcarry_u64(c, a[i], b[i], &t); r[i] = t; }
  return c;
}
static u64 sub4(u64* r, const u64* a, const u64* b) {
  unsigned long long t; unsigned char c = 0;
  for (int i = 0; i < 4; i++) { c = _subborrow_u64(c, a[i], b[i], &t); r[i] = t; }
  return c;
}
#else
static u64 add4(u64* r, const u64* a, const u64* b) {
  u128 c = 0;
  for (int i = 0; i < 4; i++) { c += (u128)a[i] + b[i]; r[i] = (u64)c; c >>= 64; }
  return (u64)c;
}
static u64 sub4(u64* r, const u64* a, const u64* b) {
  u64 br = 0;
  for (int i = 0; i < 4; i++) {
    u128 t = (u128)a[i] - b[i] - br;
    r[i] = (u64)t; br = (u64)(t >> 64) & 1;
  }
  return br;
}
#endif
static fe fe_add(const field* F, fe a, fe b) {
  fe r; u64 c = add4(r.l, a.l, b.l);
  if (c || ge4(r.l, F->p)) sub4(r.l, r.l, F->p);
  return r;
}
static fe fe_sub(const field* F, fe a, fe b) {
  fe r; if (sub4(r.l, a.l, b.l)) add4(r.l, r.l, F->p);
  return r;
}
static int fe_is_zero(fe a) { return (a.l[0] | a.l[1] | a.l[2] | a.l[3]) == 0; }
static int fe_eq(fe a, fe b) { return a.l[0] == b.l[0] && a.l[1] == b.l[1] && a.l[2] == b.l[2] && a.l[3] == b.l[3]; }
static fe fe_neg(const field* F, fe a) { fe z = {{0, 0, 0, 0}}; return fe_sub(F, z, a); }

/* Montgomery product (CIOS, rows unrolled; both moduli are 254-bit, so five words hold every partial sum) */
static fe fe_mul(const field* F, fe a, fe b) {
  const u64 p0 = F->p[0], p1 = F->p[1], p2 = F->p[2], p3 = F->p[3], ninv = F->ninv;
  u64 t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0;
  for (int i = 0; i < 4; i++) {
    const u64 bi = b.l[i];
    u128 c = (u128)a.l[0] * bi + t0; t0 = (u64)c; c >>= 64;
    c += (u128)a.l[1] * bi + t1; t1 = (u64)c; c >>= 64;
    c += (u128)a.l[2] * bi + t2; t2 = (u64)c; c >>= 64;
    c += (u128)a.l[3] * bi + t3; t3 = (u64)c; c >>= 64;
    t4 += (u64)c;
    const u64 m = t0 * ninv;
    c = (u128)m * p0 + t0; c >>= 64;
    c += (u128)m * p1 + t1; t0 = (u64)c; c >>= 64;
    c += (u128)m * p2 + t2; t1 = (u64)c; c >>= 64;
    c += (u128)m * p3 + t3; t2 = (u64)c; c >>= 64;
    c += t4; t3 = (u64)c; t4 = (u64)(c >> 64);
  }
  fe r = {{t0, t1, t2, t3}};
  if (t4 || ge4(r.l, F->p)) sub4(r.l, r.l, F->p);
  return r;
}
static fe fe_sqr(const field* F, fe a) { return fe_mul(F, a, a); }
static fe fe_pow(const field* F, fe a, const u64 e[4]) {
  fe acc = F->one;
  for (int i = 255; i >= 0; i--) {
    acc = fe_sqr(F, acc);
    if ((e[i >> 6] >> (i & 63)) & 1) acc = fe_mul(F, acc, a);
  }
  return acc;
}
static fe fe_inv(const field* F, fe a) { /* a^(p-2); inverse(0) = 0, callers check */
  u64 e[4]; u64 two[4] = {2, 0, 0, 0};
  sub4(e, F->p, two);
  return fe_pow(F, a, e);
}
static fe fe_from_canon(const field* F, const u64* c) { fe a; memcpy(a.l, c, 32); return fe_mul(F, a, F->r2); }
static void fe_to_canon(const field* F, fe a, u64* out) {
  fe one = {{1, 0, 0, 0}}; fe r = fe_mul(F, a, one); memcpy(out, r.l, 32);
}

static void field_init(field* F, const u64 p[4]) {
  memcpy(F->p, p, 32);
  u64 x = 1; /* Newton: x = p^-1 mod 2^64 */
  for (int i = 0; i < 6; i++) x *= 2 - p[0] * x;
  F->ninv = (u64)0 - x;
  /* one = 2^256 mod p by 256 modular doublings of 1; r2 = 2^512 mod p by 256 more */
  fe v = {{1, 0, 0, 0}};
  for (int i = 0; i < 512; i++) {
    v = fe_add(F, v, v);
    if (i == 255) F->one = v;
  }
  F->r2 = v;
}
static void ob_init(void) {
  if (g_init) return;
  /* r and q of BN254 (alt_bn128) */
  static const u64 r[4] = {0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
  static const u64 q[4] = {0x3c208c16d87cfd47ull, 0x97816a916871ca8dull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
  field_init(&FR, r); field_init(&FQ, q);
  g_init = 1;
}

/* ---- Fq2 = Fq[u]/(u^2+1) ---- */
typedef struct { fe c0, c1; } fe2;
static fe2 f2_add(fe2 a, fe2 b) { fe2 r = {fe_add(&FQ, a.c0, b.c0), fe_add(&FQ, a.c1, b.c1)}; return r; }
static fe2 f2_sub(fe2 a, fe2 b) { fe2 r = {fe_sub(&FQ, a.c0, b.c0), fe_sub(&FQ, a.c1, b.c1)}; return r; }
static fe2 f2_mul(fe2 a, fe2 b) { /* Karatsuba: 3 base-field products (as crate bn's Fq2 does) */
  fe aa = fe_mul(&FQ, a.c0, b.c0), bb = fe_mul(&FQ, a.c1, b.c1);
  fe2 r;
  r.c0 = fe_sub(&FQ, aa, bb);
  r.c1 = fe_sub(&FQ, fe_sub(&FQ, fe_mul(&FQ, fe_add(&FQ, a.c0, a.c1), fe_add(&FQ, b.c0, b.c1)), aa), bb);
  return r;
}
static fe2 f2_sqr(fe2 a) { /* complex squaring: (a0 + a1)(a0 - a1) + 2 a0 a1 u */
  fe ab = fe_mul(&FQ, a.c0, a.c1);
  fe2 r;
  r.c0 = fe_mul(&FQ, fe_add(&FQ, a.c0, a.c1), fe_sub(&FQ, a.c0, a.c1));
  r.c1 = fe_add(&FQ, ab, ab);
  return r;
}
static int f2_is_zero(fe2 a) { return fe_is_zero(a.c0) && fe_is_zero(a.c1); }
static int f2_eq(fe2 a, fe2 b) { return fe_eq(a.c0, b.c0) && fe_eq(a.c1, b.c1); }
static fe2 f2_inv(fe2 a) {
  fe n = fe_inv(&FQ, fe_add(&FQ, fe_sqr(&FQ, a.c0), fe_sqr(&FQ, a.c1)));
  fe2 r = {fe_mul(&FQ, a.c0, n), fe_neg(&FQ, fe_mul(&FQ, a.c1, n))};
  return r;
}

/* ---- Jacobian group law, written once for both coordinate fields ---- */
#define DEFINE_GROUP(G, T, ADD, SUB, MUL, SQR, ISZ, EQ, INV, ONE)                                   \
  typedef struct { T x, y, z; } G;                                                                  \
  static G G##_zero(void) { G r; memset(&r, 0, sizeof r); r.x = ONE; r.y = ONE; return r; }         \
  static int G##_is_zero(G p) { return ISZ(p.z); }                                                  \
  static G G##_dbl(G p) { /* dbl-2009-l */                                                          \
    if (G##_is_zero(p)) return p;                                                                   \
    T a = SQR(p.x), b = SQR(p.y), c = SQR(b);                                                       \
    T d = SUB(SUB(SQR(ADD(p.x, b)), a), c); d = ADD(d, d);                                          \
    T e = ADD(ADD(a, a), a), f = SQR(e);                                                            \
    G r; r.x = SUB(f, ADD(d, d));                                                                   \
    T c8 = ADD(c, c); c8 = ADD(c8, c8); c8 = ADD(c8, c8);                                           \
    r.z = MUL(p.y, p.z); r.z = ADD(r.z, r.z);                                                       \
    r.y = SUB(MUL(e, SUB(d, r.x)), c8);                                                             \
    return r;                                                                                       \
  }                                                                                                 \
  static G G##_add(G p, G q) { /* add-2007-bl */                                                    \
    if (G##_is_zero(p)) return q;                                                                   \
    if (G##_is_zero(q)) return p;                                                                   \
    T z1z1 = SQR(p.z), z2z2 = SQR(q.z);                                                             \
    T u1 = MUL(p.x, z2z2), u2 = MUL(q.x, z1z1);                                                     \
    T s1 = MUL(MUL(p.y, q.z), z2z2), s2 = MUL(MUL(q.y, p.z), z1z1);                                 \
    if (EQ(u1, u2)) { if (EQ(s1, s2)) return G##_dbl(p); return G##_zero(); }                     \
    T h = SUB(u2, u1), i = SQR(ADD(h, h)), j = MUL(h, i);                                           \
    T rr = SUB(s2, s1); rr = ADD(rr, rr);                                                           \
    T v = MUL(u1, i);                                                                               \
    G r; r.x = SUB(SUB(SQR(rr), j), ADD(v, v));                                                     \
    T s1j = MUL(s1, j); s1j = ADD(s1j, s1j);                                                        \
    r.y = SUB(MUL(rr, SUB(v, r.x)), s1j);                                                           \
    r.z = MUL(SUB(SUB(SQR(ADD(p.z, q.z)), z1z1), z2z2), h);                                         \
    return r;                                                                                       \
  }                                                                                                 \
  /* `Group * Fr` of crate bn: MSB-first double-and-add over the 256-bit scalar */                  \
  static G G##_mul(G p, const u64 k[4]) {                                                           \
    G acc = G##_zero(); int found = 0;                                                            \
    for (int i = 255; i >= 0; i--) {                                                                \
      if (found) acc = G##_dbl(acc);                                                                \
      if ((k[i >> 6] >> (i & 63)) & 1) { found = 1; acc = G##_add(acc, p); }                        \
    }                                                                                               \
    return acc;                                                                                     \
  }                                                                                                 \
  static void G##_to_affine(G p, T* x, T* y) {                                                      \
    if (G##_is_zero(p)) { memset(x, 0, sizeof *x); memset(y, 0, sizeof *y); return; }               \
    T zi = INV(p.z), zi2 = SQR(zi);                                                                 \
    *x = MUL(p.x, zi2); *y = MUL(p.y, MUL(zi2, zi));                                                \
  }

static fe q_add(fe a, fe b) { return fe_add(&FQ, a, b); }
static fe q_sub(fe a, fe b) { return fe_sub(&FQ, a, b); }
static fe q_mul(fe a, fe b) { return fe_mul(&FQ, a, b); }
static fe q_sqr(fe a) { return fe_sqr(&FQ, a); }
static fe q_inv(fe a) { return fe_inv(&FQ, a); }
#define Q_ONE (FQ.one)
static fe2 f2_one(void) { fe2 r; memset(&r, 0, sizeof r); r.c0 = FQ.one; return r; }
#define F2_ONE f2_one()

DEFINE_GROUP(g1, fe, q_add, q_sub, q_mul, q_sqr, fe_is_zero, fe_eq, q_inv, Q_ONE)
DEFINE_GROUP(g2, fe2, f2_add, f2_sub, f2_mul, f2_sqr, f2_is_zero, f2_eq, f2_inv, F2_ONE)

static g1 g1_from_affine(const u64* p) { /* 8 limbs canonical; all-zero = identity */
  g1 r = g1_zero();
  int z = 1; for (int i = 0; i < 8; i++) z &= p[i] == 0;
  if (z) return r;
  r.x = fe_from_canon(&FQ, p); r.y = fe_from_canon(&FQ, p + 4); r.z = FQ.one;
  return r;
}
static void g1_store(g1 p, u64* out) {
  if (g1_is_zero(p)) { memset(out, 0, 64); return; }
  fe x, y; g1_to_affine(p, &x, &y);
  fe_to_canon(&FQ, x, out); fe_to_canon(&FQ, y, out + 4);
}
static g2 g2_from_affine(const u64* p) { /* 16 limbs: x.c0 x.c1 y.c0 y.c1 */
  g2 r = g2_zero();
  int z = 1; for (int i = 0; i < 16; i++) z &= p[i] == 0;
  if (z) return r;
  r.x.c0 = fe_from_canon(&FQ, p); r.x.c1 = fe_from_canon(&FQ, p + 4);
  r.y.c0 = fe_from_canon(&FQ, p + 8); r.y.c1 = fe_from_canon(&FQ, p + 12);
  r.z = f2_one();
  return r;
}
static void g2_store(g2 p, u64* out) {
  if (g2_is_zero(p)) { memset(out, 0, 128); return; }
  fe2 x, y; g2_to_affine(p, &x, &y);
  fe_to_canon(&FQ, x.c0, out); fe_to_canon(&FQ, x.c1, out + 4);
  fe_to_canon(&FQ, y.c0, out + 8); fe_to_canon(&FQ, y.c1, out + 12);
}
static g1 g1_neg(g1 p) { p.y = fe_neg(&FQ, p.y); return p; }

/* ---- exported field / group primitives (cross-checked against Oracle A) ---- */
void ob_fr_mul(const u64* a, const u64* b, u64* out) {
  ob_init(); fe_to_canon(&FR, fe_mul(&FR, fe_from_canon(&FR, a), fe_from_canon(&FR, b)), out);
}
void ob_fr_inv(const u64* a, u64* out) { ob_init(); fe_to_canon(&FR, fe_inv(&FR, fe_from_canon(&FR, a)), out); }
void ob_fq_mul(const u64* a, const u64* b, u64* out) {
  ob_init(); fe_to_canon(&FQ, fe_mul(&FQ, fe_from_canon(&FQ, a), fe_from_canon(&FQ, b)), out);
}
void ob_g1_add(const u64* a, const u64* b, u64* out) { ob_init(); g1_store(g1_add(g1_from_affine(a), g1_from_affine(b)), out); }
void ob_g1_mul(const u64* a, const u64* k, u64* out) { ob_init(); g1_store(g1_mul(g1_from_affine(a), k), out); }
void ob_g2_add(const u64* a, const u64* b, u64* out) { ob_init(); g2_store(g2_add(g2_from_affine(a), g2_from_affine(b)), out); }
void ob_g2_mul(const u64* a, const u64* k, u64* out) { ob_init(); g2_store(g2_mul(g2_from_affine(a), k), out); }

/* `.zip().map(exp_encrypted_g1).sum()` (mod.rs:255-260): per-term scalar-mul, sequential fold from zero */
static g1 msm_g1_naive(const fe* scal /*Fr mont*/, const u64* pts, size_t n) {
  g1 acc = g1_zero();
  for (size_t i = 0; i < n; i++) {
    u64 k[4]; fe_to_canon(&FR, scal[i], k);
    acc = g1_add(acc, g1_mul(g1_from_affine(pts + 8 * i), k));
  }
  return acc;
}
static g2 msm_g2_naive(const fe* scal, const u64* pts, size_t n) {
  g2 acc = g2_zero();
  for (size_t i = 0; i < n; i++) {
    u64 k[4]; fe_to_canon(&FR, scal[i], k);
    acc = g2_add(acc, g2_mul(g2_from_affine(pts + 16 * i), k));
  }
  return acc;
}
void ob_msm_g1(const u64* scalars, const u64* pts, size_t n, u64* out) {
  ob_init();
  fe* s = malloc((n + 1) * sizeof(fe));
  for (size_t i = 0; i < n; i++) s[i] = fe_from_canon(&FR, scalars + 4 * i);
  g1_store(msm_g1_naive(s, pts, n), out); free(s);
}
void ob_msm_g2(const u64* scalars, const u64* pts, size_t n, u64* out) {
  ob_init();
  fe* s = malloc((n + 1) * sizeof(fe));
  for (size_t i = 0; i < n; i++) s[i] = fe_from_canon(&FR, scalars + 4 * i);
  g2_store(msm_g2_naive(s, pts, n), out); free(s);
}

/* ---- CoefficientPoly over Fr: (len, coeffs), little-endian degree ---- */
typedef struct { size_t len; fe* c; } poly;
static poly poly_new(size_t len) { poly p; p.len = len; p.c = calloc(len ? len : 1, sizeof(fe)); return p; }
static void poly_free(poly p) { free(p.c); }
static poly poly_clone(poly a) { poly r = poly_new(a.len); memcpy(r.c, a.c, a.len * sizeof(fe)); return r; }
/* field/mod.rs:291-297: index of the last non-zero coefficient, 0 for empty/zero polys */
static size_t poly_degree(poly a) {
  for (size_t i = a.len; i > 0; i--) if (!fe_is_zero(a.c[i - 1])) return i - 1;
  return 0;
}
/* field/mod.rs:344-355: strip trailing zeros but keep at least one coefficient */
static void poly_strip(poly* a) {
  while (a->len > 1 && fe_is_zero(a->c[a->len - 1])) a->len--;
}
/* coefficient_poly.rs:24-49 */
static poly poly_add(poly a, poly b) {
  size_t n = a.len > b.len ? a.len : b.len;
  poly r = poly_new(n);
  for (size_t i = 0; i < n; i++) {
    fe x = i < a.len ? a.c[i] : (fe){{0, 0, 0, 0}}, y = i < b.len ? b.c[i] : (fe){{0, 0, 0, 0}};
    r.c[i] = fe_add(&FR, x, y);
  }
  return r;
}
static poly poly_neg(poly a) { poly r = poly_new(a.len); for (size_t i = 0; i < a.len; i++) r.c[i] = fe_neg(&FR, a.c[i]); return r; }
static poly poly_sub(poly a, poly b) { poly nb = poly_neg(b); poly r = poly_add(a, nb); poly_free(nb); return r; }
/* coefficient_poly.rs:132-146 */
static poly poly_scale(poly a, fe s) { poly r = poly_new(a.len); for (size_t i = 0; i < a.len; i++) r.c[i] = fe_mul(&FR, a.c[i], s); return r; }
/* coefficient_poly.rs:93-130: schoolbook; operands stripped first; zero operand -> [0] */
static poly poly_mul(poly a_in, poly b_in) {
  poly a = a_in, b = b_in; poly_strip(&a); poly_strip(&b);
  if ((a.len == 1 && fe_is_zero(a.c[0])) || (b.len == 1 && fe_is_zero(b.c[0])) || a.len == 0 || b.len == 0) return poly_new(1);
  poly r = poly_new(a.len + b.len - 1);
  for (size_t i = 0; i < a.len; i++)
    for (size_t j = 0; j < b.len; j++) r.c[i + j] = fe_add(&FR, r.c[i + j], fe_mul(&FR, a.c[i], b.c[j]));
  return r;
}
/* field/mod.rs:428-469, statement by statement.  rc -1: the divisor is the zero polynomial (:433-441). */
static int poly_divrem(poly num, poly den, poly* quo, poly* rem) {
  int all_zero = 1;
  for (size_t i = 0; i < den.len; i++) if (!fe_is_zero(den.c[i])) all_zero = 0;
  if (all_zero) return -1;
  if (poly_degree(den) > poly_degree(num)) { /* :443-445 */
    *quo = poly_new(1); if (rem) *rem = poly_new(1); return 0;
  }
  poly r = poly_clone(num);
  while (r.len > 0 && fe_is_zero(r.c[r.len - 1])) r.len--; /* remove_leading_zeros may leave [] */
  size_t d = poly_degree(den);
  poly q = poly_new(poly_degree(r) + 1 - d);
  fe c = den.c[d];
  while (poly_degree(r) >= d && r.len != 0) {
    size_t dr = poly_degree(r);
    fe s = fe_mul(&FR, r.c[dr], fe_inv(&FR, c)); /* `/` = one inversion per step, fr.rs:50-56 */
    q.c[dr - d] = s;
    for (size_t k = 0; k <= d && k <= dr; k++) r.c[dr - k] = fe_sub(&FR, r.c[dr - k], fe_mul(&FR, den.c[d - k], s));
    while (r.len > 0 && fe_is_zero(r.c[r.len - 1])) r.len--; /* full re-scan, :465 */
  }
  *quo = q;
  if (rem) *rem = r; else poly_free(r);
  return 0;
}

void ob_poly_mul(const u64* a, size_t na, const u64* b, size_t nb, u64* out, size_t* nout) {
  ob_init();
  poly pa = poly_new(na), pb = poly_new(nb);
  for (size_t i = 0; i < na; i++) pa.c[i] = fe_from_canon(&FR, a + 4 * i);
  for (size_t i = 0; i < nb; i++) pb.c[i] = fe_from_canon(&FR, b + 4 * i);
  poly r = poly_mul(pa, pb);
  for (size_t i = 0; i < r.len; i++) fe_to_canon(&FR, r.c[i], out + 4 * i);
  *nout = r.len;
  poly_free(pa); poly_free(pb); poly_free(r);
}
int ob_poly_div(const u64* a, size_t na, const u64* b, size_t nb, u64* out, size_t* nout) {
  ob_init();
  poly pa = poly_new(na), pb = poly_new(nb), q;
  for (size_t i = 0; i < na; i++) pa.c[i] = fe_from_canon(&FR, a + 4 * i);
  for (size_t i = 0; i < nb; i++) pb.c[i] = fe_from_canon(&FR, b + 4 * i);
  int rc = poly_divrem(pa, pb, &q, NULL);
  if (rc == 0) { for (size_t i = 0; i < q.len; i++) fe_to_canon(&FR, q.c[i], out + 4 * i); *nout = q.len; poly_free(q); }
  poly_free(pa); poly_free(pb);
  return rc;
}

/* ---- groth16::prove, dense QAP (mod.rs:213-296) ----
 * u, v, w: m rows x `stride` coefficients each (row i uses the first row_len[i] of them; the
 * reference's polys have individual lengths).  t: nt coefficients.  CRS vectors as in zkb_crs_host.
 * weights: nw scalars.  Output proof: a (8), b (16), c (8) limbs. */
typedef struct {
  size_t m, stride, nt, n_input;
  const u64 *u, *v, *w, *t;
  const size_t *ulen, *vlen, *wlen;
  size_t n_xi, n_xit, n_sd;
  const u64 *alpha1, *beta1, *delta1, *xi1, *xi_t, *sum_delta, *beta2, *delta2, *xi2;
} ob_prove_in;

static poly weighted_sum(const u64* rows, const size_t* lens, size_t m, size_t stride, const fe* wts, size_t nw) {
  poly acc = poly_new(1); /* Sum seeds with [0], coefficient_poly.rs:84-89 */
  size_t cnt = m < nw ? m : nw; /* zip */
  for (size_t i = 0; i < cnt; i++) {
    poly row = poly_new(lens[i]); /* .clone() of the row, mod.rs:235 */
    for (size_t k = 0; k < lens[i]; k++) row.c[k] = fe_from_canon(&FR, rows + 4 * (i * stride + k));
    poly sc = poly_scale(row, wts[i]);
    poly nx = poly_add(acc, sc);
    poly_free(row); poly_free(sc); poly_free(acc);
    acc = nx;
  }
  return acc;
}

int ob_prove(const ob_prove_in* in, const u64* weights, size_t nw, const u64* r_, const u64* s_, u64* proof,
             u64* h_out, size_t* h_len) {
  ob_init();
  fe* wts = malloc((nw + 1) * sizeof(fe));
  for (size_t i = 0; i < nw; i++) wts[i] = fe_from_canon(&FR, weights + 4 * i);
  fe r = fe_from_canon(&FR, r_), s = fe_from_canon(&FR, s_);
  poly us = weighted_sum(in->u, in->ulen, in->m, in->stride, wts, nw);
  poly vs = weighted_sum(in->v, in->vlen, in->m, in->stride, wts, nw);
  poly ws = weighted_sum(in->w, in->wlen, in->m, in->stride, wts, nw);
  size_t ka = us.len < in->n_xi ? us.len : in->n_xi, kb = vs.len < in->n_xi ? vs.len : in->n_xi;
  g1 a_g1 = msm_g1_naive(us.c, in->xi1, ka);
  g1 b_g1 = msm_g1_naive(vs.c, in->xi1, kb);
  g2 b_g2 = msm_g2_naive(vs.c, in->xi2, kb);
  g1 delta1 = g1_from_affine(in->delta1);
  g2 delta2 = g2_from_affine(in->delta2);
  u64 rk[4], sk[4], rsk[4];
  fe_to_canon(&FR, r, rk); fe_to_canon(&FR, s, sk); fe_to_canon(&FR, fe_mul(&FR, r, s), rsk);
  g1 a = g1_add(g1_add(a_g1, g1_from_affine(in->alpha1)), g1_mul(delta1, rk));
  g2 b = g2_add(g2_add(b_g2, g2_from_affine(in->beta2)), g2_mul(delta2, sk));
  /* h = (u_sum * v_sum - w_sum) / t */
  poly t = poly_new(in->nt);
  for (size_t i = 0; i < in->nt; i++) t.c[i] = fe_from_canon(&FR, in->t + 4 * i);
  poly uv = poly_mul(us, vs), num = poly_sub(uv, ws), h;
  if (poly_divrem(num, t, &h, NULL) != 0) return -5;
  if (h_out) { for (size_t i = 0; i < h.len; i++) fe_to_canon(&FR, h.c[i], h_out + 4 * i); *h_len = h.len; }
  size_t kh = h.len < in->n_xit ? h.len : in->n_xit;
  g1 c = msm_g1_naive(h.c, in->xi_t, kh);
  size_t skip = in->n_input + 1;
  size_t kw = nw > skip ? nw - skip : 0; if (kw > in->n_sd) kw = in->n_sd;
  c = g1_add(c, msm_g1_naive(wts + (nw > skip ? skip : nw), in->sum_delta, kw));
  /* a is normalised to affine before s*a in bn?  No: group ops are projective; result identical. */
  c = g1_add(c, g1_mul(a, sk));
  g1 inner = g1_add(g1_add(g1_from_affine(in->beta1), b_g1), g1_mul(delta1, sk));
  c = g1_add(c, g1_mul(inner, rk));
  c = g1_add(c, g1_neg(g1_mul(delta1, rsk)));
  g1_store(a, proof); g2_store(b, proof + 8); g1_store(c, proof + 24);
  poly_free(us); poly_free(vs); poly_free(ws); poly_free(t); poly_free(uv); poly_free(num); poly_free(h);
  free(wts);
  return 0;
}

/* ---- the reference's prove() run IN FULL on a roots-of-unity QAP given as sparse evaluation rows (bench.py:
 * cpu_baseline_measured).  The dense QAP<CoefficientPoly<Fr>> the reference holds (mod.rs:60-67: 3*m coefficient
 * vectors of n entries) is built here first, OUTSIDE the timed region -- in the reference that is `QAP::from`
 * (fr.rs:140-173), a separate call -- from L_k(x) = (1/n) sum_j w^(-kj) x^j; the timed region is ob_prove alone, i.e.
 * exactly mod.rs:213-296 with the reference's algorithms.  row_ptr: m+1 offsets; gate: 0-based gate indices;
 * coeff: 4 limbs each.  omega_inv: w^-1 (canonical).  Returns seconds of ob_prove, or a negative value on failure. */
typedef struct {
  size_t log_n, m, n_input;
  const u64* row_ptr[3];
  const uint32_t* gate[3];
  const u64* coeff[3];
  const u64* omega_inv;
  size_t n_sd;
  const u64 *alpha1, *beta1, *delta1, *xi1, *xi_t, *sum_delta, *beta2, *delta2, *xi2;
} ob_full_in;

static double now(void);
double ob_prove_full_omega(const ob_full_in* in, const u64* weights, size_t nw, const u64* r_, const u64* s_, u64* proof,
                           double* dense_build_s) {
  ob_init();
  const size_t n = (size_t)1 << in->log_n, m = in->m;
  double t0 = now();
  fe winv = fe_from_canon(&FR, in->omega_inv);
  fe* pw = malloc(n * sizeof(fe)); /* w^-e, e < n */
  pw[0] = FR.one;
  for (size_t e = 1; e < n; e++) pw[e] = fe_mul(&FR, pw[e - 1], winv);
  u64 ncan[4] = {(u64)n, 0, 0, 0};
  fe ninv = fe_inv(&FR, fe_from_canon(&FR, ncan));
  u64* dense[3];
  size_t* lens[3];
  for (int t = 0; t < 3; t++) {
    dense[t] = calloc(m * n * 4, sizeof(u64));
    lens[t] = malloc(m * sizeof(size_t));
    if (!dense[t] || !lens[t]) return -1.0;
    fe* acc = malloc(n * sizeof(fe));
    for (size_t i = 0; i < m; i++) {
      size_t e0 = in->row_ptr[t][i], e1 = in->row_ptr[t][i + 1];
      if (e0 == e1) { lens[t][i] = 1; continue; } /* the sum of no basis polynomials: [0] (coefficient_poly.rs:84-89) */
      memset(acc, 0, n * sizeof(fe)); /* Montgomery form of 0 is 0 */
      for (size_t e = e0; e < e1; e++) {
        fe c = fe_mul(&FR, fe_from_canon(&FR, in->coeff[t] + 4 * e), ninv);
        size_t k = in->gate[t][e], idx = 0;
        for (size_t j = 0; j < n; j++) {
          acc[j] = fe_add(&FR, acc[j], fe_mul(&FR, c, pw[idx]));
          idx = (idx + k) & (n - 1);
        }
      }
      for (size_t j = 0; j < n; j++) fe_to_canon(&FR, acc[j], dense[t] + 4 * (i * n + j));
      lens[t][i] = n;
    }
    free(acc);
  }
  u64* tc = calloc((n + 1) * 4, sizeof(u64)); /* t = x^n - 1 */
  fe_to_canon(&FR, fe_neg(&FR, FR.one), tc);
  tc[4 * n] = 1;
  ob_prove_in pin;
  pin.m = m; pin.stride = n; pin.nt = n + 1; pin.n_input = in->n_input;
  pin.u = dense[0]; pin.v = dense[1]; pin.w = dense[2]; pin.t = tc;
  pin.ulen = lens[0]; pin.vlen = lens[1]; pin.wlen = lens[2];
  pin.n_xi = n; pin.n_xit = n - 1; pin.n_sd = in->n_sd;
  pin.alpha1 = in->alpha1; pin.beta1 = in->beta1; pin.delta1 = in->delta1; pin.xi1 = in->xi1; pin.xi_t = in->xi_t;
  pin.sum_delta = in->sum_delta; pin.beta2 = in->beta2; pin.delta2 = in->delta2; pin.xi2 = in->xi2;
  if (dense_build_s) *dense_build_s = now() - t0;
  t0 = now();
  int rc = ob_prove(&pin, weights, nw, r_, s_, proof, NULL, NULL);
  double dt = now() - t0;
  for (int t = 0; t < 3; t++) { free(dense[t]); free(lens[t]); }
  free(tc); free(pw);
  return rc == 0 ? dt : -2.0;
}

/* valid, pairwise distinct points for timing runs that have no device to make a CRS: P_i = (i + 1) * G by repeated addition */
void ob_fill_points(u64* g1_out, size_t n1, u64* g2_out, size_t n2, const u64* g2gen) {
  ob_init();
  u64 g1gen[8] = {1, 0, 0, 0, 2, 0, 0, 0};
  g1 G = g1_from_affine(g1gen), acc = G;
  for (size_t i = 0; i < n1; i++) { g1_store(acc, g1_out + 8 * i); acc = g1_add(acc, G); }
  g2 H = g2_from_affine(g2gen), acc2 = H;
  for (size_t i = 0; i < n2; i++) { g2_store(acc2, g2_out + 16 * i); acc2 = g2_add(acc2, H); }
}

/* ---- bounded timing samples of the reference algorithm at full problem width (bench.py) ----
 * Each returns seconds for the sampled work; bench.py scales to one proof. */
static double now(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }
static u64 sm64(u64* s) { u64 z = (*s += 0x9e3779b97f4a7c15ull); z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull; z = (z ^ (z >> 27)) * 0x94d049bb133111ebull; return z ^ (z >> 31); }
static fe rand_fr(u64* st) { u64 c[4] = {sm64(st), sm64(st), sm64(st), sm64(st) >> 4}; return fe_from_canon(&FR, c); }

/* k per-term G1 scalar multiplications + fold (HOT LOOP 3, mod.rs:255-266) */
double ob_time_g1_terms(size_t k, u64 seed) {
  ob_init();
  u64 st = seed; u64 kk[4];
  u64 gen[8] = {1, 0, 0, 0, 2, 0, 0, 0};
  g1 P = g1_from_affine(gen), acc = g1_zero();
  fe_to_canon(&FR, rand_fr(&st), kk); P = g1_mul(P, kk);
  u64 paff[8]; g1_store(P, paff);
  double t0 = now();
  for (size_t i = 0; i < k; i++) { fe_to_canon(&FR, rand_fr(&st), kk); acc = g1_add(acc, g1_mul(g1_from_affine(paff), kk)); }
  double t = now() - t0;
  g1_store(acc, paff);
  return paff[0] == 0xdeadbeef ? -t : t;
}
double ob_time_g2_terms(size_t k, u64 seed, const u64* g2gen) {
  ob_init();
  u64 st = seed; u64 kk[4];
  g2 P = g2_from_affine(g2gen), acc = g2_zero();
  u64 paff[16]; g2_store(P, paff);
  double t0 = now();
  for (size_t i = 0; i < k; i++) { fe_to_canon(&FR, rand_fr(&st), kk); acc = g2_add(acc, g2_mul(g2_from_affine(paff), kk)); }
  double t = now() - t0;
  g2_store(acc, paff);
  return paff[0] == 0xdeadbeef ? -t : t;
}
/* `rows` outer iterations of the schoolbook product of two length-n polys (coefficient_poly.rs:93-130) */
double ob_time_mul_rows(size_t n, size_t rows, u64 seed) {
  ob_init();
  u64 st = seed;
  fe* a = malloc(rows * sizeof(fe)); fe* b = malloc(n * sizeof(fe)); fe* c = calloc(n + rows, sizeof(fe));
  for (size_t i = 0; i < rows; i++) a[i] = rand_fr(&st);
  for (size_t j = 0; j < n; j++) b[j] = rand_fr(&st);
  double t0 = now();
  for (size_t i = 0; i < rows; i++)
    for (size_t j = 0; j < n; j++) c[i + j] = fe_add(&FR, c[i + j], fe_mul(&FR, a[i], b[j]));
  double t = now() - t0;
  u64 sink = c[n / 2].l[0];
  free(a); free(b); free(c);
  return sink == 0xdeadbeef ? -t : t;
}
/* `steps` outer steps of long division of a degree-(2n-2) poly by a degree-n one (field/mod.rs:428-469):
 * one inversion, n+1 mul-subs and a degree re-scan per step */
double ob_time_div_steps(size_t n, size_t steps, u64 seed) {
  ob_init();
  u64 st = seed;
  size_t ln = 2 * n - 1;
  fe* r = malloc(ln * sizeof(fe)); fe* d = malloc((n + 1) * sizeof(fe));
  for (size_t i = 0; i < ln; i++) r[i] = rand_fr(&st);
  for (size_t i = 0; i <= n; i++) d[i] = rand_fr(&st);
  poly pr = {ln, r};
  double t0 = now();
  size_t deg = ln - 1;
  for (size_t k = 0; k < steps && deg >= n; k++) {
    fe f = fe_mul(&FR, r[deg], fe_inv(&FR, d[n]));
    size_t sh = deg - n;
    for (size_t i = 0; i <= n; i++) r[sh + i] = fe_sub(&FR, r[sh + i], fe_mul(&FR, f, d[i]));
    deg = poly_degree(pr);
  }
  double t = now() - t0;
  u64 sink = r[0].l[0];
  free(r); free(d);
  return sink == 0xdeadbeef ? -t : t;
}
/* `rows` terms of a dense weighted sum over length-n rows (mod.rs:233-239): clone + scalar*poly + add */
double ob_time_wsum_rows(size_t n, size_t rows, u64 seed) {
  ob_init();
  u64 st = seed;
  poly row = poly_new(n), acc = poly_new(1);
  for (size_t j = 0; j < n; j++) row.c[j] = rand_fr(&st);
  double t0 = now();
  for (size_t i = 0; i < rows; i++) {
    poly cl = poly_clone(row); poly sc = poly_scale(cl, rand_fr(&st)); poly nx = poly_add(acc, sc);
    poly_free(cl); poly_free(sc); poly_free(acc); acc = nx;
  }
  double t = now() - t0;
  u64 sink = acc.c[0].l[0];
  poly_free(row); poly_free(acc);
  return sink == 0xdeadbeef ? -t : t;
}
