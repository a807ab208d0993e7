// Short-Weierstrass (a = 0) group arithmetic for BN254 G1 (over Fq) and G2 (over Fq2).
//
// Replaces crate `bn`'s G1/G2 `+`, `-`, `* Fr` as reached through
// /root/reference/src/groth16/fr.rs:114-119 (exp_encrypted_g1/g2) and :175-223 (Add/Sub/Sum).
// The reference folds Jacobian points one scalar-mul at a time; here accumulators are kept in
// extended-Jacobian "XYZZ" coordinates (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2), the cheapest mixed
// addition for bucket accumulation (8M+2S).  Results are only ever compared in AFFINE form, so the
// internal coordinate system is free.  All special cases (identity operands, P+P, P+(-P)) are
// handled: identity points are legal CRS entries (groth16/mod.rs:407).
#pragma once
#include "ff.cuh"

namespace zkb {

// Affine point.  Identity is encoded as (0, 0), which is not on either curve (b != 0).
template <class F>
struct alignas(16) Affine {
  F x, y;
  ZKB_HD bool is_inf() const { return x.is_zero() && y.is_zero(); }
  ZKB_HD static Affine inf() { Affine r; r.x = F::zero(); r.y = F::zero(); return r; }
};

// Extended Jacobian.  Identity <=> zz == 0.
template <class F>
struct alignas(16) XYZZ {
  F x, y, zz, zzz;
  ZKB_HD bool is_inf() const { return zz.is_zero(); }
  ZKB_HD static XYZZ inf() { XYZZ r; r.x = F::zero(); r.y = F::zero(); r.zz = F::zero(); r.zzz = F::zero(); return r; }
};

typedef Affine<Fq> G1Affine;
typedef Affine<Fq2> G2Affine;
typedef XYZZ<Fq> G1XYZZ;
typedef XYZZ<Fq2> G2XYZZ;

template <class F> ZKB_HD XYZZ<F> to_xyzz(const Affine<F>& p) {
  XYZZ<F> r;
  if (p.is_inf()) return XYZZ<F>::inf();
  r.x = p.x; r.y = p.y; r.zz = F::one(); r.zzz = F::one();
  return r;
}

template <class F> ZKB_HD Affine<F> neg(const Affine<F>& p) {
  Affine<F> r; r.x = p.x; r.y = neg(p.y); return r;  // -(0,0) = (0,0): identity stays identity
}
template <class F> ZKB_HD XYZZ<F> neg(const XYZZ<F>& p) {
  XYZZ<F> r = p; r.y = neg(p.y); return r;
}

// 2*P for affine P (mdbl-2008-s-1): 3M... in fact 2M + 4S here
template <class F> ZKB_HD XYZZ<F> dbl_affine(const Affine<F>& p) {
  if (p.is_inf()) return XYZZ<F>::inf();
  XYZZ<F> r;
  F u = dbl(p.y);
  F v = sqr(u);
  F w = u * v;
  F s = p.x * v;
  F xx = sqr(p.x);
  F m = dbl(xx) + xx;
  r.x = sqr(m) - dbl(s);
  r.y = m * (s - r.x) - w * p.y;
  r.zz = v;
  r.zzz = w;
  return r;  // y == 0 cannot happen on a prime-order curve (no 2-torsion); would give zz = 0 = identity anyway
}

// 2*P (dbl-2008-s-1)
template <class F> ZKB_HD XYZZ<F> dbl(const XYZZ<F>& p) {
  if (p.is_inf()) return p;
  XYZZ<F> r;
  F u = dbl(p.y);
  F v = sqr(u);
  F w = u * v;
  F s = p.x * v;
  F xx = sqr(p.x);
  F m = dbl(xx) + xx;
  r.x = sqr(m) - dbl(s);
  r.y = m * (s - r.x) - w * p.y;
  r.zz = v * p.zz;
  r.zzz = w * p.zzz;
  return r;
}

// acc + P, P affine (madd-2008-s), complete
template <class F> ZKB_HD XYZZ<F> madd(const XYZZ<F>& a, const Affine<F>& p) {
  if (p.is_inf()) return a;
  if (a.is_inf()) return to_xyzz(p);
  F u2 = p.x * a.zz;
  F s2 = p.y * a.zzz;
  F pp_ = u2 - a.x;
  F rr = s2 - a.y;
  if (pp_.is_zero()) {
    if (rr.is_zero()) return dbl_affine(p);
    return XYZZ<F>::inf();
  }
  XYZZ<F> r;
  F pp = sqr(pp_);
  F ppp = pp_ * pp;
  F q = a.x * pp;
  r.x = sqr(rr) - ppp - dbl(q);
  r.y = mul_sub_mul(rr, q - r.x, a.y, ppp);  // rr (q - x3) - y1 ppp with one reduction
  r.zz = a.zz * pp;
  r.zzz = a.zzz * ppp;
  return r;
}

// a + b (add-2008-s), complete
template <class F> ZKB_HD XYZZ<F> add(const XYZZ<F>& a, const XYZZ<F>& b) {
  if (b.is_inf()) return a;
  if (a.is_inf()) return b;
  F u1 = a.x * b.zz;
  F u2 = b.x * a.zz;
  F s1 = a.y * b.zzz;
  F s2 = b.y * a.zzz;
  F pp_ = u2 - u1;
  F rr = s2 - s1;
  if (pp_.is_zero()) {
    if (rr.is_zero()) return dbl(a);
    return XYZZ<F>::inf();
  }
  XYZZ<F> r;
  F pp = sqr(pp_);
  F ppp = pp_ * pp;
  F q = u1 * pp;
  r.x = sqr(rr) - ppp - dbl(q);
  r.y = rr * (q - r.x) - s1 * ppp;
  r.zz = a.zz * b.zz * pp;
  r.zzz = a.zzz * b.zzz * ppp;
  return r;
}

// out-of-line copies for the latency-bound reduction kernels (one body per kernel instead of one per
// call site: an inlined Fq2 add is ~9k instructions)
#if defined(__CUDACC__)
template <class F> __device__ __noinline__ XYZZ<F> add_ool(const XYZZ<F>& a, const XYZZ<F>& b) { return add(a, b); }
template <class F> __device__ __noinline__ XYZZ<F> dbl_ool(const XYZZ<F>& a) { return dbl(a); }

// ---- one addition on FOUR lanes (a "quad": lanes 4k .. 4k+3 of a warp) ----------------------------------------------
// The bucket hierarchy is a dependent chain of additions executed by a handful of warps: its duration is the latency
// of ONE thread's addition -- 14 field multiplications back to back, each a carry chain through the single CC flag
// (~25 us for an Fq2 addition) -- not the machine's throughput.  The 14 products of add-2008-s form 4 dependency
// levels of <= 4 independent products, so a quad that holds the operands replicated computes one level per step,
// every lane one product, and exchanges the results with shuffles: 4 multiplications deep instead of 14 (doubling:
// 3 instead of 9).  All 32 lanes of the warp must call these together (full-mask shuffles); the special cases
// (identity operands, P + P, P + (-P)) are resolved by selection after the arithmetic, never by early exit.
template <class F>
__device__ __forceinline__ F quad_get(const F& v, int k) {  // the value lane k of this lane's quad holds
  F r;
  const uint32_t* s = reinterpret_cast<const uint32_t*>(&v);
  uint32_t* d = reinterpret_cast<uint32_t*>(&r);
  const int src = (threadIdx.x & 28) | k;
#pragma unroll
  for (int i = 0; i < (int)(sizeof(F) / 4); i++) d[i] = __shfl_sync(0xffffffffu, s[i], src);
  return r;
}
template <class F>
__device__ __forceinline__ F sel4(int q, const F& a0, const F& a1, const F& a2, const F& a3) {
  F r;
  const uint32_t *p0 = reinterpret_cast<const uint32_t*>(&a0), *p1 = reinterpret_cast<const uint32_t*>(&a1),
                 *p2 = reinterpret_cast<const uint32_t*>(&a2), *p3 = reinterpret_cast<const uint32_t*>(&a3);
  uint32_t* d = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(F) / 4); i++) d[i] = q == 0 ? p0[i] : (q == 1 ? p1[i] : (q == 2 ? p2[i] : p3[i]));
  return r;
}
// 2 * p (dbl-2008-s-1) on a quad
template <class F>
__device__ __noinline__ XYZZ<F> qdbl(const XYZZ<F>& p) {
  const int q = threadIdx.x & 3;
  const F u = dbl(p.y);
  F t = sel4(q, u, p.x, u, p.x);
  t = sqr(t);
  const F v = quad_get(t, 0), xx = quad_get(t, 1);
  const F m = dbl(xx) + xx;
  t = sel4(q, u, p.x, m, v) * sel4(q, v, v, m, p.zz);
  const F w = quad_get(t, 0), s = quad_get(t, 1), m2 = quad_get(t, 2);
  XYZZ<F> r;
  r.zz = quad_get(t, 3);
  r.x = m2 - dbl(s);
  t = sel4(q, m, w, w, w) * sel4(q, s - r.x, p.y, p.zzz, p.zzz);
  r.y = quad_get(t, 0) - quad_get(t, 1);
  r.zzz = quad_get(t, 2);
  if (p.is_inf()) return XYZZ<F>::inf();
  return r;
}
// a + b (add-2008-s) on a quad, complete
template <class F>
__device__ __noinline__ XYZZ<F> qadd(const XYZZ<F>& a, const XYZZ<F>& b) {
  const int q = threadIdx.x & 3;
  F t = sel4(q, a.x, b.x, a.y, b.y) * sel4(q, b.zz, a.zz, b.zzz, a.zzz);
  const F u1 = quad_get(t, 0), s1 = quad_get(t, 2);
  const F pp_ = quad_get(t, 1) - u1, rr = quad_get(t, 3) - s1;
  t = sel4(q, pp_, rr, a.zz, a.zzz) * sel4(q, pp_, rr, b.zz, b.zzz);
  const F pp = quad_get(t, 0), rr2 = quad_get(t, 1), zzab = quad_get(t, 2), zzzab = quad_get(t, 3);
  t = sel4(q, pp_, u1, zzab, s1) * pp;
  const F ppp = quad_get(t, 0), qq = quad_get(t, 1), s1pp = quad_get(t, 3);
  XYZZ<F> r;
  r.zz = quad_get(t, 2);
  r.x = rr2 - ppp - dbl(qq);
  t = sel4(q, rr, s1pp, zzzab, zzzab) * sel4(q, qq - r.x, pp_, ppp, ppp);
  r.y = quad_get(t, 0) - quad_get(t, 1);
  r.zzz = quad_get(t, 2);
  const bool ai = a.is_inf(), bi = b.is_inf();
  const bool same_x = !ai && !bi && pp_.is_zero();
  const bool twice = same_x && rr.is_zero();
  if (__any_sync(0xffffffffu, twice)) {  // P + P somewhere in the warp: every lane takes the doubling together
    const XYZZ<F> d = qdbl(a);
    if (twice) r = d;
  }
  if (same_x && !twice) r = XYZZ<F>::inf();
  if (bi) r = a;
  else if (ai) r = b;
  return r;
}
#endif

template <class F> ZKB_HD Affine<F> to_affine(const XYZZ<F>& p) {
  if (p.is_inf()) return Affine<F>::inf();
  // x = X/ZZ, y = Y/ZZZ; one inversion: i = 1/(ZZ*ZZZ)
  F i = inverse(p.zz * p.zzz);
  Affine<F> r;
  r.x = p.x * (i * p.zzz);
  r.y = p.y * (i * p.zz);
  return r;
}

// k*P by MSB-first double-and-add; k is a plain 256-bit integer (8 u32 limbs, canonical Fr residue)
template <class F> ZKB_HD XYZZ<F> scalar_mul(const Affine<F>& p, const uint32_t k[8]) {
  XYZZ<F> acc = XYZZ<F>::inf();
  for (int i = 7; i >= 0; i--) {
    for (int b = 31; b >= 0; b--) {
      acc = dbl(acc);
      if ((k[i] >> b) & 1u) acc = madd(acc, p);
    }
  }
  return acc;
}

}  // namespace zkb
