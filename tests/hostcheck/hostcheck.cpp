// TEST INFRASTRUCTURE: compiles the product's __host__ __device__ field / curve formulas
// (zksnark-rs_b200/csrc/ff.cuh, ec.cuh) with g++ so their algebra can be checked against Oracle A
// on a machine without a GPU.  Never linked into libzkb200.so and never used by the product path.
#include <cstring>
#include <vector>
#include "pairing.cuh"
#include "affine_level.cuh"
using namespace zkb;

template <class F> static F ld(const uint64_t* p) { F r; memcpy(r.v, p, 32); return to_mont(r); }
template <class F> static void st(uint64_t* p, const F& m) { F c = from_mont(m); memcpy(p, c.v, 32); }
static Fq2 ld2(const uint64_t* p) { Fq2 r; r.c0 = ld<Fq>(p); r.c1 = ld<Fq>(p + 4); return r; }
static void st2(uint64_t* p, const Fq2& a) { st(p, a.c0); st(p + 4, a.c1); }
static G1Affine ldg1(const uint64_t* p) { G1Affine r; r.x = ld<Fq>(p); r.y = ld<Fq>(p + 4); return r; }
static void stg1(uint64_t* p, const G1Affine& a) { st(p, a.x); st(p + 4, a.y); }
static G2Affine ldg2(const uint64_t* p) { G2Affine r; r.x = ld2(p); r.y = ld2(p + 8); return r; }
static void stg2(uint64_t* p, const G2Affine& a) { st2(p, a.x); st2(p + 8, a.y); }

// The batched-affine pair tree (affine_level.cuh) over one sorted record array, run item by item as the kernel does
// (one prefix scratch of B elements, stride 1).  group 1 / 2; table: n_tab affine points (canonical); sorted: n_rec records
// (entry | sign << 31) ordered by bucket; offs0: nbk + 1 offsets.  Runs `levels` levels with B = batch (4, 16 or 32) and
// writes the last level's offsets (nbk + 1) and its elements (canonical, identity = zeros); returns the element count.
template <class F, int B>
static uint32_t pair_tree(const std::vector<Affine<F>>& tab, const uint32_t* sorted, const uint32_t* offs0, uint32_t nbk, int levels,
                          std::vector<uint32_t>& offs_out, std::vector<Affine<F>>& elems) {
  std::vector<uint32_t> offs_in(offs0, offs0 + nbk + 1);
  std::vector<Affine<F>> in;
  F prefix[B];
  for (int l = 1; l <= levels; l++) {
    std::vector<uint32_t> o(nbk + 1);
    uint32_t run = 0;
    for (uint32_t b = 0; b < nbk; b++) { o[b] = run; run += affine_level_count(offs0[b + 1] - offs0[b], l); }
    o[nbk] = run;
    std::vector<Affine<F>> out(run);
    const uint32_t items = (run + B - 1) / B;
    for (uint32_t it = 0; it < items + 2; it++) {  // two items past the end: must be no-ops
      if (l == 1) affine_level_item<F, B, true>(it, tab.data(), sorted, offs_in.data(), o.data(), nbk, out.data(), prefix, 1);
      else affine_level_item<F, B, false>(it, in.data(), nullptr, offs_in.data(), o.data(), nbk, out.data(), prefix, 1);
    }
    in.swap(out);
    offs_in.swap(o);
  }
  offs_out = offs_in;
  elems = in;
  return (uint32_t)in.size();
}

extern "C" {
// n inverses at once, both ways: out = inverse(a) (safegcd), out2 = inverse_fermat(a); field: 0 Fr 1 Fq
void hc_inverse_many(int field, int n, const uint64_t* a, uint64_t* out, uint64_t* out2) {
  for (int i = 0; i < n; i++) {
    if (field == 0) { Fr x = ld<Fr>(a + 4 * i); st(out + 4 * i, inverse(x)); st(out2 + 4 * i, inverse_fermat(x)); }
    else { Fq x = ld<Fq>(a + 4 * i); st(out + 4 * i, inverse(x)); st(out2 + 4 * i, inverse_fermat(x)); }
  }
}
// op: 0 mul 1 add 2 sub 3 inverse(a) ; field: 0 Fr 1 Fq
void hc_field(int field, int op, const uint64_t* a, const uint64_t* b, uint64_t* out) {
  if (field == 0) {
    Fr x = ld<Fr>(a), y = ld<Fr>(b), r;
    r = op == 0 ? x * y : op == 1 ? x + y : op == 2 ? x - y : inverse(x);
    st(out, r);
  } else {
    Fq x = ld<Fq>(a), y = ld<Fq>(b), r;
    r = op == 0 ? x * y : op == 1 ? x + y : op == 2 ? x - y : inverse(x);
    st(out, r);
  }
}
void hc_fq2(int op, const uint64_t* a, const uint64_t* b, uint64_t* out) {
  Fq2 x = ld2(a), y = ld2(b);
  Fq2 r = op == 0 ? x * y : op == 1 ? x + y : op == 2 ? x - y : op == 3 ? inverse(x) : sqr(x);
  st2(out, r);
}
// out = a + b (both affine, through XYZZ: to_xyzz(a) madd b), out2 = to_affine(add(xyzz(a), xyzz(b)))
void hc_g1_add(const uint64_t* a, const uint64_t* b, uint64_t* out, uint64_t* out2) {
  G1Affine A = ldg1(a), B = ldg1(b);
  stg1(out, to_affine(madd(to_xyzz(A), B)));
  // exercise the general add with non-trivial zz on both sides: (2A - A) + (2B - B)
  G1XYZZ a2 = add(dbl(to_xyzz(A)), neg(to_xyzz(A)));
  G1XYZZ b2 = add(dbl_affine(B), to_xyzz(neg(B)));
  stg1(out2, to_affine(add(a2, b2)));
}
void hc_g1_mul(const uint64_t* a, const uint64_t* k, uint64_t* out) {
  uint32_t kk[8]; memcpy(kk, k, 32);
  stg1(out, to_affine(scalar_mul(ldg1(a), kk)));
}
void hc_g2_add(const uint64_t* a, const uint64_t* b, uint64_t* out, uint64_t* out2) {
  G2Affine A = ldg2(a), B = ldg2(b);
  stg2(out, to_affine(madd(to_xyzz(A), B)));
  G2XYZZ a2 = add(dbl(to_xyzz(A)), neg(to_xyzz(A)));
  G2XYZZ b2 = add(dbl_affine(B), to_xyzz(neg(B)));
  stg2(out2, to_affine(add(a2, b2)));
}
void hc_g2_mul(const uint64_t* a, const uint64_t* k, uint64_t* out) {
  uint32_t kk[8]; memcpy(kk, k, 32);
  stg2(out, to_affine(scalar_mul(ldg2(a), kk)));
}
// GT element as 12 Fq residues: the Fq2 coefficient (c0, c1) of w^i, i = 0..5.  stage 0: product of the n Miller
// loops only, 1: its final exponentiation; +2: the plain versions (affine steps, 761-bit exponent)
static void st12(uint64_t* out, const Fq12& f) { for (int i = 0; i < 6; i++) st2(out + 8 * i, f.w(i)); }
void hc_pairing(int stage, int n, const uint64_t* g1s, const uint64_t* g2s, uint64_t* out) {
  Fq12 f = Fq12::one();
  for (int i = 0; i < n; i++)
    f = f * ((stage & 2) ? miller_loop_affine(ldg1(g1s + 8 * i), ldg2(g2s + 16 * i)) : miller_loop(ldg1(g1s + 8 * i), ldg2(g2s + 16 * i)));
  if (stage & 1) f = (stage & 2) ? final_exponentiation_plain(f) : final_exponentiation(f);
  st12(out, f);
}
// Fq12 arithmetic on its own: op 0 a*b, 1 a^2, 2 1/a, 3 a^(q^2), 4 a^(q^6), 5 a^q, 6 a^u (a in the
// cyclotomic subgroup), 7 a * sparse line (b_0, b_1, b_3), 8 cyclotomic square, 9 the easy part a^((q^6-1)(q^2+1))
void hc_fq12(int op, const uint64_t* a, const uint64_t* b, uint64_t* out) {
  Fq12 x, y;
  for (int i = 0; i < 6; i++) { x.w(i) = ld2(a + 8 * i); y.w(i) = ld2(b + 8 * i); }
  if (op == 7) { mul_by_line(x, y.w(0), y.w(1), y.w(3)); st12(out, x); return; }
  if (op == 8) { st12(out, cyclotomic_sqr(x)); return; }
  if (op == 9) { Fq12 t = conj(x) * inverse(x); st12(out, frobenius2(t) * t); return; }
  st12(out, op == 0 ? x * y : op == 1 ? sqr(x) : op == 2 ? inverse(x) : op == 3 ? frobenius2(x) : op == 4 ? conj(x) : op == 5 ? frobenius(x) : pow_u(x));
}

uint32_t hc_pair_tree(int group, int batch, uint32_t n_tab, const uint64_t* table, uint32_t nbk, const uint32_t* offs0,
                      const uint32_t* sorted, int levels, uint32_t* offs_out, uint64_t* elems_out) {
  std::vector<uint32_t> oo;
  uint32_t n = 0;
  if (group == 1) {
    std::vector<G1Affine> tab(n_tab), el;
    for (uint32_t i = 0; i < n_tab; i++) {
      const uint64_t* p = table + 8 * i;
      bool zero = true;
      for (int k = 0; k < 8; k++) zero = zero && p[k] == 0;
      tab[i] = zero ? G1Affine::inf() : ldg1(p);
    }
    n = batch == 4 ? pair_tree<Fq, 4>(tab, sorted, offs0, nbk, levels, oo, el)
        : batch == 16 ? pair_tree<Fq, 16>(tab, sorted, offs0, nbk, levels, oo, el) : pair_tree<Fq, 32>(tab, sorted, offs0, nbk, levels, oo, el);
    for (uint32_t i = 0; i < n; i++) stg1(elems_out + 8 * i, el[i]);
  } else {
    std::vector<G2Affine> tab(n_tab), el;
    for (uint32_t i = 0; i < n_tab; i++) {
      const uint64_t* p = table + 16 * i;
      bool zero = true;
      for (int k = 0; k < 16; k++) zero = zero && p[k] == 0;
      tab[i] = zero ? G2Affine::inf() : ldg2(p);
    }
    n = batch == 4 ? pair_tree<Fq2, 4>(tab, sorted, offs0, nbk, levels, oo, el)
        : batch == 16 ? pair_tree<Fq2, 16>(tab, sorted, offs0, nbk, levels, oo, el) : pair_tree<Fq2, 32>(tab, sorted, offs0, nbk, levels, oo, el);
    for (uint32_t i = 0; i < n; i++) stg2(elems_out + 16 * i, el[i]);
  }
  for (uint32_t b = 0; b <= nbk; b++) offs_out[b] = oo[b];
  return n;
}
}
