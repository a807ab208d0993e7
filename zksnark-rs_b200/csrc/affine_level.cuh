// One level of the batched-affine bucket accumulation ("pair tree").
//
// The sorted record array of an MSM (msm_impl.cuh) holds, bucket by bucket, the table entries whose sum is the bucket.
// The XYZZ chain spends 8M + 2S per record.  An AFFINE addition is 1 inversion + 2M + 1S -- and the inversions of B
// independent additions cost one inversion and 3 (B - 1) products together (Montgomery's trick), i.e. 5M + 1S per
// addition once B is large enough to pay for the inversion (ff.cuh: safegcd, ~15 products of the multiplier pipe).
// Independent additions are what a bucket's records are once they are summed as a TREE: level l adds the elements
// 2i, 2i+1 of every bucket (bucket-local indices; an odd last element passes through), so a bucket of k records has
// ceil(k / 2^l) elements after l levels and the counts / offsets of every level follow from the level-0 offsets alone.
// After a few levels the remaining elements (~ 1 / 2^l of the records) go through the XYZZ chain as before.
//
// This header is the level itself, written as a __host__ __device__ function of one WORK ITEM (B consecutive output
// elements) so that the same code runs under g++ (tests/hostcheck) and inside k_affine_level (msm_impl.cuh).
// Replaces, like the rest of the MSM, the per-term double-and-add folds of /root/reference/src/groth16/mod.rs:255-272,
// 279-290 (fr.rs:114-119, 191-223); group addition is exact, so any summation order gives the same affine result.
#pragma once
#include "ec.cuh"

namespace zkb {

// What P + Q needs (both affine, either may be the identity (0, 0)):
//   0: the chord, denominator d = Q.x - P.x != 0          1: the tangent (P == Q), d = 2 P.y != 0
//   2: P is the identity, sum = Q      3: Q is the identity, sum = P      4: Q == -P (or a 2-torsion point doubled): sum = identity
template <class F>
ZKB_HD int pair_kind(const Affine<F>& P, const Affine<F>& Q, F& d) {
  if (P.is_inf()) return 2;
  if (Q.is_inf()) return 3;
  d = Q.x - P.x;
  if (!d.is_zero()) return 0;
  if (P.y == Q.y) {
    d = dbl(P.y);
    return d.is_zero() ? 4 : 1;
  }
  return 4;
}

namespace aff {
static const uint32_t IDX = 0x7fffffffu;  // index bits of a source word; bit 31 = negate (level 0: the record's sign)

ZKB_HD void prefetch_l2(const void* p) {
#if defined(__CUDA_ARCH__)
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
  (void)p;
#endif
}
template <class F>
ZKB_HD void prefetch_point(const Affine<F>* p) {
  prefetch_l2(p);
  if (sizeof(Affine<F>) > 64) prefetch_l2(reinterpret_cast<const char*>(p) + 64);
}
// 16-byte loads; 64-byte table entries (G1) with the L2::64B prefetch size so that a gather fills half a line, not a
// whole one (msm_impl.cuh: ld_table_entry, profiles/r02_l2_fetch_granularity.txt)
template <class T, bool HINT64>
ZKB_HD T load_words(const void* p) {
#if defined(__CUDA_ARCH__)
  if (HINT64) {
    T r;
    uint32_t* w = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
    for (int k = 0; k < (int)(sizeof(T) / 16); k++)
      asm volatile("ld.global.L2::64B.v4.u32 {%0, %1, %2, %3}, [%4];"
                   : "=r"(w[4 * k]), "=r"(w[4 * k + 1]), "=r"(w[4 * k + 2]), "=r"(w[4 * k + 3])
                   : "l"(reinterpret_cast<const char*>(p) + 16 * k));
    return r;
  }
#endif
  return *reinterpret_cast<const T*>(p);
}
template <class F, bool HINT64>
ZKB_HD Affine<F> load_point(const Affine<F>* pts, uint32_t word) {
  Affine<F> P = load_words<Affine<F>, HINT64>(pts + (word & IDX));
  if (word >> 31) P = neg(P);
  return P;
}
}  // namespace aff

// Work item `item` of one level: the output elements [item * B, item * B + B) of the level's output array.
//   pts / sorted   FIRST: the window table and the sorted records (entry index | sign << 31); else: the previous level's
//                  output array (sorted unused)
//   offs_in/_out   bucket offsets (nbk + 1 each) of the input and output arrays; out counts = ceil(in counts / 2)
//   prefix         this thread's B-element scratch for the running products, element j at prefix[j * pstride]
// PF: how many pairs ahead the operands are pulled into the L2 (device only; both passes)
template <class F, int B, bool FIRST, int PF = 4>
ZKB_HD void affine_level_item(uint32_t item, const Affine<F>* __restrict__ pts, const uint32_t* __restrict__ sorted,
                              const uint32_t* __restrict__ offs_in, const uint32_t* __restrict__ offs_out, uint32_t nbk,
                              Affine<F>* __restrict__ out, F* __restrict__ prefix, size_t pstride) {
  static_assert(B >= 1 && B <= 32, "the pair mask is one 32-bit word");
  // gathers of 64-byte table entries fill half an L2 line; the level arrays are read in order (whole lines wanted)
  constexpr bool H = FIRST && sizeof(Affine<F>) == 64;
  const uint32_t total = offs_out[nbk];
  const uint32_t start = item * (uint32_t)B;
  if (start >= total) return;
  const uint32_t cnt = total - start < (uint32_t)B ? total - start : (uint32_t)B;
  // g: offs_out[g] <= start < offs_out[g + 1]
  uint32_t lo = 0, hi = nbk;
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (offs_out[mid] <= start) lo = mid + 1; else hi = mid;
  }
  uint32_t g = lo - 1;
  uint32_t o_lo = offs_out[g], o_hi = offs_out[g + 1], i_lo = offs_in[g], i_hi = offs_in[g + 1];
  // plan: source words of every output element (s1 only where the element is a pair)
  uint32_t s0[B], s1[B], pairmask = 0;
  for (uint32_t j = 0; j < cnt; j++) {
    const uint32_t p = start + j;
    while (p >= o_hi) {  // next non-empty bucket
      g++;
      o_lo = o_hi; o_hi = offs_out[g + 1];
      i_lo = offs_in[g]; i_hi = offs_in[g + 1];
    }
    const uint32_t i0 = i_lo + 2 * (p - o_lo);
    const bool pair = i0 + 1 < i_hi;
    if (FIRST) {
      s0[j] = sorted[i0];
      s1[j] = pair ? sorted[i0 + 1] : 0;
    } else {
      s0[j] = i0;
      s1[j] = i0 + 1;
    }
    if (pair) {
      pairmask |= 1u << j;
      if (j < (uint32_t)PF) { aff::prefetch_l2(pts + (s0[j] & aff::IDX)); aff::prefetch_l2(pts + (s1[j] & aff::IDX)); }
    }
  }
  // forward: running product of the denominators
  F acc = F::one();
  for (uint32_t j = 0; j < cnt; j++) {
    if (j + PF < cnt && ((pairmask >> (j + PF)) & 1u)) {
      aff::prefetch_l2(pts + (s0[j + PF] & aff::IDX));
      aff::prefetch_l2(pts + (s1[j + PF] & aff::IDX));
    }
    if (!((pairmask >> j) & 1u)) continue;
    const Affine<F>* a0 = pts + (s0[j] & aff::IDX);
    const Affine<F>* a1 = pts + (s1[j] & aff::IDX);
    const F x0 = aff::load_words<F, H>(&a0->x), x1 = aff::load_words<F, H>(&a1->x);
    F d = x1 - x0;
    if (d.is_zero() || x0.is_zero() || x1.is_zero()) {  // rare: identity operand, P + P, P - P
      const Affine<F> P0 = aff::load_point<F, H>(pts, s0[j]), P1 = aff::load_point<F, H>(pts, s1[j]);
      if (pair_kind(P0, P1, d) >= 2) continue;  // no denominator
    }
    prefix[j * pstride] = acc;
    acc = acc * d;
  }
  for (uint32_t q = 0; q < (uint32_t)PF && q < cnt; q++) {  // the backward pass starts at the end
    const uint32_t j = cnt - 1 - q;
    aff::prefetch_point(pts + (s0[j] & aff::IDX));
    if ((pairmask >> j) & 1u) aff::prefetch_point(pts + (s1[j] & aff::IDX));
  }
  F inv = inverse(acc);
  // backward: 1 / d_j = inv * prefix_j; the sums
  for (uint32_t j = cnt; j-- > 0;) {
    if (j >= (uint32_t)PF) {
      aff::prefetch_point(pts + (s0[j - PF] & aff::IDX));
      if ((pairmask >> (j - PF)) & 1u) aff::prefetch_point(pts + (s1[j - PF] & aff::IDX));
    }
    const Affine<F> P0 = aff::load_point<F, H>(pts, s0[j]);
    if (!((pairmask >> j) & 1u)) {  // odd last element of its bucket
      out[start + j] = P0;
      continue;
    }
    const Affine<F> P1 = aff::load_point<F, H>(pts, s1[j]);
    F d;
    const int kind = pair_kind(P0, P1, d);
    Affine<F> R;
    if (kind >= 2) {
      R = kind == 2 ? P1 : (kind == 3 ? P0 : Affine<F>::inf());
    } else {
      const F dinv = inv * prefix[j * pstride];
      inv = inv * d;
      F num;
      if (kind == 0) {
        num = P1.y - P0.y;
      } else {
        const F xx = sqr(P0.x);
        num = dbl(xx) + xx;
      }
      const F lam = num * dinv;
      R.x = sqr(lam) - P0.x - P1.x;
      R.y = lam * (P0.x - R.x) - P0.y;
    }
    out[start + j] = R;
  }
}

// offsets of level l >= 1 from the level-0 bucket sizes: ceil(ceil(k / 2) / 2) = ceil(k / 4), ...
ZKB_HD uint32_t affine_level_count(uint32_t k, int l) { return (uint32_t)(((uint64_t)k + ((1u << l) - 1u)) >> l); }

}  // namespace zkb
