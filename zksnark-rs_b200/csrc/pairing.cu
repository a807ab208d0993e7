// groth16::verify (/root/reference/src/groth16/mod.rs:299-320) and `EllipticEncryptable::pairing`
// (fr.rs:120-122) on the device.
//
//   e(alpha1, beta2) * e(sum_term, gamma2) * e(C, delta2) == e(A, B)
//     <=>  FE( ML(-A, B) * ML(alpha1, beta2) * ML(sum_term, gamma2) * ML(C, delta2) ) == 1
//
// with sum_term = sum_j a_j * sum_gamma[j] over zip(sum_gamma, [1] ++ inputs) (the shorter side ends
// the sum, like the reference's `zip`).  One thread per Miller loop (the four loops of one proof sit
// in one warp and run in lock step: same loop bits), one thread per final exponentiation; a batch of
// proofs fills the machine with independent verifications.  This is latency-class work (O(1) per
// proof), not one of the two hot kernels.
#define ZKB_FP_OOL
#define ZKB_FQ2_OOL
#include "common.cuh"
#include "pairing.cuh"

namespace zkb {

__device__ __forceinline__ Fq ld_canon(const uint32_t* p) {
  Fq c;
#pragma unroll
  for (int i = 0; i < 8; i++) c.v[i] = p[i];
  return to_mont(c);
}
__device__ __forceinline__ bool below_q(const uint32_t* p) {  // canonical residue?
  const Fq m = Fq::modulus();
  for (int i = 7; i >= 0; i--) {
    if (p[i] < m.v[i]) return true;
    if (p[i] > m.v[i]) return false;
  }
  return false;
}
__device__ __forceinline__ Fq fq_raw(const uint32_t k[8]) { Fq r; for (int i = 0; i < 8; i++) r.v[i] = k[i]; return r; }

// y^2 == x^3 + b or identity; false also for non-canonical coordinates
__device__ bool g1_load_checked(const uint32_t* p, G1Affine* out) {
  bool ok = below_q(p) && below_q(p + 8);
  out->x = ld_canon(p);
  out->y = ld_canon(p + 8);
  if (out->is_inf()) return ok;
  const uint32_t b[8] = ZKB_G1_B;
  return ok && sqr(out->y) == sqr(out->x) * out->x + fq_raw(b);
}
__device__ bool g2_load_checked(const uint32_t* p, G2Affine* out) {
  bool ok = below_q(p) && below_q(p + 8) && below_q(p + 16) && below_q(p + 24);
  out->x.c0 = ld_canon(p); out->x.c1 = ld_canon(p + 8);
  out->y.c0 = ld_canon(p + 16); out->y.c1 = ld_canon(p + 24);
  if (out->is_inf()) return ok;
  const uint32_t b0[8] = ZKB_G2_B0, b1[8] = ZKB_G2_B1;
  Fq2 b; b.c0 = fq_raw(b0); b.c1 = fq_raw(b1);
  return ok && sqr(out->y) == sqr(out->x) * out->x + b;
}

// thread (i, j): a_j * sum_gamma[j] of proof i, a_0 = 1 (mod.rs:312-316)
__global__ void k_verify_terms(const G1Affine* __restrict__ sum_gamma, const uint32_t* __restrict__ inputs, size_t n_inputs,
                               size_t terms, size_t count, G1XYZZ* __restrict__ out) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= terms * count) return;
  size_t i = t / terms, j = t % terms;
  if (j == 0) { out[t] = to_xyzz(sum_gamma[0]); return; }
  uint32_t k[8];
  for (int l = 0; l < 8; l++) k[l] = inputs[(i * n_inputs + (j - 1)) * 8 + l];
  out[t] = scalar_mul(sum_gamma[j], k);
}

// thread i: the four pairs of proof i -> P[4i..], Q[4i..]; bad[i] = 1 if a proof point is off its curve
__global__ void k_verify_pairs(const G1Affine* __restrict__ alpha1, const G2Affine* __restrict__ beta2,
                               const G2Affine* __restrict__ gamma2, const G2Affine* __restrict__ delta2,
                               const uint32_t* __restrict__ proofs, const G1XYZZ* __restrict__ term, size_t terms, size_t count,
                               G1Affine* __restrict__ P, G2Affine* __restrict__ Q, int* __restrict__ bad) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const uint32_t* pr = proofs + i * 64;  // a: 16 u32 | b: 32 u32 | c: 16 u32
  G1Affine a, c;
  G2Affine b;
  bool ok = g1_load_checked(pr, &a);
  ok = g2_load_checked(pr + 16, &b) && ok;
  ok = g1_load_checked(pr + 48, &c) && ok;
  bad[i] = ok ? 0 : 1;
  G1XYZZ s = G1XYZZ::inf();
  for (size_t j = 0; j < terms; j++) s = add(s, term[i * terms + j]);
  P[4 * i + 0] = neg(a);          Q[4 * i + 0] = b;
  P[4 * i + 1] = *alpha1;         Q[4 * i + 1] = *beta2;
  P[4 * i + 2] = to_affine(s);    Q[4 * i + 2] = *gamma2;
  P[4 * i + 3] = c;               Q[4 * i + 3] = *delta2;
}

// canonical host pairs -> Montgomery (zkb_pairing)
__global__ void k_pairs_load(const uint32_t* __restrict__ g1s, const uint32_t* __restrict__ g2s, size_t n, G1Affine* __restrict__ P,
                             G2Affine* __restrict__ Q, int* __restrict__ bad) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  bool ok = g1_load_checked(g1s + i * 16, &P[i]);
  ok = g2_load_checked(g2s + i * 32, &Q[i]) && ok;
  if (!ok) atomicOr(bad, 1);
}

// proof.b / pairing arguments must lie in the order-r subgroup of the twist (cofactor != 1): [r]Q == O by double-and-add.
// The reference's G2Local values are always bn-constructed subgroup elements; raw coordinates over the ABI are not.
// One thread per point Q[i * stride]; per_item: bad[i] |= 2, else *bad |= 2.  Runs beside the Miller loops.
__global__ void __launch_bounds__(32) k_g2_subgroup(const G2Affine* __restrict__ Q, size_t stride, size_t count, int* __restrict__ bad,
                                                    int per_item) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const G2Affine q = Q[i * stride];
  if (q.is_inf()) return;
  const Fr m = Fr::modulus();
  if (!scalar_mul(q, m.v).is_inf()) atomicOr(per_item ? bad + i : bad, 2);
}

__global__ void __launch_bounds__(32) k_miller(const G1Affine* __restrict__ P, const G2Affine* __restrict__ Q, size_t n,
                                               Fq12* __restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = miller_loop(P[i], Q[i]);
}

// thread i: product of `per` Miller values, final exponentiation; gt (optional): canonical residues,
// 12 Fq per element (c0, c1 of the w^i coefficient, i = 0..5); ok (optional): result == 1 and !bad
__global__ void __launch_bounds__(32) k_final_exp(const Fq12* __restrict__ ml, size_t per, size_t count, const int* __restrict__ bad,
                                                  uint32_t* __restrict__ gt, int* __restrict__ ok) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  Fq12 f = Fq12::one();
  for (size_t j = 0; j < per; j++) f = f * ml[i * per + j];
  f = final_exponentiation(f);
  if (gt)
    for (int k = 0; k < 6; k++) {
      Fq c0 = from_mont(f.w(k).c0), c1 = from_mont(f.w(k).c1);
      for (int l = 0; l < 8; l++) { gt[(i * 12 + 2 * k) * 8 + l] = c0.v[l]; gt[(i * 12 + 2 * k + 1) * 8 + l] = c1.v[l]; }
    }
  if (ok) ok[i] = (f == Fq12::one()) && !(bad && bad[i]);
}

}  // namespace zkb

using namespace zkb;

extern "C" {

int zkb_verify_batch(zkb_ctx* ctx, const zkb_crs* crs, const uint64_t* inputs, size_t n_inputs, const zkb_proof* proofs,
                     size_t count, int* ok) {
  if (!ctx || !crs || (!inputs && n_inputs) || (!proofs && count) || (!ok && count))
    return set_err(ctx, ZKB_ERR_ARG, "zkb_verify: NULL argument");
  if (!count) return ZKB_OK;
  ZKB_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const size_t terms = crs->n_sum_gamma < n_inputs + 1 ? crs->n_sum_gamma : n_inputs + 1;  // zip truncation
  const size_t in_bytes = count * n_inputs * 32, pr_bytes = count * sizeof(zkb_proof);
  const size_t tm_bytes = (count * terms + 1) * sizeof(G1XYZZ);
  const size_t bytes = in_bytes + pr_bytes + tm_bytes + 4 * count * (sizeof(G1Affine) + sizeof(G2Affine) + sizeof(Fq12)) +
                       2 * count * sizeof(int) + 256;
  void* p;
  ZKB_TRY(scratch_get(ctx, 8, bytes, &p));
  char* q = (char*)p;
  Fq12* ml = (Fq12*)q;          q += 4 * count * sizeof(Fq12);
  G2Affine* Q = (G2Affine*)q;   q += 4 * count * sizeof(G2Affine);
  G1XYZZ* term = (G1XYZZ*)q;    q += tm_bytes;
  G1Affine* P = (G1Affine*)q;   q += 4 * count * sizeof(G1Affine);
  uint32_t* d_pr = (uint32_t*)q; q += pr_bytes;
  uint32_t* d_in = (uint32_t*)q; q += in_bytes;
  int* d_bad = (int*)q;         q += count * sizeof(int);
  int* d_ok = (int*)q;
  ZKB_CUDA(ctx, cudaMemcpyAsync(d_pr, proofs, pr_bytes, cudaMemcpyHostToDevice, st));
  if (in_bytes) ZKB_CUDA(ctx, cudaMemcpyAsync(d_in, inputs, in_bytes, cudaMemcpyHostToDevice, st));
  if (terms) ZKB_LAUNCH(ctx, k_verify_terms, cdiv(terms * count, 64), 64, 0, st, crs->sum_gamma, d_in, n_inputs, terms, count, term);
  ZKB_LAUNCH(ctx, k_verify_pairs, cdiv(count, 32), 32, 0, st, crs->g1 + crs->off_fixed(), crs->g2 + crs->nxi(), crs->gamma2,
             crs->g2 + crs->nxi() + 1, d_pr, term, terms, count, P, Q, d_bad);
  // subgroup membership of proof.b on the second stream, beside the Miller loops
  cudaEvent_t ev_pairs = ctx->lanes[0].ev[0], ev_sub = ctx->lanes[0].ev[1];
  ZKB_CUDA(ctx, cudaEventRecord(ev_pairs, st));
  ZKB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream2, ev_pairs, 0));
  ZKB_LAUNCH(ctx, k_g2_subgroup, cdiv(count, 32), 32, 0, ctx->stream2, Q, (size_t)4, count, d_bad, 1);
  ZKB_CUDA(ctx, cudaEventRecord(ev_sub, ctx->stream2));
  ZKB_LAUNCH(ctx, k_miller, cdiv(4 * count, 32), 32, 0, st, P, Q, 4 * count, ml);
  ZKB_CUDA(ctx, cudaStreamWaitEvent(st, ev_sub, 0));
  ZKB_LAUNCH(ctx, k_final_exp, cdiv(count, 32), 32, 0, st, ml, (size_t)4, count, d_bad, (uint32_t*)nullptr, d_ok);
  ZKB_CUDA(ctx, cudaMemcpyAsync(ok, d_ok, count * sizeof(int), cudaMemcpyDeviceToHost, st));
  ZKB_CUDA(ctx, cudaStreamSynchronize(st));
  return ZKB_OK;
}

int zkb_verify(zkb_ctx* ctx, const zkb_crs* crs, const uint64_t* inputs, size_t n_inputs, const zkb_proof* proof, int* ok) {
  return zkb_verify_batch(ctx, crs, inputs, n_inputs, proof, 1, ok);
}

int zkb_pairing(zkb_ctx* ctx, const uint64_t* g1s, const uint64_t* g2s, size_t n, uint64_t* gt) {
  if (!ctx || (!g1s && n) || (!g2s && n) || !gt) return set_err(ctx, ZKB_ERR_ARG, "zkb_pairing: NULL argument");
  ZKB_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const size_t bytes = (n + 1) * (sizeof(Fq12) + sizeof(G2Affine) + sizeof(G1Affine) + 64 + 128) + 48 * 8 + 256;
  void* p;
  ZKB_TRY(scratch_get(ctx, 8, bytes, &p));
  char* q = (char*)p;
  Fq12* ml = (Fq12*)q;           q += (n + 1) * sizeof(Fq12);
  G2Affine* Q = (G2Affine*)q;    q += (n + 1) * sizeof(G2Affine);
  G1Affine* P = (G1Affine*)q;    q += (n + 1) * sizeof(G1Affine);
  uint32_t* d_g2 = (uint32_t*)q; q += (n + 1) * 128;
  uint32_t* d_g1 = (uint32_t*)q; q += (n + 1) * 64;
  uint32_t* d_gt = (uint32_t*)q; q += 48 * 8;
  int* d_bad = (int*)q;
  ZKB_CUDA(ctx, cudaMemsetAsync(d_bad, 0, sizeof(int), st));
  if (n) {
    ZKB_CUDA(ctx, cudaMemcpyAsync(d_g1, g1s, n * 64, cudaMemcpyHostToDevice, st));
    ZKB_CUDA(ctx, cudaMemcpyAsync(d_g2, g2s, n * 128, cudaMemcpyHostToDevice, st));
    ZKB_LAUNCH(ctx, k_pairs_load, cdiv(n, 32), 32, 0, st, d_g1, d_g2, n, P, Q, d_bad);
    cudaEvent_t ev_pairs = ctx->lanes[0].ev[0], ev_sub = ctx->lanes[0].ev[1];
    ZKB_CUDA(ctx, cudaEventRecord(ev_pairs, st));
    ZKB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream2, ev_pairs, 0));
    ZKB_LAUNCH(ctx, k_g2_subgroup, cdiv(n, 32), 32, 0, ctx->stream2, Q, (size_t)1, n, d_bad, 0);
    ZKB_CUDA(ctx, cudaEventRecord(ev_sub, ctx->stream2));
    ZKB_LAUNCH(ctx, k_miller, cdiv(n, 32), 32, 0, st, P, Q, n, ml);
    ZKB_CUDA(ctx, cudaStreamWaitEvent(st, ev_sub, 0));
  }
  ZKB_LAUNCH(ctx, k_final_exp, 1, 32, 0, st, ml, n, (size_t)1, (const int*)nullptr, d_gt, (int*)nullptr);
  int bad = 0;
  ZKB_CUDA(ctx, cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, st));
  ZKB_CUDA(ctx, cudaMemcpyAsync(gt, d_gt, 48 * 8, cudaMemcpyDeviceToHost, st));
  ZKB_CUDA(ctx, cudaStreamSynchronize(st));
  if (bad & 1) return set_err(ctx, ZKB_ERR_ARG, "zkb_pairing: a point is not on its curve (or a coordinate is not a canonical residue)");
  if (bad & 2) return set_err(ctx, ZKB_ERR_ARG, "zkb_pairing: a G2 point is not in the order-r subgroup");
  return ZKB_OK;
}

}  // extern "C"
