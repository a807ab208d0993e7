// groth16::prove on the device (include/zkb200.h: zkb_qap_*, zkb_prove*, zkb_qap_h).
//
// Reference path: /root/reference/src/groth16/mod.rs:213-296.  Stages here:
//   1. witness upload, canonical -> Montgomery
//   2. k_matvec          A_k = sum_i a_i u_i(w^k), B_k likewise (sparse rows; replaces the dense
//                        weighted sums mod.rs:233-253) and AB_k = A_k * B_k
//   3. inverse NTTs      u_sum, v_sum coefficients (unique degree<n interpolants)
//   4. coset NTTs + pointwise product + inverse NTT: with g = omega_2n (g^n = -1),
//                        c = iNTT(A.B) = p_lo + p_hi and d_k g^-k = p_lo - p_hi, p = u_sum * v_sum;
//                        t = x^n - 1, so the quotient (mod.rs:277, field/mod.rs:428-469) is
//                        h = p_hi = (c - d g^-k) / 2.  w_sum has degree < n = deg t and cannot
//                        influence the quotient (the remainder is discarded, coefficient_poly.rs:155).
//   5. k_msm_scalars     the MSM scalar vectors, canonical.  The reference's
//                          A = a_g1 + alpha1 + r delta1                                  (mod.rs:274)
//                          B = b_g2 + beta2 + s delta2                                   (mod.rs:275)
//                          C = c_h + c_w + s A + r (beta1 + b_g1 + s delta1) - rs delta1 (mod.rs:279-293)
//                        expand (group law, exact) to three MSMs over the CRS tables:
//                          A = <[u_sum | 1, 0, r], [xi1 | alpha1, beta1, delta1]>
//                          B = <[v_sum | 1, s],    [xi2 | beta2, delta2]>
//                          C = <[s u_sum + r v_sum | s, r, rs | h | a_(l+1..)], [xi1 | alpha1, beta1, delta1 | xi_t | sum_delta]>
//                        so no scalar multiplication of a fresh point is left.
//   6. MSMs              A and C as two jobs of one G1 call; B (G2) on a second stream
//   7. k_finish          three affine normalisations -> Proof{a, b, c}
#include <string.h>
#include <algorithm>
#include "common.cuh"

namespace zkb {

// ------------------------------------------------------------------------------------------------
// kernels of the polynomial stage
__global__ void k_matvec(const uint32_t* __restrict__ gptr_u, const uint32_t* __restrict__ wire_u, const Fr* __restrict__ coef_u,
                         const uint32_t* __restrict__ gptr_v, const uint32_t* __restrict__ wire_v, const Fr* __restrict__ coef_v,
                         const Fr* __restrict__ a, size_t m, size_t k0, size_t kstride, size_t count, Fr* __restrict__ A,
                         Fr* __restrict__ B, Fr* __restrict__ AB) {
  // output i <-> gate k = k0 + kstride * i (one GPU: all gates; a sharded proof: the gates rank, rank + G, ...)
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const size_t k = k0 + kstride * i;
  Fr su = Fr::zero(), sv = Fr::zero();
  for (uint32_t p = gptr_u[k], e = gptr_u[k + 1]; p < e; p++) {
    uint32_t w = wire_u[p];
    if (w < m) su = su + coef_u[p] * a[w];  // wires beyond the witness count as zero (zip truncation)
  }
  for (uint32_t p = gptr_v[k], e = gptr_v[k + 1]; p < e; p++) {
    uint32_t w = wire_v[p];
    if (w < m) sv = sv + coef_v[p] * a[w];
  }
  A[i] = su;
  B[i] = sv;
  AB[i] = su * sv;
}

__global__ void k_mul2(Fr* __restrict__ o1, const Fr* __restrict__ a1, Fr* __restrict__ o2, const Fr* __restrict__ a2,
                       const Fr* __restrict__ tab, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr t = tab[i];
  o1[i] = a1[i] * t;
  o2[i] = a2[i] * t;
}

// h_br[i] = c_br[i] * inv2n - d_br[i] * Q[i]
__global__ void k_hfinal(Fr* __restrict__ h, const Fr* __restrict__ c, const Fr* __restrict__ d, const Fr* __restrict__ Q,
                         Fr inv2n, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  h[i] = c[i] * inv2n - d[i] * Q[i];
}

// coset tables in bit-reversed position order
__global__ void k_coset_tables(Fr* P, Fr* Q, Fr g, Fr ginv, Fr invn, Fr inv2n, uint32_t log_n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t n = (size_t)1 << log_n;
  if (i >= n) return;
  uint64_t e = log_n ? (uint64_t)(__brevll((unsigned long long)i) >> (64 - log_n)) : 0;
  P[i] = invn * pow_u64(g, e);
  Q[i] = inv2n * pow_u64(ginv, e);
}


// ------------------------------------------------------------------------------------------------
// Generic root domain (explicit roots, n <= 32768): the reference's own algorithms restated as dense
// O(n^2) kernels.  Upload builds t(x) = prod (x - r_k) (root_poly, coefficient_poly.rs:192-200), the
// coefficient rows of the Lagrange basis L_k = t / ((x - r_k) t'(r_k)) (lagrange_basis, :173-190)
// and g = 1 / rev(t) mod x^(n-1); a proof then needs u = sum_k A_k L_k, v likewise (the unique
// interpolants: same coefficient vectors as mod.rs:233-246), the high half of the schoolbook product
// (coefficient_poly.rs:93-130) and the quotient by the monic t (field/mod.rs:428-469) as
// rev(h) = rev(p_hi) * g mod x^(n-1).  As on the fast domain, w_sum (degree < n) cannot reach the quotient.
__global__ void __launch_bounds__(1024) k_gen_tcoef(const Fr* __restrict__ roots, size_t n, Fr* buf0, Fr* buf1, Fr* __restrict__ out) {
  for (size_t i = threadIdx.x; i <= n; i += blockDim.x) buf0[i] = i == 0 ? Fr::one() : Fr::zero();
  __syncthreads();
  Fr *cur = buf0, *nxt = buf1;
  for (size_t k = 0; k < n; k++) {
    const Fr r = roots[k];
    for (size_t i = threadIdx.x; i <= k + 1; i += blockDim.x) {
      Fr lo = i > 0 ? cur[i - 1] : Fr::zero();
      Fr hi = i <= k ? cur[i] : Fr::zero();
      nxt[i] = lo - r * hi;
    }
    __syncthreads();
    Fr* t = cur; cur = nxt; nxt = t;
  }
  for (size_t i = threadIdx.x; i <= n; i += blockDim.x) out[i] = cur[i];
}

__global__ void k_gen_lagrange_coef(const Fr* __restrict__ roots, const Fr* __restrict__ tc, size_t n, Fr* __restrict__ Lc,
                                    Fr* __restrict__ dinv, int* bad) {
  size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const Fr r = roots[k];
  Fr* row = Lc + k * n;
  Fr b = tc[n], d = b;  // synthetic division of t by (x - r); d = quotient evaluated at r = t'(r)
  row[n - 1] = b;
  for (size_t i = n - 1; i >= 1; i--) {
    b = tc[i] + r * b;
    row[i - 1] = b;
    d = d * r + b;
  }
  if (d.is_zero()) { *bad = 1; dinv[k] = Fr::zero(); return; }  // repeated root
  const Fr inv = inverse(d);
  dinv[k] = inv;
  for (size_t i = 0; i < n; i++) row[i] = row[i] * inv;
}

__global__ void __launch_bounds__(512) k_gen_ginv(const Fr* __restrict__ tc, size_t n, Fr* g) {
  __shared__ Fr sh[512];
  if (n < 2) return;
  if (threadIdx.x == 0) g[0] = Fr::one();  // t is monic
  __syncthreads();
  for (size_t j = 1; j + 1 < n; j++) {
    Fr acc = Fr::zero();
    for (size_t i = 1 + threadIdx.x; i <= j; i += blockDim.x) acc = acc + tc[n - i] * g[j - i];
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (uint32_t off = blockDim.x >> 1; off > 0; off >>= 1) {
      if (threadIdx.x < off) sh[threadIdx.x] = sh[threadIdx.x] + sh[threadIdx.x + off];
      __syncthreads();
    }
    if (threadIdx.x == 0) g[j] = neg(sh[0]);
    __syncthreads();
  }
}

__global__ void k_gen_interp(const Fr* __restrict__ A, const Fr* __restrict__ B, const Fr* __restrict__ Lc, size_t n,
                             Fr* __restrict__ u, Fr* __restrict__ v) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr su = Fr::zero(), sv = Fr::zero();
  for (size_t k = 0; k < n; k++) {
    const Fr l = Lc[k * n + i];
    su = su + A[k] * l;
    sv = sv + B[k] * l;
  }
  u[i] = su;
  v[i] = sv;
}

// rp[n-2-k] = (u * v)[n + k] = sum_{i=k+1}^{n-1} u[i] v[n+k-i],  k = 0 .. n-2
__global__ void k_gen_prod_high(const Fr* __restrict__ u, const Fr* __restrict__ v, size_t n, Fr* __restrict__ rp) {
  size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k + 1 >= n) return;
  Fr acc = Fr::zero();
  for (size_t i = k + 1; i < n; i++) acc = acc + u[i] * v[n + k - i];
  rp[n - 2 - k] = acc;
}

// h[n-2-j] = sum_{i=0}^{j} rp[i] g[j-i],  j = 0 .. n-2;  h[n-1] = 0
__global__ void k_gen_quotient(const Fr* __restrict__ rp, const Fr* __restrict__ g, size_t n, Fr* __restrict__ h) {
  size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  if (j == n - 1) { h[n - 1] = Fr::zero(); return; }
  Fr acc = Fr::zero();
  for (size_t i = 0; i <= j; i++) acc = acc + rp[i] * g[j - i];
  h[n - 2 - j] = acc;
}

// MSM scalar vectors (canonical).  un, vn, hn: Montgomery, natural order; wm: witness, Montgomery.
struct ScalarPlan {
  uint64_t nxi, nxt, nsd, xi_lo, xit_lo, sd_lo;  // this rank's shard
  uint64_t n_input, m;
  int lead;                                      // rank 0 carries the fixed-point terms
};
__global__ void k_msm_scalars(const Fr* __restrict__ un, const Fr* __restrict__ vn, const Fr* __restrict__ hn,
                              const Fr* __restrict__ wm, ScalarPlan pl, Fr r_m, Fr s_m, Fr* __restrict__ SA, Fr* __restrict__ SB,
                              Fr* __restrict__ SC) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t off_xit = pl.nxi + 3, off_sd = off_xit + pl.nxt, total = off_sd + pl.nsd;
  if (i >= total) return;
  if (i < pl.nxi) {
    Fr u = un[pl.xi_lo + i], v = vn[pl.xi_lo + i];
    SA[i] = from_mont(u);
    SB[i] = from_mont(v);
    SC[i] = from_mont(s_m * u + r_m * v);
  } else if (i < off_xit) {
    const int k = (int)(i - pl.nxi);
    Fr one = Fr::zero();
    one.v[0] = 1;
    Fr r = from_mont(r_m), s = from_mont(s_m), z = Fr::zero();
    if (!pl.lead) { r = z; s = z; one = z; }
    SA[i] = k == 0 ? one : (k == 1 ? z : r);                               // alpha1 + r delta1
    SC[i] = k == 0 ? s : (k == 1 ? r : (pl.lead ? from_mont(r_m * s_m) : z));  // s alpha1 + r beta1 + rs delta1
    if (k < 2) SB[i] = k == 0 ? one : s;                                   // beta2 + s delta2
  } else if (i < off_sd) {
    SC[i] = from_mont(hn[pl.xit_lo + (i - off_xit)]);
  } else {
    const uint64_t w = pl.n_input + 1 + pl.sd_lo + (i - off_sd);  // zip truncation: missing weights count as zero
    SC[i] = w < pl.m ? from_mont(wm[w]) : Fr::zero();
  }
}

// affine, canonical proof (or per-rank partial sums): out = a (16 u32) | b (32 u32) | c (16 u32)
__global__ void __launch_bounds__(96) k_finish(const G1XYZZ* __restrict__ ac, const G2XYZZ* __restrict__ b, uint32_t* __restrict__ out) {
  const int w = threadIdx.x >> 5;
  if (threadIdx.x & 31) return;
  if (w < 2) {
    G1Affine p = to_affine(ac[w]);
    Fq x = from_mont(p.x), y = from_mont(p.y);
    uint32_t* o = out + (w == 0 ? 0 : 48);
#pragma unroll
    for (int i = 0; i < 8; i++) { o[i] = x.v[i]; o[8 + i] = y.v[i]; }
  } else {
    G2Affine p = to_affine(*b);
    Fq c[4] = {from_mont(p.x.c0), from_mont(p.x.c1), from_mont(p.y.c0), from_mont(p.y.c1)};
    for (int q = 0; q < 4; q++)
#pragma unroll
      for (int i = 0; i < 8; i++) out[16 + q * 8 + i] = c[q].v[i];
  }
}

// Fold per-rank partial sums (canonical affine, proof layout, partials[w][count][64 u32]) into proofs:
// one block per proof, warp 0 -> A, warp 1 -> C, warp 2 -> B (sequential mixed additions over the ranks).
__device__ __forceinline__ Fq ld_fq_canon(const uint32_t* p) {
  Fq c;
#pragma unroll
  for (int i = 0; i < 8; i++) c.v[i] = p[i];
  return to_mont(c);
}
__global__ void __launch_bounds__(96) k_combine_partials(const uint32_t* __restrict__ partials, int world, size_t count,
                                                         uint32_t* __restrict__ out) {
  const size_t k = blockIdx.x;
  const int w = threadIdx.x >> 5;
  if (k >= count || (threadIdx.x & 31)) return;
  uint32_t* o = out + k * 64;
  if (w < 2) {
    const int off = w == 0 ? 0 : 48;
    G1XYZZ acc = G1XYZZ::inf();
    for (int r = 0; r < world; r++) {
      const uint32_t* rec = partials + ((size_t)r * count + k) * 64 + off;
      G1Affine p;
      p.x = ld_fq_canon(rec);
      p.y = ld_fq_canon(rec + 8);
      acc = madd(acc, p);
    }
    G1Affine a = to_affine(acc);
    Fq x = from_mont(a.x), y = from_mont(a.y);
#pragma unroll
    for (int i = 0; i < 8; i++) { o[off + i] = x.v[i]; o[off + 8 + i] = y.v[i]; }
  } else {
    G2XYZZ acc = G2XYZZ::inf();
    for (int r = 0; r < world; r++) {
      const uint32_t* rec = partials + ((size_t)r * count + k) * 64 + 16;
      G2Affine p;
      p.x.c0 = ld_fq_canon(rec); p.x.c1 = ld_fq_canon(rec + 8);
      p.y.c0 = ld_fq_canon(rec + 16); p.y.c1 = ld_fq_canon(rec + 24);
      acc = madd(acc, p);
    }
    G2Affine a = to_affine(acc);
    Fq c[4] = {from_mont(a.x.c0), from_mont(a.x.c1), from_mont(a.y.c0), from_mont(a.y.c1)};
    for (int q = 0; q < 4; q++)
#pragma unroll
      for (int i = 0; i < 8; i++) o[16 + q * 8 + i] = c[q].v[i];
  }
}

int matvec_launch(zkb_ctx* ctx, const zkb_qap* q, const Fr* wmont, size_t k0, size_t kstride, size_t count, Fr* A, Fr* B, Fr* AB,
                  cudaStream_t st) {
  ZKB_LAUNCH(ctx, k_matvec, cdiv(count, 128), 128, 0, st, q->d_gptr[0], q->d_wire[0], q->d_coeff[0], q->d_gptr[1], q->d_wire[1],
             q->d_coeff[1], wmont, (size_t)q->m, k0, kstride, count, A, B, AB);
  return ZKB_OK;
}
int combine_partials_launch(zkb_ctx* ctx, const uint32_t* d_partials, int world, size_t count, uint32_t* d_out, cudaStream_t st) {
  ZKB_LAUNCH(ctx, k_combine_partials, (unsigned)count, 96, 0, st, d_partials, world, count, d_out);
  return ZKB_OK;
}

}  // namespace zkb

using namespace zkb;

extern "C" {

// ---- QAP ----------------------------------------------------------------------------------------
void zkb_qap_free(zkb_ctx* ctx, zkb_qap* q) {
  if (!q) return;
  if (ctx) cudaSetDevice(ctx->device);
  for (int t = 0; t < 3; t++) {
    cudaFree(q->d_gptr[t]); cudaFree(q->d_wire[t]); cudaFree(q->d_coeff[t]);
    cudaFree(q->d_rptr[t]); cudaFree(q->d_gate[t]); cudaFree(q->d_rcoeff[t]);
  }
  cudaFree(q->d_cosP); cudaFree(q->d_cosQ);
  cudaFree(q->d_roots); cudaFree(q->d_Lc); cudaFree(q->d_dinv); cudaFree(q->d_tc); cudaFree(q->d_ginv);
  delete q;
}

int zkb_qap_upload(zkb_ctx* ctx, const zkb_qap_host* h, zkb_qap** out) {
  if (!ctx || !h || !out) return set_err(ctx, ZKB_ERR_ARG, "zkb_qap_upload: NULL argument");
  *out = nullptr;
  uint64_t n = h->n, m = h->m;
  const bool generic = h->roots != nullptr;
  if (!generic && (n < 2 || (n & (n - 1)) || n > ((uint64_t)1 << 27)))
    return set_err(ctx, ZKB_ERR_ARG, "qap.n = %llu must be a power of two in [2, 2^27] on the roots-of-unity domain",
                   (unsigned long long)n);
  if (generic && (n < 1 || n > ZKB_GENERIC_MAX_N))
    return set_err(ctx, ZKB_ERR_UNSUPPORTED, "qap.n = %llu: explicit root domains are limited to n <= %llu (dense O(n^2) path)",
                   (unsigned long long)n, (unsigned long long)ZKB_GENERIC_MAX_N);
  if (m == 0 || m >= ((uint64_t)1 << 31) || h->n_input + 1 > m)
    return set_err(ctx, ZKB_ERR_ARG, "qap.m = %llu / n_input = %llu invalid", (unsigned long long)m,
                   (unsigned long long)h->n_input);
  for (int t = 0; t < 3; t++)
    if (!h->row_ptr[t] || (h->row_ptr[t][m] && (!h->gate[t] || !h->coeff[t])))
      return set_err(ctx, ZKB_ERR_ARG, "qap rows %d: NULL pointer", t);
  ZKB_CUDA(ctx, cudaSetDevice(ctx->device));
  zkb_qap* q = new zkb_qap();
  q->n = n; q->m = m; q->n_input = h->n_input;
  while (((uint64_t)1 << q->log_n) < n) q->log_n++;
  cudaStream_t st = ctx->stream;
  int rc = ZKB_OK;
  auto fail = [&](int code) { zkb_qap_free(ctx, q); return code; };
  for (int t = 0; t < 3 && rc == ZKB_OK; t++) {
    const uint64_t* rp = h->row_ptr[t];
    uint64_t nnz = rp[m];
    q->nnz[t] = nnz;
    if (nnz >= ((uint64_t)1 << 32)) return fail(set_err(ctx, ZKB_ERR_ARG, "nnz too large"));
    if (rp[0] != 0) return fail(set_err(ctx, ZKB_ERR_ARG, "row_ptr[0] must be 0"));
    for (uint64_t i = 0; i < m; i++)
      if (rp[i] > rp[i + 1]) return fail(set_err(ctx, ZKB_ERR_ARG, "row_ptr not monotone"));
    // by-gate transpose (counting sort on the host: one-time format conversion at upload)
    std::vector<uint32_t> gptr(n + 1, 0), wire(nnz), rptr32(m + 1), perm(nnz);
    for (uint64_t e = 0; e < nnz; e++) {
      if (h->gate[t][e] >= n) return fail(set_err(ctx, ZKB_ERR_ARG, "gate index out of range"));
      gptr[h->gate[t][e] + 1]++;
    }
    for (uint64_t k = 0; k < n; k++) gptr[k + 1] += gptr[k];
    {
      std::vector<uint32_t> cur(gptr.begin(), gptr.end() - 1);
      for (uint64_t i = 0; i < m; i++)
        for (uint64_t e = rp[i]; e < rp[i + 1]; e++) {
          uint32_t pos = cur[h->gate[t][e]]++;
          wire[pos] = (uint32_t)i;
          perm[pos] = (uint32_t)e;
        }
    }
    for (uint64_t i = 0; i <= m; i++) rptr32[i] = (uint32_t)rp[i];
    std::vector<uint64_t> coef_g(nnz * 4);
    for (uint64_t p = 0; p < nnz; p++) memcpy(&coef_g[p * 4], &h->coeff[t][(uint64_t)perm[p] * 4], 32);
    size_t nz = nnz ? nnz : 1;
    if (cudaMalloc(&q->d_gptr[t], (n + 1) * 4) || cudaMalloc(&q->d_wire[t], nz * 4) ||
        cudaMalloc(&q->d_coeff[t], nz * 32) || cudaMalloc(&q->d_rptr[t], (m + 1) * 4) ||
        cudaMalloc(&q->d_gate[t], nz * 4) || cudaMalloc(&q->d_rcoeff[t], nz * 32))
      return fail(set_err(ctx, ZKB_ERR_ALLOC, "qap upload: cudaMalloc failed"));
    cudaMemcpyAsync(q->d_gptr[t], gptr.data(), (n + 1) * 4, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(q->d_rptr[t], rptr32.data(), (m + 1) * 4, cudaMemcpyHostToDevice, st);
    if (nnz) {
      cudaMemcpyAsync(q->d_wire[t], wire.data(), nnz * 4, cudaMemcpyHostToDevice, st);
      cudaMemcpyAsync(q->d_coeff[t], coef_g.data(), nnz * 32, cudaMemcpyHostToDevice, st);
      cudaMemcpyAsync(q->d_gate[t], h->gate[t], nnz * 4, cudaMemcpyHostToDevice, st);
      cudaMemcpyAsync(q->d_rcoeff[t], h->coeff[t], nnz * 32, cudaMemcpyHostToDevice, st);
    }
    cudaError_t e = cudaStreamSynchronize(st);  // host vectors go out of scope
    if (e != cudaSuccess) return fail(set_err(ctx, ZKB_ERR_CUDA, "qap upload copy: %s", cudaGetErrorString(e)));
    rc = vec_to_mont(ctx, q->d_coeff[t], nnz, true, st);
    if (rc == ZKB_OK) rc = vec_to_mont(ctx, q->d_rcoeff[t], nnz, true, st);
    q->h_gptr[t] = std::move(gptr);
    q->h_wire[t] = std::move(wire);
  }
  if (rc != ZKB_OK) return fail(rc);
  q->generic = generic;
  if (!generic) {
    if (cudaMalloc(&q->d_cosP, n * 32) || cudaMalloc(&q->d_cosQ, n * 32))
      return fail(set_err(ctx, ZKB_ERR_ALLOC, "qap upload: workspace cudaMalloc failed"));
    Fr g = host_omega(q->log_n + 1, false), ginv = host_omega(q->log_n + 1, true);
    Fr invn = inverse(fr_from_u64(n)), inv2n = inverse(fr_from_u64(2 * n));
    {
      zkb_ctx* c = ctx;
      k_coset_tables<<<cdiv(n, 256), 256, 0, st>>>(q->d_cosP, q->d_cosQ, g, ginv, invn, inv2n, q->log_n);
      c->launches++;
    }
    Fr* tw;
    rc = get_twiddles(ctx, q->log_n, false, &tw);
    if (rc == ZKB_OK) rc = get_twiddles(ctx, q->log_n, true, &tw);
    if (rc != ZKB_OK) return fail(rc);
  } else {
    q->h_roots.resize(n);
    for (uint64_t k = 0; k < n; k++) q->h_roots[k] = fr_from_limbs(h->roots + 4 * k);
    void* p;
    if ((rc = scratch_get(ctx, 8, 2 * (n + 1) * sizeof(Fr), &p)) != ZKB_OK) return fail(rc);
    Fr* buf = (Fr*)p;
    if ((rc = scratch_get(ctx, 9, sizeof(int), &p)) != ZKB_OK) return fail(rc);
    int* d_bad = (int*)p;
    if (cudaMalloc(&q->d_roots, n * 32) || cudaMalloc(&q->d_Lc, n * n * 32) || cudaMalloc(&q->d_dinv, n * 32) ||
        cudaMalloc(&q->d_tc, (n + 1) * 32) || cudaMalloc(&q->d_ginv, n * 32))
      return fail(set_err(ctx, ZKB_ERR_ALLOC, "qap upload: generic-domain tables (%.1f MiB) cudaMalloc failed",
                          (double)n * n * 32 / 1048576.0));
    cudaMemsetAsync(d_bad, 0, sizeof(int), st);
    cudaMemcpyAsync(q->d_roots, q->h_roots.data(), n * 32, cudaMemcpyHostToDevice, st);
    k_gen_tcoef<<<1, 1024, 0, st>>>(q->d_roots, n, buf, buf + n + 1, q->d_tc);
    k_gen_lagrange_coef<<<cdiv(n, 64), 64, 0, st>>>(q->d_roots, q->d_tc, n, q->d_Lc, q->d_dinv, d_bad);
    k_gen_ginv<<<1, 512, 0, st>>>(q->d_tc, n, q->d_ginv);
    ctx->launches += 3;
    int bad = 0;
    cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, st);
    cudaError_t e2 = cudaStreamSynchronize(st);
    if (e2 != cudaSuccess) return fail(set_err(ctx, ZKB_ERR_CUDA, "qap upload (generic domain): %s", cudaGetErrorString(e2)));
    if (bad) return fail(set_err(ctx, ZKB_ERR_DIV_ZERO, "qap roots are not pairwise distinct (the reference's lagrange_basis "
                                                        "would divide by zero, coefficient_poly.rs:183)"));
  }
  cudaError_t e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return fail(set_err(ctx, ZKB_ERR_CUDA, "qap upload: %s", cudaGetErrorString(e)));
  *out = q;
  return ZKB_OK;
}

// ---- polynomial stage ---------------------------------------------------------------------------
// lane workspace: 8 vectors of n Fr (0 A->u_br  1 B->v_br  2 AB->c_br  3 uc  4 vc->d_br  5 u_nat  6 v_nat
// 7 h_nat), the witness as uploaded (canonical) and in Montgomery form
struct Work {
  Fr *ws, *wcanon, *wmont;
};
static int work_get(zkb_ctx* ctx, zkb_lane* L, const zkb_qap* q, Work* w, size_t ws_elems = 0) {
  void* p;
  ZKB_TRY(scratch_get_in(ctx, L->scratch, 11, (ws_elems ? ws_elems : 8 * q->n) * sizeof(Fr), &p));
  w->ws = (Fr*)p;
  ZKB_TRY(scratch_get_in(ctx, L->scratch, 12, q->m * sizeof(Fr), &p));
  w->wcanon = (Fr*)p;
  ZKB_TRY(scratch_get_in(ctx, L->scratch, 13, q->m * sizeof(Fr), &p));
  w->wmont = (Fr*)p;
  return ZKB_OK;
}

static int poly_stage(zkb_ctx* ctx, const zkb_qap* q, const Work& w, const Fr* d_w_canon, cudaStream_t st) {
  const size_t n = q->n, m = q->m;
  Fr* ws = w.ws;
  Fr *A = ws, *B = ws + n, *AB = ws + 2 * n, *uc = ws + 3 * n, *vc = ws + 4 * n, *un = ws + 5 * n, *vn = ws + 6 * n,
     *hn = ws + 7 * n;
  ZKB_CUDA(ctx, cudaMemcpyAsync(w.wmont, d_w_canon, m * 32, cudaMemcpyDeviceToDevice, st));
  ZKB_TRY(vec_to_mont(ctx, w.wmont, m, true, st));
  ZKB_TRY(matvec_launch(ctx, q, w.wmont, 0, 1, n, A, B, AB, st));
  if (q->generic) {  // dense path: interpolate, high half of the product, quotient by t
    ZKB_LAUNCH(ctx, k_gen_interp, cdiv(n, 64), 64, 0, st, A, B, q->d_Lc, n, un, vn);
    if (n > 1) ZKB_LAUNCH(ctx, k_gen_prod_high, cdiv(n, 64), 64, 0, st, un, vn, n, AB);
    ZKB_LAUNCH(ctx, k_gen_quotient, cdiv(n, 64), 64, 0, st, AB, q->d_ginv, n, hn);
    return ZKB_OK;
  }
  ZKB_TRY(ntt_dif(ctx, A, q->log_n, true, st));   // n * u_sum, bit-reversed
  ZKB_TRY(ntt_dif(ctx, B, q->log_n, true, st));
  ZKB_TRY(ntt_dif(ctx, AB, q->log_n, true, st));  // n * (p_lo + p_hi), bit-reversed
  Fr invn = inverse(fr_from_u64(n));
  ZKB_TRY(bitrev_permute(ctx, un, A, q->log_n, &invn, st));
  ZKB_TRY(bitrev_permute(ctx, vn, B, q->log_n, &invn, st));
  ZKB_LAUNCH(ctx, k_mul2, cdiv(n, 256), 256, 0, st, uc, A, vc, B, q->d_cosP, n);  // u_k g^k, v_k g^k (bit-reversed)
  ZKB_TRY(ntt_dit(ctx, uc, q->log_n, false, st));  // u_sum(g w^i), natural
  ZKB_TRY(ntt_dit(ctx, vc, q->log_n, false, st));
  ZKB_TRY(vec_mul(ctx, vc, uc, vc, n, st));
  ZKB_TRY(ntt_dif(ctx, vc, q->log_n, true, st));   // n * g^k (p_lo - p_hi), bit-reversed
  Fr inv2n = inverse(fr_from_u64(2 * n));
  ZKB_LAUNCH(ctx, k_hfinal, cdiv(n, 256), 256, 0, st, uc, AB, vc, q->d_cosQ, inv2n, n);
  ZKB_TRY(bitrev_permute(ctx, hn, uc, q->log_n, nullptr, st));
  return ZKB_OK;
}

// The reference's prove() zips whatever it is given (mod.rs:237..288); a CRS made for another QAP would silently give a
// proof that cannot verify.  setup() fixes n = qap.degree, sum_gamma = input + 1 and sum_delta = m - input - 1 entries
// (mod.rs:146-168), so all three must agree with the QAP.
static int check_pair(zkb_ctx* ctx, const zkb_qap* q, const zkb_crs* c) {
  if (c->n != q->n) return set_err(ctx, ZKB_ERR_ARG, "CRS (n=%llu) does not belong to QAP (n=%llu)",
                                   (unsigned long long)c->n, (unsigned long long)q->n);
  if (c->n_sum_gamma != q->n_input + 1 || c->n_sum_gamma + c->n_sum_delta != q->m)
    return set_err(ctx, ZKB_ERR_ARG, "CRS (sum_gamma %llu + sum_delta %llu entries) does not belong to QAP (m=%llu wires, %llu inputs)",
                   (unsigned long long)c->n_sum_gamma, (unsigned long long)c->n_sum_delta, (unsigned long long)q->m,
                   (unsigned long long)q->n_input);
  return ZKB_OK;
}

struct ProveOut {
  G1XYZZ* ac;       // A, C accumulators
  G2XYZZ* b;
  uint32_t* proof;  // 64 u32
};
static int prove_out(zkb_ctx* ctx, DevBuf* slots, ProveOut* o) {
  void* p;
  ZKB_TRY(scratch_get_in(ctx, slots, 6, 2 * sizeof(G1XYZZ) + sizeof(G2XYZZ) + 1024, &p));  // + partial | proof | status
  o->ac = (G1XYZZ*)p;
  o->b = (G2XYZZ*)(o->ac + 2);
  o->proof = (uint32_t*)(o->b + 1);
  return ZKB_OK;
}

// Enqueue stages 1-7 over this rank's CRS shard on lane L; nothing waits on the host.  The 256-byte
// result (proof for world 1, else the rank's partial sums) lands in L->h_proof when L->hi drains.
//   L->hi:  [H2D] polynomial stage, scalars, sort G2, sort G1 ... tail G2 ........ finish, D2H
//   L->lo:                                    accumulate G2, accumulate G1
//   L->hi2:                                                                tail G1
// comm != NULL: ONE proof over all ranks of `comm` (CRS in layout 1): the polynomial stage is sharded (shard.cu, three
// exchanges over peer memory on channel `ch`), the partial sums are exchanged and folded on the device, and every rank
// ends up with the complete proof.
// batch: other proofs are in flight on other lanes (their accumulations hide this proof's bucket hierarchy, so the plan
// with the least multiplier work wins); else the caller waits for THIS proof and the hierarchy runs on quads.
static int prove_enqueue(zkb_ctx* ctx, zkb_lane* L, const zkb_qap* q, const zkb_crs* c, const uint64_t* weights,
                         int on_device, const uint64_t* r, const uint64_t* s, zkb_comm* comm = nullptr, int ch = 0, bool batch = false) {
  cudaStream_t hi = L->hi, lo = L->lo;
  // Every allocation, table and plan first, launches after: scratch growth synchronises the device and the first use of a
  // twiddle table synchronises a stream, and neither may happen behind a kernel that is waiting for a peer.
  Work w;
  ZKB_TRY(work_get(ctx, L, q, &w, comm ? 7 * (q->n >> comm->lg) : 0));
  ScalarPlan sp;
  sp.nxi = c->nxi(); sp.nxt = c->nxt(); sp.nsd = c->nsd();
  sp.xi_lo = c->xi_lo; sp.xit_lo = c->xit_lo; sp.sd_lo = c->sd_lo;
  sp.n_input = q->n_input; sp.m = q->m; sp.lead = c->rank == 0;
  void* p;
  ZKB_TRY(scratch_get_in(ctx, L->scratch, 10, (c->g1_cnt + (sp.nxi + 3) + c->g2_cnt + 4) * sizeof(Fr), &p));
  Fr* SC = (Fr*)p;
  Fr* SA = SC + c->g1_cnt;
  Fr* SB = SA + sp.nxi + 3;
  ProveOut o;
  ZKB_TRY(prove_out(ctx, L->scratch, &o));
  MsmJob jb = {SB, c->g2_cnt};
  MsmJob j1[2] = {{SA, sp.nxi + 3}, {SC, c->g1_cnt}};
  MsmPlan P2, P1;
  ZKB_TRY(msm_prepare(ctx, L->scratch, 3, 2, c->g2, c->g2_cnt, c->c2, &jb, 1, o.b, &P2));
  ZKB_TRY(msm_prepare(ctx, L->scratch, 0, 1, c->g1, c->g1_cnt, c->c1, j1, 2, o.ac, &P1));
  // hierarchy plan: quads unless this is one of several proofs in flight AND the accumulations are long enough to hide
  // the (then cheaper) thread-level hierarchy -- a rank of an 8-GPU proof at 2^20 has 6.8 M G1 records: quads
  P1.tail = P2.tail = !batch ? 0 : (P1.max_recs >= ((size_t)1 << 24) ? 1 : 2);
  if (comm) ZKB_TRY(shard_prepare(ctx, comm, q->log_n));
  const Fr* d_w = (const Fr*)weights;
  if (!on_device) {
    ZKB_CUDA(ctx, cudaMemcpyAsync(w.wcanon, weights, q->m * 32, cudaMemcpyHostToDevice, hi));
    d_w = w.wcanon;
  }
  const size_t n = q->n;
  Fr *un = w.ws + 5 * n, *vn = w.ws + 6 * n, *hn = w.ws + 7 * n;
  uint32_t epoch = 0;
  if (comm) {
    const size_t ml = n >> comm->lg;
    epoch = ++comm->epoch[ch];
    ZKB_CUDA(ctx, cudaMemcpyAsync(w.wmont, d_w, q->m * 32, cudaMemcpyDeviceToDevice, hi));
    ZKB_TRY(vec_to_mont(ctx, w.wmont, q->m, true, hi));
    ZKB_TRY(shard_poly_stage(ctx, comm, ch, epoch, q, w.ws, w.wmont, hi));
    un = w.ws + 3 * ml; vn = w.ws + 4 * ml; hn = w.ws + 6 * ml;
  } else {
    ZKB_TRY(poly_stage(ctx, q, w, d_w, hi));
  }
  ZKB_LAUNCH(ctx, k_msm_scalars, cdiv(c->g1_cnt, 256), 256, 0, hi, un, vn, hn, w.wmont, sp, fr_from_limbs(r), fr_from_limbs(s), SA,
             SB, SC);
  ZKB_TRY(msm_sort(ctx, P2, hi));
  ZKB_CUDA(ctx, cudaEventRecord(L->ev[2], hi));
  ZKB_TRY(msm_sort(ctx, P1, hi));
  ZKB_CUDA(ctx, cudaEventRecord(L->ev[0], hi));
  ZKB_CUDA(ctx, cudaStreamWaitEvent(lo, L->ev[2], 0));
  ZKB_TRY(msm_accumulate(ctx, P2, lo));
  ZKB_CUDA(ctx, cudaEventRecord(L->ev[3], lo));
  ZKB_CUDA(ctx, cudaStreamWaitEvent(lo, L->ev[0], 0));
  ZKB_TRY(msm_accumulate(ctx, P1, lo));
  ZKB_CUDA(ctx, cudaEventRecord(L->ev[1], lo));
  ZKB_CUDA(ctx, cudaStreamWaitEvent(hi, L->ev[3], 0));
  ZKB_TRY(msm_tail(ctx, P2, hi));
  // the G1 tail on its own latency-class stream: at small sizes the G2 tail is still running when the G1
  // accumulation ends, and both are chains of low-occupancy kernels (2^16: single-proof latency 4.5 -> 3.8 ms)
  ZKB_CUDA(ctx, cudaStreamWaitEvent(L->hi2, L->ev[1], 0));
  ZKB_TRY(msm_tail(ctx, P1, L->hi2));
  ZKB_CUDA(ctx, cudaEventRecord(L->ev[4], L->hi2));
  ZKB_CUDA(ctx, cudaStreamWaitEvent(hi, L->ev[4], 0));
  ZKB_LAUNCH(ctx, k_finish, 1, 96, 0, hi, o.ac, o.b, o.proof);
  if (comm) {  // partial sums -> every rank's window; wait for all of them; fold; the proof replaces the partial record
    int* d_status = reinterpret_cast<int*>(o.proof + 128);
    ZKB_TRY(shard_exchange_partials(ctx, comm, ch, epoch, o.proof, o.proof + 64, d_status, hi));
    ZKB_CUDA(ctx, cudaMemcpyAsync(L->h_proof, o.proof + 64, 256, cudaMemcpyDeviceToHost, hi));
    ZKB_CUDA(ctx, cudaMemcpyAsync((char*)L->h_proof + 256, d_status, sizeof(int), cudaMemcpyDeviceToHost, hi));
    return ZKB_OK;
  }
  ZKB_CUDA(ctx, cudaMemcpyAsync(L->h_proof, o.proof, 256, cudaMemcpyDeviceToHost, hi));
  return ZKB_OK;
}

static int prove_collect(zkb_ctx* ctx, zkb_lane* L, void* out, bool sharded = false) {
  ZKB_CUDA(ctx, cudaStreamSynchronize(L->hi));
  memcpy(out, L->h_proof, 256);
  if (sharded && *reinterpret_cast<const int*>((const char*)L->h_proof + 256) != 0)
    return set_err(ctx, ZKB_ERR_COMM, "sharded proof: a peer did not arrive within the exchange timeout (ZKB_COMM_TIMEOUT_MS); "
                                      "the communicator is unusable from here on");
  return ZKB_OK;
}

static int prove_common(zkb_ctx* ctx, const zkb_qap* q, const zkb_crs* c, const uint64_t* weights, int on_device,
                        const uint64_t* r, const uint64_t* s, uint64_t* out) {
  ZKB_TRY(check_pair(ctx, q, c));
  ZKB_CUDA(ctx, cudaSetDevice(ctx->device));
  ZKB_TRY(prove_enqueue(ctx, &ctx->lanes[0], q, c, weights, on_device, r, s));
  return prove_collect(ctx, &ctx->lanes[0], out);
}

int zkb_prove(zkb_ctx* ctx, const zkb_qap* q, const zkb_crs* c, const uint64_t* weights, const uint64_t r[4],
              const uint64_t s[4], zkb_proof* out) {
  if (!ctx || !q || !c || !weights || !r || !s || !out) return set_err(ctx, ZKB_ERR_ARG, "zkb_prove: NULL argument");
  if (c->world != 1) return set_err(ctx, ZKB_ERR_ARG, "zkb_prove needs an unsharded CRS; use zkb_prove_partial");
  return prove_common(ctx, q, c, weights, 0, r, s, (uint64_t*)out);
}
int zkb_prove_dev(zkb_ctx* ctx, const zkb_qap* q, const zkb_crs* c, const uint64_t* d_weights, const uint64_t r[4],
                  const uint64_t s[4], zkb_proof* out) {
  if (!ctx || !q || !c || !d_weights || !r || !s || !out) return set_err(ctx, ZKB_ERR_ARG, "zkb_prove_dev: NULL argument");
  if (c->world != 1) return set_err(ctx, ZKB_ERR_ARG, "zkb_prove needs an unsharded CRS; use zkb_prove_partial");
  return prove_common(ctx, q, c, d_weights, 1, r, s, (uint64_t*)out);
}

// Throughput mode: `count` independent proofs over the same QAP / CRS, several in flight (one per lane: batch_lane_count),
// so the latency-class stages of proof i+1 and the tails of proof i fill the gaps of the bucket
// accumulations.  Same results as `count` zkb_prove calls.
// Proofs in flight.  Measured on B200 (profiles/r01_lanes_by_size.txt, ms per proof for 1 / 2 / 3 / 4 lanes):
// 2^16 4.27 / 2.82 / 2.34 / 2.19, 2^18 8.66 / 7.22 / 7.06 / 7.23, 2^20 23.8 / 23.2 / 23.2 / -- : small proofs are
// dominated by the latency-bound tails (bucket hierarchy, head folding), which more proofs in flight hide.
// Witnesses that come from HOST memory add a copy stage in front of every proof: at 2^20 a third proof in flight hides it
// (ms per proof with 2 / 3 / 4 lanes, device-resident witnesses 23.24 / 23.33 / 23.94, host witnesses 23.65 / 23.32 / 23.54:
// profiles/r02_lanes_host_witness.txt), so host witnesses get three lanes up to 2^22 (per-lane scratch: ~1.3 GB at 2^20).
static size_t batch_lane_count(const zkb_ctx* ctx, const zkb_qap* q, bool host_witness) {
  if (ctx->batch_lanes >= 1 && ctx->batch_lanes <= 4) return (size_t)ctx->batch_lanes;
  if (q->n <= ((uint64_t)1 << 17)) return 4;
  if (q->n <= ((uint64_t)1 << 19)) return 3;
  return host_witness && q->n <= ((uint64_t)1 << 22) ? 3 : 2;
}

int zkb_prove_batch(zkb_ctx* ctx, const zkb_qap* q, const zkb_crs* c, const uint64_t* const* weights, int on_device,
                    const uint64_t* r, const uint64_t* s, size_t count, zkb_proof* out) {
  if (!ctx || !q || !c || (count && (!weights || !r || !s || !out))) return set_err(ctx, ZKB_ERR_ARG, "zkb_prove_batch: NULL argument");
  ZKB_TRY(check_pair(ctx, q, c));
  ZKB_CUDA(ctx, cudaSetDevice(ctx->device));
  for (size_t i = 0; i < count; i++)
    if (!weights[i]) return set_err(ctx, ZKB_ERR_ARG, "zkb_prove_batch: weights[%zu] is NULL", i);
  int rc = ZKB_OK;
  size_t done = 0;
  const size_t nl = batch_lane_count(ctx, q, !on_device);
  for (size_t l = 0; l < nl && l < count; l++) ZKB_TRY(lane_get(ctx, (int)l, nullptr));
  for (size_t i = 0; i < count && rc == ZKB_OK; i++) {
    zkb_lane* L = &ctx->lanes[i % nl];
    if (i >= nl) {
      rc = prove_collect(ctx, L, &out[i - nl]);
      done = i - nl + 1;
      if (rc != ZKB_OK) break;
    }
    rc = prove_enqueue(ctx, L, q, c, weights[i], on_device, r + 4 * i, s + 4 * i, nullptr, 0, count > 1);
  }
  for (size_t i = done; i < count && rc == ZKB_OK; i++) rc = prove_collect(ctx, &ctx->lanes[i % nl], &out[i]);
  if (rc != ZKB_OK) cudaDeviceSynchronize();  // leave no work in flight behind an error
  return rc;
}

// ---- one proof over all ranks of a communicator ------------------------------------------------------
static int check_shard(zkb_ctx* ctx, zkb_comm* comm, const zkb_qap* q, const zkb_crs* c) {
  ZKB_TRY(check_pair(ctx, q, c));
  if (q->generic) return set_err(ctx, ZKB_ERR_UNSUPPORTED, "sharded proofs need the roots-of-unity domain");
  ZKB_TRY(shard_check(ctx, comm, q->log_n));
  if (c->layout != 1 || c->world != comm->world || c->rank != comm->rank)
    return set_err(ctx, ZKB_ERR_ARG, "sharded proof: the CRS was not made by zkb_setup_shard / zkb_crs_upload_shard for this communicator");
  return ZKB_OK;
}

int zkb_prove_shard_enqueue(zkb_ctx* ctx, zkb_comm* comm, const zkb_qap* q, const zkb_crs* c, const uint64_t* weights, int on_device,
                            const uint64_t r[4], const uint64_t s[4], int lane) {
  if (!ctx || !comm || !q || !c || !weights || !r || !s) return set_err(ctx, ZKB_ERR_ARG, "zkb_prove_shard_enqueue: NULL argument");
  if (lane < 0 || lane >= 4) return set_err(ctx, ZKB_ERR_ARG, "zkb_prove_shard_enqueue: lane %d out of range [0, 4)", lane);
  ZKB_TRY(check_shard(ctx, comm, q, c));
  ZKB_CUDA(ctx, cudaSetDevice(ctx->device));
  ZKB_TRY(lane_get(ctx, lane, nullptr));
  return prove_enqueue(ctx, &ctx->lanes[lane], q, c, weights, on_device, r, s, comm, lane);
}
int zkb_prove_shard_collect(zkb_ctx* ctx, zkb_comm* comm, int lane, zkb_proof* out) {
  if (!ctx || !comm || !out || lane < 0 || lane >= 4 || !ctx->lanes[lane].hi) return set_err(ctx, ZKB_ERR_ARG, "zkb_prove_shard_collect: bad argument");
  ZKB_CUDA(ctx, cudaSetDevice(ctx->device));
  return prove_collect(ctx, &ctx->lanes[lane], out, true);
}
int zkb_prove_shard(zkb_ctx* ctx, zkb_comm* comm, const zkb_qap* q, const zkb_crs* c, const uint64_t* weights, int on_device,
                    const uint64_t r[4], const uint64_t s[4], zkb_proof* out) {
  if (!out) return set_err(ctx, ZKB_ERR_ARG, "zkb_prove_shard: NULL argument");
  ZKB_TRY(zkb_prove_shard_enqueue(ctx, comm, q, c, weights, on_device, r, s, 0));
  return zkb_prove_shard_collect(ctx, comm, 0, out);
}
// `count` proofs, each over all ranks, several in flight (one per lane / exchange channel).  Every rank must make the
// same sequence of calls with the same count.
int zkb_prove_shard_batch(zkb_ctx* ctx, zkb_comm* comm, const zkb_qap* q, const zkb_crs* c, const uint64_t* const* weights, int on_device,
                          const uint64_t* r, const uint64_t* s, size_t count, zkb_proof* out) {
  if (!ctx || !comm || !q || !c || (count && (!weights || !r || !s || !out))) return set_err(ctx, ZKB_ERR_ARG, "zkb_prove_shard_batch: NULL argument");
  ZKB_TRY(check_shard(ctx, comm, q, c));
  ZKB_CUDA(ctx, cudaSetDevice(ctx->device));
  for (size_t i = 0; i < count; i++)
    if (!weights[i]) return set_err(ctx, ZKB_ERR_ARG, "zkb_prove_shard_batch: weights[%zu] is NULL", i);
  size_t nl = 4;
  if (ctx->batch_lanes >= 1 && ctx->batch_lanes <= 4) nl = (size_t)ctx->batch_lanes;
  else if ((q->n >> comm->lg) > ((uint64_t)1 << 19)) nl = 2;
  else if ((q->n >> comm->lg) > ((uint64_t)1 << 17)) nl = 3;
  for (size_t l = 0; l < nl && l < count; l++) ZKB_TRY(lane_get(ctx, (int)l, nullptr));
  int rc = ZKB_OK;
  size_t done = 0;
  for (size_t i = 0; i < count && rc == ZKB_OK; i++) {
    const int lane = (int)(i % nl);
    if (i >= nl) {
      rc = prove_collect(ctx, &ctx->lanes[lane], &out[i - nl], true);
      done = i - nl + 1;
      if (rc != ZKB_OK) break;
    }
    rc = prove_enqueue(ctx, &ctx->lanes[lane], q, c, weights[i], on_device, r + 4 * i, s + 4 * i, comm, lane, count > 1);
  }
  for (size_t i = done; i < count && rc == ZKB_OK; i++) rc = prove_collect(ctx, &ctx->lanes[i % nl], &out[i], true);
  if (rc != ZKB_OK) cudaDeviceSynchronize();
  return rc;
}

int zkb_prove_partial(zkb_ctx* ctx, const zkb_qap* q, const zkb_crs* c, const uint64_t* weights, int on_device,
                      const uint64_t r[4], const uint64_t s[4], uint64_t* out_partial) {
  if (!ctx || !q || !c || !weights || !r || !s || !out_partial)
    return set_err(ctx, ZKB_ERR_ARG, "zkb_prove_partial: NULL argument");
  return prove_common(ctx, q, c, weights, on_device, r, s, out_partial);
}

int zkb_prove_combine_batch(zkb_ctx* ctx, const uint64_t* partials, int world, size_t count, zkb_proof* out) {
  if (!ctx || (count && (!partials || !out)) || world < 1) return set_err(ctx, ZKB_ERR_ARG, "zkb_prove_combine: bad argument");
  if (!count) return ZKB_OK;
  ZKB_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const size_t in_bytes = (size_t)world * count * 256, out_bytes = count * 256;
  void* p;
  ZKB_TRY(scratch_get(ctx, 7, in_bytes + out_bytes, &p));
  uint32_t* d_in = (uint32_t*)p;
  uint32_t* d_out = d_in + in_bytes / 4;
  ZKB_CUDA(ctx, cudaMemcpyAsync(d_in, partials, in_bytes, cudaMemcpyHostToDevice, st));
  ZKB_TRY(combine_partials_launch(ctx, d_in, world, count, d_out, st));
  ZKB_CUDA(ctx, cudaMemcpyAsync(out, d_out, out_bytes, cudaMemcpyDeviceToHost, st));
  ZKB_CUDA(ctx, cudaStreamSynchronize(st));
  return ZKB_OK;
}

int zkb_prove_combine(zkb_ctx* ctx, const uint64_t* partials, int world, zkb_proof* out) {
  return zkb_prove_combine_batch(ctx, partials, world, 1, out);
}

int zkb_qap_h(zkb_ctx* ctx, const zkb_qap* q, const uint64_t* weights, uint64_t* u_sum, uint64_t* v_sum, uint64_t* h) {
  if (!ctx || !q || !weights) return set_err(ctx, ZKB_ERR_ARG, "zkb_qap_h: NULL argument");
  ZKB_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  Work w;
  ZKB_TRY(work_get(ctx, &ctx->lanes[0], q, &w));
  ZKB_CUDA(ctx, cudaMemcpyAsync(w.wcanon, weights, q->m * 32, cudaMemcpyHostToDevice, st));
  ZKB_TRY(poly_stage(ctx, q, w, w.wcanon, st));
  const size_t n = q->n;
  uint64_t* dst[3] = {u_sum, v_sum, h};
  for (int k = 0; k < 3; k++) {
    if (!dst[k]) continue;
    Fr* v = w.ws + (5 + k) * n;
    Fr* tmp = w.ws + 3 * n;
    ZKB_CUDA(ctx, cudaMemcpyAsync(tmp, v, n * 32, cudaMemcpyDeviceToDevice, st));
    ZKB_TRY(vec_to_mont(ctx, tmp, n, false, st));
    ZKB_CUDA(ctx, cudaMemcpyAsync(dst[k], tmp, n * 32, cudaMemcpyDeviceToHost, st));
  }
  ZKB_CUDA(ctx, cudaStreamSynchronize(st));
  return ZKB_OK;
}

}  // extern "C"
