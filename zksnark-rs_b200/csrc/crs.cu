// CRS objects of libzkb200.so: `SigmaG1<G1Local>` / `SigmaG2<G2Local>` (/root/reference/src/groth16/
// mod.rs:105-121) resident in HBM as the window-expanded base tables the MSMs read, plus
// groth16::setup (mod.rs:134-197) on the device for the roots-of-unity domain.
//
// G1 table row 0 = [xi1 shard | alpha1 beta1 delta1 | xi_t shard | sum_delta shard]; G2 table row
// 0 = [xi2 shard | beta2 delta2]; row j = 2^(c*j) times row 0 (msm_impl.cuh).  Putting the fixed
// points into the tables lets prove() fold alpha1 + r*delta1 etc. into the MSMs as extra terms.
#include <stdlib.h>
#include <string.h>
#include "common.cuh"

namespace zkb {

// ------------------------------------------------------------------------------------------------
// setup kernels (groth16/mod.rs:134-197 on the omega domain)
// L_k(x) = (x^n - 1) w^k / (n (x - w^k))
__global__ void k_lagrange(Fr* L, Fr x, Fr tx_over_n, Fr omega, size_t n, int* bad) {
  size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  Fr wk = pow_u64(omega, (uint64_t)k);
  Fr den = x - wk;
  if (den.is_zero()) { *bad = 1; L[k] = Fr::zero(); return; }
  L[k] = tx_over_n * wk * inverse(den);
}

// generic root domain: L_k(x) = t(x) / ((x - r_k) t'(r_k))
__global__ void k_lagrange_generic(Fr* L, Fr x, Fr tx, const Fr* __restrict__ roots, const Fr* __restrict__ dinv, size_t n, int* bad) {
  size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  Fr den = x - roots[k];
  if (den.is_zero()) { *bad = 1; L[k] = Fr::zero(); return; }
  L[k] = tx * dinv[k] * inverse(den);
}

// one warp per wire: lin_i = (beta u_i(x) + alpha v_i(x) + w_i(x)) * (i <= n_input ? 1/gamma : 1/delta)
__device__ __forceinline__ Fr warp_sum(Fr v) {
  for (int off = 16; off > 0; off >>= 1) {
    Fr o;
#pragma unroll
    for (int i = 0; i < 8; i++) o.v[i] = __shfl_down_sync(0xffffffffu, v.v[i], off);
    v = v + o;
  }
  return v;
}
__device__ __forceinline__ Fr row_eval(const uint32_t* rptr, const uint32_t* gate, const Fr* coef, const Fr* L, size_t i,
                                       int lane) {
  Fr acc = Fr::zero();
  for (uint32_t p = rptr[i] + lane, e = rptr[i + 1]; p < e; p += 32) acc = acc + coef[p] * L[gate[p]];
  return warp_sum(acc);
}
__global__ void k_lin(const uint32_t* ru, const uint32_t* gu, const Fr* cu, const uint32_t* rv, const uint32_t* gv,
                      const Fr* cv, const uint32_t* rw, const uint32_t* gw, const Fr* cw, const Fr* L, size_t m,
                      size_t n_input, Fr alpha, Fr beta, Fr inv_gamma, Fr inv_delta, Fr* out) {
  size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (i >= m) return;
  Fr u = row_eval(ru, gu, cu, L, i, lane);
  Fr v = row_eval(rv, gv, cv, L, i, lane);
  Fr w = row_eval(rw, gw, cw, L, i, lane);
  if (lane == 0) out[i] = (beta * u + alpha * v + w) * (i <= n_input ? inv_gamma : inv_delta);
}

// layout 1 (shard.cu, layout S): local index s = k1*q + t <-> global index (rank*q + t) + m*k1
__global__ void k_gather_S(Fr* __restrict__ dst, const Fr* __restrict__ src, uint64_t rank, uint64_t q, uint64_t m, uint64_t count) {
  const uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= count) return;
  const uint64_t k1 = s / q, t = s - k1 * q;
  dst[s] = src[rank * q + t + m * k1];
}

}  // namespace zkb

using namespace zkb;

extern "C" {

void zkb_crs_free(zkb_ctx* ctx, zkb_crs* c) {
  if (!c) return;
  if (ctx) cudaSetDevice(ctx->device);
  cudaFree(c->g1); cudaFree(c->g2); cudaFree(c->sum_gamma); cudaFree(c->gamma2);
  delete c;
}

static void shard(uint64_t len, int rank, int world, uint64_t* lo, uint64_t* hi) {
  *lo = len * (uint64_t)rank / (uint64_t)world;
  *hi = len * (uint64_t)(rank + 1) / (uint64_t)world;
}

static int crs_new(zkb_ctx* ctx, uint64_t n, uint64_t nsg, uint64_t nsd, int rank, int world, int layout, zkb_crs** out) {
  zkb_crs* c = new zkb_crs();
  c->n = n; c->n_sum_gamma = nsg; c->n_sum_delta = nsd;
  c->rank = rank; c->world = world; c->layout = layout;
  if (layout == 1) {  // xi / xi_t in the sharded transform's output layout: n / world entries, local indices
    if (n % ((uint64_t)world * world)) { delete c; return set_err(ctx, ZKB_ERR_ARG, "crs: n = %llu is not a multiple of world^2", (unsigned long long)n); }
    c->xi_lo = 0; c->xi_hi = n / world;
    c->xit_lo = 0; c->xit_hi = n / world - (rank == world - 1 ? 1 : 0);  // xi_t has n - 1 entries: the last index belongs to the last rank
  } else {
    shard(n, rank, world, &c->xi_lo, &c->xi_hi);
    shard(n - 1, rank, world, &c->xit_lo, &c->xit_hi);
  }
  shard(nsd, rank, world, &c->sd_lo, &c->sd_hi);
  c->g1_cnt = c->nxi() + 3 + c->nxt() + c->nsd();
  c->g2_cnt = c->nxi() + 2;
  c->c1 = msm_pick_c(c->g1_cnt);
  c->c2 = msm_pick_c(c->g2_cnt);
  if (const char* e = getenv("ZKB_MSM_C2")) {  // developer switch: window size of the G2 table only
    int v = atoi(e);
    if (v >= 2 && v <= 23) c->c2 = v;
  }
  if (cudaMalloc(&c->g1, (size_t)msm_windows(c->c1) * c->g1_cnt * sizeof(G1Affine)) ||
      cudaMalloc(&c->g2, (size_t)msm_windows(c->c2) * c->g2_cnt * sizeof(G2Affine)) ||
      cudaMalloc(&c->sum_gamma, (nsg + 1) * sizeof(G1Affine)) || cudaMalloc(&c->gamma2, sizeof(G2Affine))) {
    zkb_crs_free(ctx, c);
    return set_err(ctx, ZKB_ERR_ALLOC, "crs: cudaMalloc failed (G1 table %.1f MiB, G2 table %.1f MiB)",
                   (double)msm_windows(c->c1) * c->g1_cnt * 64 / 1048576.0, (double)msm_windows(c->c2) * c->g2_cnt * 128 / 1048576.0);
  }
  *out = c;
  return ZKB_OK;
}

static int crs_expand(zkb_ctx* ctx, zkb_crs* c, cudaStream_t st) {
  ZKB_TRY(expand_table_g1(ctx, c->g1, c->g1_cnt, c->g1_cnt, c->c1, st));
  ZKB_TRY(expand_table_g2(ctx, c->g2, c->g2_cnt, c->g2_cnt, c->c2, st));
  if (cudaStreamSynchronize(st) != cudaSuccess)
    return set_err(ctx, ZKB_ERR_CUDA, "crs: table expansion failed: %s", cudaGetErrorString(cudaGetLastError()));
  return ZKB_OK;
}

static int crs_upload_impl(zkb_ctx* ctx, const zkb_crs_host* h, int rank, int world, int layout, zkb_crs** out) {
  if (!ctx || !h || !out) return set_err(ctx, ZKB_ERR_ARG, "zkb_crs_upload: NULL argument");
  *out = nullptr;
  if (world < 1 || rank < 0 || rank >= world) return set_err(ctx, ZKB_ERR_ARG, "bad rank/world");
  if (h->n < 1) return set_err(ctx, ZKB_ERR_ARG, "crs.n must be >= 1");
  ZKB_CUDA(ctx, cudaSetDevice(ctx->device));
  zkb_crs* c;
  ZKB_TRY(crs_new(ctx, h->n, h->n_sum_gamma, h->n_sum_delta, rank, world, layout, &c));
  cudaStream_t st = ctx->stream;
  const size_t nxi = c->nxi(), nxt = c->nxt(), nsd = c->nsd();
  G1Affine* fx = c->g1 + c->off_fixed();
  if (!h->alpha1 || !h->beta1 || !h->delta1 || !h->beta2 || !h->gamma2 || !h->delta2 || !h->xi1 || !h->xi2 ||
      (h->n > 1 && !h->xi_t) || (h->n_sum_gamma && !h->sum_gamma) || (h->n_sum_delta && !h->sum_delta)) {
    zkb_crs_free(ctx, c);
    return set_err(ctx, ZKB_ERR_ARG, "zkb_crs_upload: NULL vector in zkb_crs_host");
  }
  cudaError_t ce = cudaSuccess;
  auto cp = [&](void* dst, const void* src, size_t bytes) {
    if (ce == cudaSuccess && bytes) ce = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st);
  };
  cp(fx + 0, h->alpha1, 64);
  cp(fx + 1, h->beta1, 64);
  cp(fx + 2, h->delta1, 64);
  cp(c->g2 + nxi + 0, h->beta2, 128);
  cp(c->g2 + nxi + 1, h->delta2, 128);
  cp(c->gamma2, h->gamma2, 128);
  if (layout == 1) {  // world runs of q consecutive entries: local [k1*q, (k1+1)*q) <-> global [m*k1 + rank*q, ...)
    const uint64_t m = h->n / world, q = m / world;
    for (uint64_t k1 = 0; k1 < (uint64_t)world; k1++) {
      const uint64_t g0 = m * k1 + (uint64_t)rank * q;
      const uint64_t cnt_t = k1 * q + q <= nxt ? q : (nxt > k1 * q ? nxt - k1 * q : 0);
      cp(c->g1 + k1 * q, h->xi1 + g0 * 8, q * 64);
      cp(c->g2 + k1 * q, h->xi2 + g0 * 16, q * 128);
      cp(c->g1 + c->off_xit() + k1 * q, h->xi_t + g0 * 8, cnt_t * 64);
    }
  } else {
    cp(c->g1, h->xi1 + c->xi_lo * 8, nxi * 64);
    cp(c->g1 + c->off_xit(), h->xi_t + c->xit_lo * 8, nxt * 64);
    cp(c->g2, h->xi2 + c->xi_lo * 16, nxi * 128);
  }
  cp(c->g1 + c->off_sd(), h->sum_delta + c->sd_lo * 8, nsd * 64);
  cp(c->sum_gamma, h->sum_gamma, c->n_sum_gamma * 64);
  int rc = ce == cudaSuccess ? ZKB_OK : set_err(ctx, ZKB_ERR_CUDA, "crs upload: copy failed: %s", cudaGetErrorString(ce));
  if (rc == ZKB_OK) rc = fq_to_mont(ctx, (Fq*)c->g1, c->g1_cnt * 2, true, st);
  if (rc == ZKB_OK) rc = fq_to_mont(ctx, (Fq*)c->g2, c->g2_cnt * 4, true, st);
  if (rc == ZKB_OK) rc = fq_to_mont(ctx, (Fq*)c->gamma2, 4, true, st);
  if (rc == ZKB_OK) rc = fq_to_mont(ctx, (Fq*)c->sum_gamma, c->n_sum_gamma * 2, true, st);
  // Raw coordinates from outside: every point must be the identity or on its curve, and every G2 point in the order-r
  // subgroup (the reference's SigmaG1 / SigmaG2 can only hold bn-constructed group elements, mod.rs:105-121; a point
  // outside the subgroup would make the proofs and the verifier's pairings meaningless).  ZKB_CRS_TRUSTED=1 skips the
  // subgroup scalar multiplications (one per G2 point: ~0.2 s at 2^20) for a CRS the caller produced itself.
  int bad = 0;
  if (rc == ZKB_OK) {
    void* p;
    rc = scratch_get(ctx, 9, sizeof(int), &p);
    int* d_bad = (int*)p;
    const char* tr = getenv("ZKB_CRS_TRUSTED");
    const bool sub = !(tr && atoi(tr) == 1);
    if (rc == ZKB_OK && cudaMemsetAsync(d_bad, 0, sizeof(int), st) != cudaSuccess) rc = set_err(ctx, ZKB_ERR_CUDA, "crs upload: memset failed");
    if (rc == ZKB_OK) rc = check_points_g1(ctx, c->g1, c->g1_cnt, d_bad, st);
    if (rc == ZKB_OK) rc = check_points_g1(ctx, c->sum_gamma, c->n_sum_gamma, d_bad, st);
    if (rc == ZKB_OK) rc = check_points_g2(ctx, c->g2, c->g2_cnt, sub, d_bad, st);
    if (rc == ZKB_OK) rc = check_points_g2(ctx, c->gamma2, 1, sub, d_bad, st);
    if (rc == ZKB_OK && cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess)
      rc = set_err(ctx, ZKB_ERR_CUDA, "crs upload: copy failed");
  }
  if (rc == ZKB_OK && cudaStreamSynchronize(st) != cudaSuccess)
    rc = set_err(ctx, ZKB_ERR_CUDA, "crs upload failed: %s", cudaGetErrorString(cudaGetLastError()));
  if (rc == ZKB_OK && (bad & 1)) rc = set_err(ctx, ZKB_ERR_ARG, "zkb_crs_upload: a CRS point is not on its curve");
  if (rc == ZKB_OK && (bad & 2)) rc = set_err(ctx, ZKB_ERR_ARG, "zkb_crs_upload: a G2 point of the CRS is not in the order-r subgroup");
  if (rc == ZKB_OK) rc = crs_expand(ctx, c, st);
  if (rc != ZKB_OK) { zkb_crs_free(ctx, c); return rc; }
  *out = c;
  return ZKB_OK;
}

int zkb_crs_upload(zkb_ctx* ctx, const zkb_crs_host* h, int rank, int world, zkb_crs** out) {
  return crs_upload_impl(ctx, h, rank, world, 0, out);
}
int zkb_crs_upload_shard(zkb_ctx* ctx, const zkb_comm* comm, const zkb_crs_host* h, zkb_crs** out) {
  if (!comm) return set_err(ctx, ZKB_ERR_ARG, "zkb_crs_upload_shard: NULL communicator");
  return crs_upload_impl(ctx, h, comm->rank, comm->world, 1, out);
}

static int setup_impl(zkb_ctx* ctx, const zkb_qap* q, const uint64_t* toxic, int rank, int world, int layout, zkb_crs** out) {
  if (!ctx || !q || !toxic || !out) return set_err(ctx, ZKB_ERR_ARG, "zkb_setup: NULL argument");
  *out = nullptr;
  ZKB_CUDA(ctx, cudaSetDevice(ctx->device));
  Fr alpha = fr_from_limbs(toxic), beta = fr_from_limbs(toxic + 4), gamma = fr_from_limbs(toxic + 8),
     delta = fr_from_limbs(toxic + 12), x = fr_from_limbs(toxic + 16);
  if (alpha.is_zero() || beta.is_zero() || gamma.is_zero() || delta.is_zero() || x.is_zero())
    return set_err(ctx, ZKB_ERR_DIV_ZERO, "setup: toxic values must be non-zero (random_elem, fr.rs:90-99)");
  const uint64_t n = q->n, m = q->m;
  zkb_crs* c;
  if (layout == 1 && q->generic) return set_err(ctx, ZKB_ERR_UNSUPPORTED, "sharded setup needs the roots-of-unity domain");
  ZKB_TRY(crs_new(ctx, n, q->n_input + 1, m - q->n_input - 1, rank, world, layout, &c));
  int rc;
  auto fail = [&](int code) { zkb_crs_free(ctx, c); return code; };
  cudaStream_t st = ctx->stream;
  const size_t nxi = c->nxi(), nxt = c->nxt(), nsd = c->nsd();
  // scalars
  void* p;
  if ((rc = scratch_get(ctx, 8, (n + n + m + 8 + 2 * (n / world + 1)) * sizeof(Fr), &p)) != ZKB_OK) return fail(rc);
  Fr* d_L = (Fr*)p;           // n   Lagrange basis at x; later reused for xi_t scalars
  Fr* d_pow = d_L + n;        // n   x^i
  Fr* d_lin = d_pow + n;      // m
  Fr* d_six = d_lin + m;      // alpha, beta, delta | beta, delta | gamma
  Fr* d_gat = d_six + 8;      // layout 1: this rank's xi scalars | xi_t scalars, gathered into local order
  if ((rc = scratch_get(ctx, 9, sizeof(int), &p)) != ZKB_OK) return fail(rc);
  int* d_bad = (int*)p;
  cudaMemsetAsync(d_bad, 0, sizeof(int), st);
  Fr tx;
  if (q->generic) {  // t(x) = prod (x - r_k)
    tx = Fr::one();
    for (const Fr& rk : q->h_roots) tx = tx * (x - rk);
  } else {
    tx = pow_u64(x, n) - Fr::one();  // t(x) = x^n - 1
  }
  Fr inv_delta = inverse(delta), inv_gamma = inverse(gamma);
  Fr tx_over_n = tx * inverse(fr_from_u64(n));
  Fr omega = q->generic ? Fr::one() : host_omega(q->log_n, false);
  auto launch_fail = [&](const char* what) { return fail(set_err(ctx, ZKB_ERR_CUDA, "setup: %s: %s", what, cudaGetErrorString(cudaGetLastError()))); };
  if (q->generic) k_lagrange_generic<<<cdiv(n, 128), 128, 0, st>>>(d_L, x, tx, q->d_roots, q->d_dinv, n, d_bad);
  else k_lagrange<<<cdiv(n, 128), 128, 0, st>>>(d_L, x, tx_over_n, omega, n, d_bad);
  ctx->launches++;
  if (cudaGetLastError() != cudaSuccess) return launch_fail("k_lagrange");
  k_lin<<<cdiv(m * 32, 256), 256, 0, st>>>(q->d_rptr[0], q->d_gate[0], q->d_rcoeff[0], q->d_rptr[1], q->d_gate[1],
                                           q->d_rcoeff[1], q->d_rptr[2], q->d_gate[2], q->d_rcoeff[2], d_L, m,
                                           q->n_input, alpha, beta, inv_gamma, inv_delta, d_lin);
  ctx->launches++;
  if (cudaGetLastError() != cudaSuccess) return launch_fail("k_lin");
  int bad = 0;
  cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, st);
  if (cudaStreamSynchronize(st) != cudaSuccess) return launch_fail("sync");
  if (bad) return fail(set_err(ctx, ZKB_ERR_UNSUPPORTED, "setup: x is one of the domain roots"));
  if ((rc = fill_powers(ctx, d_pow, x, Fr::one(), n, st)) != ZKB_OK) return fail(rc);
  if ((rc = fill_powers(ctx, d_L, x, tx * inv_delta, n, st)) != ZKB_OK) return fail(rc);  // xi_t scalars
  Fr six[6] = {alpha, beta, delta, beta, delta, gamma};
  cudaMemcpyAsync(d_six, six, sizeof six, cudaMemcpyHostToDevice, st);
  if ((rc = fixed_base_g1(ctx, c->g1 + c->off_fixed(), d_six, 3, st)) != ZKB_OK) return fail(rc);
  if ((rc = fixed_base_g2(ctx, c->g2 + nxi, d_six + 3, 2, st)) != ZKB_OK) return fail(rc);
  if ((rc = fixed_base_g2(ctx, c->gamma2, d_six + 5, 1, st)) != ZKB_OK) return fail(rc);
  const Fr *xi_scal = d_pow + c->xi_lo, *xit_scal = d_L + c->xit_lo;
  if (layout == 1) {
    const uint64_t ml = n / world, ql = ml / world;
    k_gather_S<<<cdiv(nxi, 256), 256, 0, st>>>(d_gat, d_pow, (uint64_t)rank, ql, ml, nxi);
    if (nxt) k_gather_S<<<cdiv(nxt, 256), 256, 0, st>>>(d_gat + ml, d_L, (uint64_t)rank, ql, ml, nxt);
    ctx->launches += 2;
    if (cudaGetLastError() != cudaSuccess) return launch_fail("k_gather_S");
    xi_scal = d_gat; xit_scal = d_gat + ml;
  }
  if ((rc = fixed_base_g1(ctx, c->g1, xi_scal, nxi, st)) != ZKB_OK) return fail(rc);
  if ((rc = fixed_base_g2(ctx, c->g2, xi_scal, nxi, st)) != ZKB_OK) return fail(rc);
  if ((rc = fixed_base_g1(ctx, c->g1 + c->off_xit(), xit_scal, nxt, st)) != ZKB_OK) return fail(rc);
  if ((rc = fixed_base_g1(ctx, c->sum_gamma, d_lin, c->n_sum_gamma, st)) != ZKB_OK) return fail(rc);
  if ((rc = fixed_base_g1(ctx, c->g1 + c->off_sd(), d_lin + c->n_sum_gamma + c->sd_lo, nsd, st)) != ZKB_OK) return fail(rc);
  if (cudaStreamSynchronize(st) != cudaSuccess) return launch_fail("fixed-base");
  if ((rc = crs_expand(ctx, c, st)) != ZKB_OK) return fail(rc);
  *out = c;
  return ZKB_OK;
}

int zkb_setup(zkb_ctx* ctx, const zkb_qap* q, const uint64_t* toxic, int rank, int world, zkb_crs** out) {
  if (world < 1 || rank < 0 || rank >= world) return set_err(ctx, ZKB_ERR_ARG, "bad rank/world");
  return setup_impl(ctx, q, toxic, rank, world, 0, out);
}
int zkb_setup_shard(zkb_ctx* ctx, const zkb_comm* comm, const zkb_qap* q, const uint64_t* toxic, zkb_crs** out) {
  if (!comm) return set_err(ctx, ZKB_ERR_ARG, "zkb_setup_shard: NULL communicator");
  return setup_impl(ctx, q, toxic, comm->rank, comm->world, 1, out);
}

int zkb_crs_dims(const zkb_crs* c, uint64_t* n, uint64_t* nsg, uint64_t* nsd) {
  if (!c) return ZKB_ERR_ARG;
  if (n) *n = c->n;
  if (nsg) *nsg = c->n_sum_gamma;
  if (nsd) *nsd = c->n_sum_delta;
  return ZKB_OK;
}

int zkb_crs_download(zkb_ctx* ctx, const zkb_crs* c, zkb_crs_host* d) {
  if (!ctx || !c || !d) return set_err(ctx, ZKB_ERR_ARG, "zkb_crs_download: NULL argument");
  if (c->world != 1) return set_err(ctx, ZKB_ERR_UNSUPPORTED, "crs download needs an unsharded CRS");
  ZKB_CUDA(ctx, cudaSetDevice(ctx->device));
  d->n = c->n; d->n_sum_gamma = c->n_sum_gamma; d->n_sum_delta = c->n_sum_delta;
  const G1Affine* fx = c->g1 + c->off_fixed();
  ZKB_TRY(download_fq(ctx, (uint64_t*)d->alpha1, fx + 0, 2));
  ZKB_TRY(download_fq(ctx, (uint64_t*)d->beta1, fx + 1, 2));
  ZKB_TRY(download_fq(ctx, (uint64_t*)d->delta1, fx + 2, 2));
  ZKB_TRY(download_fq(ctx, (uint64_t*)d->beta2, c->g2 + c->n + 0, 4));
  ZKB_TRY(download_fq(ctx, (uint64_t*)d->delta2, c->g2 + c->n + 1, 4));
  ZKB_TRY(download_fq(ctx, (uint64_t*)d->gamma2, c->gamma2, 4));
  ZKB_TRY(download_fq(ctx, (uint64_t*)d->xi1, c->g1, c->n * 2));
  ZKB_TRY(download_fq(ctx, (uint64_t*)d->xi_t, c->g1 + c->off_xit(), (c->n - 1) * 2));
  ZKB_TRY(download_fq(ctx, (uint64_t*)d->sum_gamma, c->sum_gamma, c->n_sum_gamma * 2));
  ZKB_TRY(download_fq(ctx, (uint64_t*)d->sum_delta, c->g1 + c->off_sd(), c->n_sum_delta * 2));
  ZKB_TRY(download_fq(ctx, (uint64_t*)d->xi2, c->g2, c->n * 4));
  return ZKB_OK;
}

}  // extern "C"
