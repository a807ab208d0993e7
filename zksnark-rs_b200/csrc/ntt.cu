// Radix-2 number-theoretic transform over BN254 Fr for sm_100a.
//
// Replaces the polynomial arithmetic behind `h = (u_sum * v_sum - w_sum) / t`
// (/root/reference/src/groth16/mod.rs:277): CoefficientPoly::Mul (schoolbook, coefficient_poly.rs:
// 93-130) and polynomial_division (field/mod.rs:428-469).  Transform convention is the reference's
// dft/idft (field/mod.rs:508-537): X[i] = sum_j x[j] * root^(i*j), idft scales by 1/n.
//
// Structure: a size-2^k transform is split into passes of up to 10 butterfly stages; each pass
// stages a 1024-element tile (32 KB) in shared memory as two 16-byte planes (conflict-free
// LDS.128), runs its stages there, and writes the tile back -- so HBM sees one read and one write
// of the vector per pass (2-3 passes for 2^20..2^26).  Tiles of the strided passes are
// (2^B rows) x (C = 2^(10-B) adjacent columns) so that every global access is a >=128-byte run.
// Twiddles omega^k (k < n/2) live in a per-size table (L2-resident: 16 MB at 2^20).
#include "common.cuh"

namespace zkb {

static const uint32_t kOmega28[8] = ZKB_FR_OMEGA28;
static const uint32_t kOmega28Inv[8] = ZKB_FR_OMEGA28_INV;

Fr host_omega(uint32_t log_n, bool inverse) {
  Fr w;
  for (int i = 0; i < 8; i++) w.v[i] = inverse ? kOmega28Inv[i] : kOmega28[i];
  for (uint32_t i = log_n; i < 28; i++) w = w * w;
  return w;
}

// ------------------------------------------------------------------------------------------------
__global__ void k_fill_powers(Fr* out, Fr base, Fr first, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = first * pow_u64(base, (uint64_t)i);
}

__global__ void k_scale_powers(Fr* d, Fr base, Fr first, uint32_t log_n, int bitrev) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t n = (size_t)1 << log_n;
  if (i >= n) return;
  uint64_t e = bitrev ? (uint64_t)(__brevll((unsigned long long)i) >> (64 - log_n)) : (uint64_t)i;
  if (log_n == 0) e = 0;
  d[i] = d[i] * (first * pow_u64(base, e));
}

__global__ void k_to_mont(Fr* d, size_t n, int to) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  d[i] = to ? to_mont(d[i]) : from_mont(d[i]);
}

__global__ void k_vec_mul(Fr* out, const Fr* a, const Fr* b, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = a[i] * b[i];
}

__global__ void k_bitrev(Fr* out, const Fr* in, uint32_t log_n, Fr scale, int has_scale) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t n = (size_t)1 << log_n;
  if (i >= n) return;
  size_t j = log_n ? (size_t)(__brevll((unsigned long long)i) >> (64 - log_n)) : 0;
  Fr v = in[i];
  if (has_scale) v = v * scale;
  out[j] = v;
}

int get_twiddles(zkb_ctx* ctx, uint32_t log_n, bool inverse, Fr** out) {
  if (log_n < 1 || log_n > 27) return set_err(ctx, ZKB_ERR_ARG, "ntt size 2^%u unsupported", log_n);
  Fr*& t = ctx->tw[log_n][inverse ? 1 : 0];
  if (!t) {
    size_t half = (size_t)1 << (log_n - 1);
    ZKB_CUDA(ctx, cudaMalloc(&t, half * sizeof(Fr)));
    ZKB_LAUNCH(ctx, k_fill_powers, cdiv(half, 256), 256, 0, ctx->stream, t, host_omega(log_n, inverse), Fr::one(), half);
    ZKB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  *out = t;
  return ZKB_OK;
}

// ------------------------------------------------------------------------------------------------
// one pass = stages [lo, hi) of the size-2^log_n transform on 2^(hi-lo+logC)-element tiles.
// DIF (Gentleman-Sande) runs the stages downwards, DIT (Cooley-Tukey) upwards.
struct alignas(16) Half { uint32_t v[4]; };

template <bool DIT>
__global__ void __launch_bounds__(512) k_ntt_pass(Fr* __restrict__ d, const Fr* __restrict__ tw, uint32_t log_n,
                                                  uint32_t hi, uint32_t lo, uint32_t logC) {
  extern __shared__ uint4 smem[];
  const uint32_t B = hi - lo;
  const uint32_t logT = B + logC;
  const uint32_t T = 1u << logT;
  uint4* p0 = smem;       // limbs 0..3 of tile element e
  uint4* p1 = smem + T;   // limbs 4..7
  const uint32_t t = threadIdx.x;
  const uint32_t C = 1u << logC;
  const uint64_t groups_per_H = ((uint64_t)1 << lo) >> logC;
  const uint64_t H = blockIdx.x / groups_per_H;
  const uint64_t Lbase = (blockIdx.x % groups_per_H) << logC;
  const uint64_t base = (H << hi) + Lbase;
  const uint4* g4 = reinterpret_cast<const uint4*>(d);
  uint4* gw4 = reinterpret_cast<uint4*>(d);

  // load: thread handles tile elements t and t + T/2
#pragma unroll
  for (int r = 0; r < 2; r++) {
    uint32_t e = t + r * (T >> 1);
    uint64_t gi = base + ((uint64_t)(e >> logC) << lo) + (e & (C - 1));
    p0[e] = g4[gi * 2];
    p1[e] = g4[gi * 2 + 1];
  }
  __syncthreads();

  for (uint32_t k = 0; k < B; k++) {
    const uint32_t ls = DIT ? k : (B - 1 - k);  // local stage
    const uint32_t s = lo + ls;
    const uint32_t sh = ls + logC;              // log2 of the half distance in tile units
    const uint32_t half = 1u << sh;
    const uint32_t e0 = ((t >> sh) << (sh + 1)) | (t & (half - 1));
    const uint32_t e1 = e0 + half;
    // twiddle exponent: (global index of e0 mod 2^s) << (log_n - 1 - s)
    const uint64_t m0 = (e0 >> logC) & ((1u << ls) - 1);
    const uint64_t j = (m0 << lo) + Lbase + (e0 & (C - 1));
    const uint64_t jj = (s == 0) ? 0 : (j & (((uint64_t)1 << s) - 1));
    const Fr w = tw[jj << (log_n - 1 - s)];
    Fr a, b;
    *reinterpret_cast<uint4*>(&a.v[0]) = p0[e0];
    *reinterpret_cast<uint4*>(&a.v[4]) = p1[e0];
    *reinterpret_cast<uint4*>(&b.v[0]) = p0[e1];
    *reinterpret_cast<uint4*>(&b.v[4]) = p1[e1];
    Fr x, y;
    if (DIT) {
      Fr wb = w * b;
      x = a + wb;
      y = a - wb;
    } else {
      x = a + b;
      y = (a - b) * w;
    }
    p0[e0] = *reinterpret_cast<uint4*>(&x.v[0]);
    p1[e0] = *reinterpret_cast<uint4*>(&x.v[4]);
    p0[e1] = *reinterpret_cast<uint4*>(&y.v[0]);
    p1[e1] = *reinterpret_cast<uint4*>(&y.v[4]);
    __syncthreads();
  }

#pragma unroll
  for (int r = 0; r < 2; r++) {
    uint32_t e = t + r * (T >> 1);
    uint64_t gi = base + ((uint64_t)(e >> logC) << lo) + (e & (C - 1));
    gw4[gi * 2] = p0[e];
    gw4[gi * 2 + 1] = p1[e];
  }
}

struct Pass { uint32_t lo, hi, logC; };

static std::vector<Pass> plan(uint32_t log_n) {
  std::vector<Pass> ps;  // ordered from the low stages up
  uint32_t l0 = log_n < 10 ? log_n : 10;
  ps.push_back({0, l0, 0});
  uint32_t rem = log_n - l0;
  if (rem) {
    uint32_t np = (rem + 7) / 8;
    uint32_t lo = l0;
    for (uint32_t i = 0; i < np; i++) {
      uint32_t b = rem / np + (i < rem % np ? 1 : 0);
      ps.push_back({lo, lo + b, 10 - b});
      lo += b;
    }
  }
  return ps;
}

template <bool DIT>
static int run_ntt(zkb_ctx* ctx, Fr* d, uint32_t log_n, bool inverse, cudaStream_t st) {
  if (log_n == 0) return ZKB_OK;
  Fr* tw;
  ZKB_TRY(get_twiddles(ctx, log_n, inverse, &tw));
  std::vector<Pass> ps = plan(log_n);
  int np = (int)ps.size();
  for (int q = 0; q < np; q++) {
    const Pass& p = DIT ? ps[q] : ps[np - 1 - q];
    uint32_t logT = p.hi - p.lo + p.logC;
    uint32_t T = 1u << logT;
    unsigned grid = (unsigned)(((size_t)1 << log_n) >> logT);
    unsigned block = T / 2;
    size_t smem = (size_t)T * 32;
    if (ctx->profile) ctx->prof_units[PK_NTT] += (uint64_t)1 << log_n;
    ZKB_LAUNCH_K(ctx, PK_NTT, k_ntt_pass<DIT>, grid, block, smem, st, d, tw, log_n, p.hi, p.lo, p.logC);
  }
  return ZKB_OK;
}

int ntt_dif(zkb_ctx* ctx, Fr* d, uint32_t log_n, bool inverse, cudaStream_t st) {
  return run_ntt<false>(ctx, d, log_n, inverse, st);
}
int ntt_dit(zkb_ctx* ctx, Fr* d, uint32_t log_n, bool inverse, cudaStream_t st) {
  return run_ntt<true>(ctx, d, log_n, inverse, st);
}

int bitrev_permute(zkb_ctx* ctx, Fr* out, const Fr* in, uint32_t log_n, const Fr* h_scale, cudaStream_t st) {
  size_t n = (size_t)1 << log_n;
  Fr sc = h_scale ? *h_scale : Fr::one();
  ZKB_LAUNCH(ctx, k_bitrev, cdiv(n, 256), 256, 0, st, out, in, log_n, sc, h_scale ? 1 : 0);
  return ZKB_OK;
}

int scale_powers(zkb_ctx* ctx, Fr* d, uint32_t log_n, const Fr& base, const Fr& first, bool bitrev, cudaStream_t st) {
  size_t n = (size_t)1 << log_n;
  ZKB_LAUNCH(ctx, k_scale_powers, cdiv(n, 256), 256, 0, st, d, base, first, log_n, bitrev ? 1 : 0);
  return ZKB_OK;
}

int vec_to_mont(zkb_ctx* ctx, Fr* d, size_t n, bool to, cudaStream_t st) {
  if (!n) return ZKB_OK;
  ZKB_LAUNCH(ctx, k_to_mont, cdiv(n, 256), 256, 0, st, d, n, to ? 1 : 0);
  return ZKB_OK;
}

int vec_mul(zkb_ctx* ctx, Fr* out, const Fr* a, const Fr* b, size_t n, cudaStream_t st) {
  if (!n) return ZKB_OK;
  ZKB_LAUNCH(ctx, k_vec_mul, cdiv(n, 256), 256, 0, st, out, a, b, n);
  return ZKB_OK;
}

int fill_powers(zkb_ctx* ctx, Fr* out, const Fr& base, const Fr& first, size_t n, cudaStream_t st) {
  if (!n) return ZKB_OK;
  ZKB_LAUNCH(ctx, k_fill_powers, cdiv(n, 256), 256, 0, st, out, base, first, n);
  return ZKB_OK;
}

}  // namespace zkb
