// Pippenger bucket multi-scalar multiplication over BN254 G1 / G2 for sm_100a.
//
// Replaces the reference's MSM pattern `coeffs.zip(points).map(exp_encrypted_g*).sum()`
// (/root/reference/src/groth16/mod.rs:255-272, 279-290), i.e. n independent MSB-first
// double-and-add scalar multiplications (fr.rs:114-119) folded sequentially (fr.rs:191-223).
// Group addition is exact, so any summation order gives the same affine result.
//
// Pipeline (all on one stream, no host round trips):
//   1. k_digits_count   signed-digit recoding of every scalar into W windows of c bits
//                       (digits in [-2^(c-1), 2^(c-1)]), histogram of bucket sizes
//   2. scan             exclusive prefix sum of the W * 2^(c-1) bucket sizes
//   3. k_digits_scatter counting-sort the (point index, sign) records by bucket
//   4. k_accumulate_chunks  one thread per 32 sorted records: XYZZ mixed additions, flush at bucket
//                       boundaries; k_fix_heads: per-warp fold of buckets that span chunks (shuffle tree)
//   5. k_reduce_chunks  per window, running-sum reduction of L-bucket chunks
//   6. k_window_finish  per window: sum_t (T_t + v0_t * S_t) with a shared-memory tree
//   7. k_combine        Horner over the windows (c doublings per window)
#pragma once
#include "common.cuh"

namespace zkb {

// ------------------------------------------------------------------------------------------------
// signed-digit recoding.  k: canonical 254-bit scalar.  Calls f(j, digit) for every window.
struct DigitPlan {
  int c;        // window bits
  int W;        // number of windows = floor(254 / c) + 1
  uint32_t nb;  // buckets per window = 2^(c-1); bucket value v in 1..nb
};

__host__ __device__ inline DigitPlan make_plan(int c) {
  DigitPlan p;
  p.c = c;
  p.W = 254 / c + 1;
  p.nb = 1u << (c - 1);
  return p;
}

__device__ __forceinline__ uint32_t get_bits(const uint32_t k[8], int pos, int c) {
  // bits [pos, pos+c) of the 256-bit integer k (c <= 24)
  int w = pos >> 5, o = pos & 31;
  if (w >= 8) return 0;
  uint64_t lo = k[w];
  uint64_t hi = (w + 1 < 8) ? k[w + 1] : 0;
  uint64_t v = (lo | (hi << 32)) >> o;
  return (uint32_t)v & ((1u << c) - 1);
}

__device__ __forceinline__ Fr load_scalar(const Fr* scalars, size_t i, int mont) {
  Fr k = scalars[i];
  if (mont) k = from_mont(k);
  return k;
}

static __global__ void k_digits_count(const Fr* __restrict__ scalars, int mont, size_t n, DigitPlan pl,
                               uint32_t* __restrict__ hist) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr k = load_scalar(scalars, i, mont);
  uint32_t carry = 0;
  for (int j = 0; j < pl.W; j++) {
    uint32_t d = get_bits(k.v, j * pl.c, pl.c) + carry;
    carry = d > pl.nb;
    uint32_t mag = carry ? ((1u << pl.c) - d) : d;
    if (mag) atomicAdd(&hist[(size_t)j * pl.nb + mag - 1], 1u);
  }
}

static __global__ void k_digits_scatter(const Fr* __restrict__ scalars, int mont, size_t n, DigitPlan pl,
                                 uint32_t* __restrict__ cursor, uint32_t* __restrict__ sorted) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr k = load_scalar(scalars, i, mont);
  uint32_t carry = 0;
  for (int j = 0; j < pl.W; j++) {
    uint32_t d = get_bits(k.v, j * pl.c, pl.c) + carry;
    carry = d > pl.nb;
    uint32_t mag = carry ? ((1u << pl.c) - d) : d;
    if (mag) {
      uint32_t pos = atomicAdd(&cursor[(size_t)j * pl.nb + mag - 1], 1u);
      sorted[pos] = (uint32_t)i | (carry << 31);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// exclusive scan of uint32 (three small kernels; total <= 2^24 entries)
static const int SCAN_B = 1024;  // elements per block (256 threads x 4)

static __global__ void k_scan_block(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, uint32_t* __restrict__ sums,
                             size_t n) {
  __shared__ uint32_t sh[256];
  size_t base = (size_t)blockIdx.x * SCAN_B + threadIdx.x * 4;
  uint32_t v[4], tot = 0;
#pragma unroll
  for (int q = 0; q < 4; q++) {
    v[q] = base + q < n ? in[base + q] : 0;
    tot += v[q];
  }
  sh[threadIdx.x] = tot;
  __syncthreads();
  for (int off = 1; off < 256; off <<= 1) {
    uint32_t x = threadIdx.x >= off ? sh[threadIdx.x - off] : 0;
    __syncthreads();
    sh[threadIdx.x] += x;
    __syncthreads();
  }
  uint32_t excl = sh[threadIdx.x] - tot;
#pragma unroll
  for (int q = 0; q < 4; q++) {
    if (base + q < n) out[base + q] = excl;
    excl += v[q];
  }
  if (threadIdx.x == 255) sums[blockIdx.x] = sh[255];
}

static __global__ void k_scan_sums(uint32_t* sums, size_t nblocks, uint32_t* total) {
  // single block, sequential over chunks of 1024
  __shared__ uint32_t sh[1024];
  __shared__ uint32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (size_t base = 0; base < nblocks; base += 1024) {
    size_t i = base + threadIdx.x;
    uint32_t v = i < nblocks ? sums[i] : 0;
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
      uint32_t x = threadIdx.x >= off ? sh[threadIdx.x - off] : 0;
      __syncthreads();
      sh[threadIdx.x] += x;
      __syncthreads();
    }
    if (i < nblocks) sums[i] = carry + sh[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry += sh[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}

static __global__ void k_scan_add(uint32_t* __restrict__ out, uint32_t* __restrict__ out2, const uint32_t* __restrict__ sums,
                           size_t n, const uint32_t* total) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    uint32_t v = out[i] + sums[i / SCAN_B];
    out[i] = v;
    out2[i] = v;
  }
  if (i == n) out[n] = *total;  // offsets[n] = total
}

// ------------------------------------------------------------------------------------------------
// Bucket accumulation, load-balanced: the sorted record array is cut into chunks of S records and
// every thread sums exactly one chunk with XYZZ mixed additions, flushing at bucket boundaries.
// A bucket that begins inside the chunk is written to buckets[g]; the leading piece of a bucket
// that began in an earlier chunk goes to heads[t] and is folded in by
// k_fix_heads.  buckets[] is zero-filled (= identity) beforehand, empty buckets are never touched.

template <class F, int S>
__global__ void __launch_bounds__(128) k_accumulate_chunks(const Affine<F>* __restrict__ pts, const uint32_t* __restrict__ offs,
                                                           const uint32_t* __restrict__ sorted, uint32_t nbk, size_t nchunks,
                                                           XYZZ<F>* __restrict__ buckets, XYZZ<F>* __restrict__ heads) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nchunks) return;
  const uint32_t total = offs[nbk];
  const size_t start64 = t * (size_t)S;
  if (start64 >= total) return;
  const uint32_t start = (uint32_t)start64;
  const uint32_t end = (total - start > (uint32_t)S) ? start + S : total;
  // g: offs[g] <= start < offs[g+1]
  uint32_t lo = 0, hi = nbk;
  while (lo < hi) {
    uint32_t mid = (lo + hi) >> 1;
    if (offs[mid] <= start) lo = mid + 1; else hi = mid;
  }
  uint32_t g = lo - 1;
  bool is_head = offs[g] < start;
  uint32_t bend = offs[g + 1];
  XYZZ<F> acc = XYZZ<F>::inf();
  for (uint32_t p = start; p < end; p++) {
    if (p == bend) {
      if (is_head) { heads[t] = acc; is_head = false; } else buckets[g] = acc;
      acc = XYZZ<F>::inf();
      do { g++; bend = offs[g + 1]; } while (bend <= p);
    }
    uint32_t rec = sorted[p];
    Affine<F> P = pts[rec & 0x7fffffffu];
    if (rec >> 31) P = neg(P);
    acc = madd(acc, P);
  }
  if (is_head) heads[t] = acc; else buckets[g] = acc;
}

template <class F>
__device__ __forceinline__ XYZZ<F> shfl_down_xyzz(const XYZZ<F>& p, int off) {
  XYZZ<F> r;
  const uint32_t* s = reinterpret_cast<const uint32_t*>(&p);
  uint32_t* d = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(XYZZ<F>) / 4); i++) d[i] = __shfl_down_sync(0xffffffffu, s[i], off);
  return r;
}

template <class F>
__device__ __forceinline__ XYZZ<F> shfl_xyzz(const XYZZ<F>& p, int src) {
  XYZZ<F> r;
  const uint32_t* s = reinterpret_cast<const uint32_t*>(&p);
  uint32_t* d = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(XYZZ<F>) / 4); i++) d[i] = __shfl_sync(0xffffffffu, s[i], src);
  return r;
}

// Fold the head pieces into their buckets.  One thread per bucket g: the chunks whose first record
// lies strictly inside bucket g are t with offs[g] < t*S < offs[g+1] (computed from the offsets, so
// no search).  Short runs (the common case: ~1 head per bucket when the mean bucket size is about
// S) are summed by the owning thread; long runs (skewed scalars: one huge bucket) are summed by
// the whole warp, lanes striding over the run followed by a shuffle tree.
template <class F, int S>
__global__ void __launch_bounds__(128) k_fix_heads(const uint32_t* __restrict__ offs, uint32_t nbk,
                                                   XYZZ<F>* __restrict__ buckets, const XYZZ<F>* __restrict__ heads) {
  const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  uint32_t t0 = 1, t1 = 0;
  if (g < nbk) {
    const uint32_t lo = offs[g], hi = offs[g + 1];
    if (hi > lo) { t0 = lo / S + 1; t1 = (hi - 1) / S; }
  }
  const uint32_t cnt = t1 >= t0 ? t1 - t0 + 1 : 0;
  const bool big = cnt > 8;
  XYZZ<F> acc = XYZZ<F>::inf();
  if (!big)
    for (uint32_t t = t0; t <= t1; t++) acc = add(acc, heads[t]);
  unsigned todo = __ballot_sync(0xffffffffu, big);
  while (todo) {
    const int src = __ffs(todo) - 1;
    todo &= todo - 1;
    const uint32_t b0 = __shfl_sync(0xffffffffu, t0, src), b1 = __shfl_sync(0xffffffffu, t1, src);
    XYZZ<F> part = XYZZ<F>::inf();
    for (uint32_t t = b0 + lane; t <= b1; t += 32) part = add(part, heads[t]);
    for (int off = 16; off > 0; off >>= 1) {
      XYZZ<F> o = shfl_down_xyzz(part, off);
      part = add(part, o);
    }
    XYZZ<F> tot = shfl_xyzz(part, 0);
    if (lane == src) acc = tot;
  }
  if (cnt) buckets[g] = add(buckets[g], acc);
}

// chunk t of window j covers bucket values v0+1 .. v0+L (v0 = t*L).  S = sum B_v, T = sum (v - v0) B_v.
template <class F>
__global__ void __launch_bounds__(128) k_reduce_chunks(const XYZZ<F>* __restrict__ buckets, uint32_t nb, uint32_t L,
                                                       size_t nchunks_total, XYZZ<F>* __restrict__ S,
                                                       XYZZ<F>* __restrict__ T) {
  size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= nchunks_total) return;
  uint32_t per_win = nb / L;
  size_t j = g / per_win;
  uint32_t t = (uint32_t)(g % per_win);
  const XYZZ<F>* b = buckets + j * nb + (size_t)t * L;
  XYZZ<F> run = XYZZ<F>::inf(), acc = XYZZ<F>::inf();
  for (int v = (int)L - 1; v >= 0; v--) {
    run = add(run, b[v]);
    acc = add(acc, run);
  }
  S[g] = run;
  T[g] = acc;
}

template <class F>
__device__ XYZZ<F> small_mul(const XYZZ<F>& p, uint32_t k) {
  XYZZ<F> acc = XYZZ<F>::inf();
  if (k == 0 || p.is_inf()) return acc;
  int top = 31 - __clz(k);
  for (int b = top; b >= 0; b--) {
    acc = dbl(acc);
    if ((k >> b) & 1u) acc = add(acc, p);
  }
  return acc;
}

// one block per window: X_t = T_t + (t*L) * S_t, then tree-sum over t.
template <class F>
__global__ void __launch_bounds__(128) k_window_finish(const XYZZ<F>* __restrict__ S, const XYZZ<F>* __restrict__ T,
                                                       uint32_t per_win, uint32_t L, XYZZ<F>* __restrict__ wsum) {
  extern __shared__ uint4 smem_raw[];
  XYZZ<F>* sh = reinterpret_cast<XYZZ<F>*>(smem_raw);
  size_t j = blockIdx.x;
  XYZZ<F> acc = XYZZ<F>::inf();
  for (uint32_t t = threadIdx.x; t < per_win; t += blockDim.x) {
    XYZZ<F> x = add(T[j * per_win + t], small_mul(S[j * per_win + t], t * L));
    acc = add(acc, x);
  }
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (uint32_t off = blockDim.x >> 1; off > 0; off >>= 1) {
    if (threadIdx.x < off) sh[threadIdx.x] = add(sh[threadIdx.x], sh[threadIdx.x + off]);
    __syncthreads();
  }
  if (threadIdx.x == 0) wsum[j] = sh[0];
}

template <class F>
__global__ void k_combine(const XYZZ<F>* __restrict__ wsum, int W, int c, XYZZ<F>* __restrict__ out) {
  if (threadIdx.x | blockIdx.x) return;
  XYZZ<F> acc = wsum[W - 1];
  for (int j = W - 2; j >= 0; j--) {
    for (int q = 0; q < c; q++) acc = dbl(acc);
    acc = add(acc, wsum[j]);
  }
  *out = acc;
}

template <class F>
__global__ void k_set_inf(XYZZ<F>* out) {
  if (threadIdx.x | blockIdx.x) return;
  *out = XYZZ<F>::inf();
}

static int pick_c(size_t n) {
  // minimise W * (10 n + 2 * 14 * 2 * 2^(c-1)) modmuls: N*W mixed adds (10M) + bucket reduction
  // (2 full adds of 14M per bucket, weighted x2 for its lower parallelism)
  int best = 4;
  double best_cost = 1e300;
  for (int c = 4; c <= 18; c++) {
    double W = 254 / c + 1;
    double cost = W * (10.0 * (double)n + 56.0 * (double)((size_t)1 << (c - 1)));
    if (cost < best_cost) { best_cost = cost; best = c; }
  }
  return best;
}

template <class F>
static int msm_impl(zkb_ctx* ctx, const Affine<F>* pts, const Fr* scalars, bool mont, size_t n, int c, XYZZ<F>* d_out,
                    int slot, cudaStream_t st) {
  if (n == 0) {
    ZKB_LAUNCH(ctx, k_set_inf<F>, 1, 1, 0, st, d_out);
    return ZKB_OK;
  }
  if (n >= ((size_t)1 << 31)) return set_err(ctx, ZKB_ERR_ARG, "msm: n too large");
  if (c <= 0) c = pick_c(n);
  if (c < 2 || c > 20) return set_err(ctx, ZKB_ERR_ARG, "msm: window_bits %d out of range [2,20]", c);
  DigitPlan pl = make_plan(c);
  size_t nbk = (size_t)pl.W * pl.nb;
  uint32_t L = pl.nb >= 64 ? 16 : (pl.nb >= 8 ? 4 : 1);  // chunk length for the running-sum reduction
  uint32_t per_win = pl.nb / L;
  size_t nchunks = (size_t)pl.W * per_win;
  size_t nscan_blocks = (nbk + SCAN_B - 1) / SCAN_B;

  uint32_t *hist, *offs, *cursor, *sums, *sorted;
  XYZZ<F>*buckets, *S, *T, *wsum;
  void* p;
  // layout of the u32 scratch: hist[nbk] | offs[nbk+1] | cursor[nbk] | sums[nscan_blocks+1]
  size_t u32_words = nbk * 3 + 1 + nscan_blocks + 8;
  ZKB_TRY(scratch_get(ctx, slot + 0, u32_words * 4, &p));
  hist = (uint32_t*)p;
  offs = hist + nbk;
  cursor = offs + nbk + 1;
  sums = cursor + nbk;
  ZKB_TRY(scratch_get(ctx, slot + 1, (size_t)pl.W * n * 4, &p));
  sorted = (uint32_t*)p;
  const int ACC_S = 32;  // records per accumulation chunk
  size_t nacc = ((size_t)pl.W * n + ACC_S - 1) / ACC_S;
  ZKB_TRY(scratch_get(ctx, slot + 2, (nbk + 2 * nchunks + pl.W + nacc) * sizeof(XYZZ<F>), &p));
  buckets = (XYZZ<F>*)p;
  S = buckets + nbk;
  T = S + nchunks;
  wsum = T + nchunks;
  XYZZ<F>* heads = wsum + pl.W;

  ZKB_CUDA(ctx, cudaMemsetAsync(hist, 0, nbk * 4, st));
  ZKB_LAUNCH(ctx, k_digits_count, cdiv(n, 256), 256, 0, st, scalars, mont ? 1 : 0, n, pl, hist);
  ZKB_LAUNCH(ctx, k_scan_block, (unsigned)nscan_blocks, 256, 0, st, hist, offs, sums, nbk);
  ZKB_LAUNCH(ctx, k_scan_sums, 1, 1024, 0, st, sums, nscan_blocks, sums + nscan_blocks);
  ZKB_LAUNCH(ctx, k_scan_add, cdiv(nbk + 1, 256), 256, 0, st, offs, cursor, sums, nbk, sums + nscan_blocks);
  ZKB_LAUNCH(ctx, k_digits_scatter, cdiv(n, 256), 256, 0, st, scalars, mont ? 1 : 0, n, pl, cursor, sorted);
  ZKB_CUDA(ctx, cudaMemsetAsync(buckets, 0, nbk * sizeof(XYZZ<F>), st));  // all-zero XYZZ = identity
  ZKB_LAUNCH_K(ctx, sizeof(F) == sizeof(Fq) ? PK_ACC_G1 : PK_ACC_G2, (k_accumulate_chunks<F, ACC_S>), cdiv(nacc, 128), 128, 0, st,
               pts, offs, sorted, (uint32_t)nbk, nacc, buckets, heads);
  if (ctx->profile) ctx->prof_units[sizeof(F) == sizeof(Fq) ? PK_ACC_G1 : PK_ACC_G2] += (uint64_t)pl.W * n;
  ZKB_LAUNCH(ctx, (k_fix_heads<F, ACC_S>), cdiv(nbk, 128), 128, 0, st, offs, (uint32_t)nbk, buckets, heads);
  ZKB_LAUNCH(ctx, k_reduce_chunks<F>, cdiv(nchunks, 128), 128, 0, st, buckets, pl.nb, L, nchunks, S, T);
  unsigned fin_threads = per_win >= 128 ? 128 : (per_win >= 32 ? 32 : 1);
  // round per_win down to a power of two thread count (per_win is a power of two)
  ZKB_LAUNCH(ctx, k_window_finish<F>, pl.W, fin_threads, fin_threads * sizeof(XYZZ<F>), st, S, T, per_win, L, wsum);
  ZKB_LAUNCH(ctx, k_combine<F>, 1, 32, 0, st, wsum, pl.W, pl.c, d_out);
  return ZKB_OK;
}

// ------------------------------------------------------------------------------------------------
// fixed-base scalar multiplication: out[i] = k_i * base, affine (used by setup / bench scaffolding;
// replaces encrypt_g1 / encrypt_g2, fr.rs:106-113)
__device__ __forceinline__ G1Affine g1_base() {
  const uint32_t bx[8] = ZKB_G1_BASE_X, by[8] = ZKB_G1_BASE_Y;
  G1Affine b;
#pragma unroll
  for (int i = 0; i < 8; i++) { b.x.v[i] = bx[i]; b.y.v[i] = by[i]; }
  return b;
}
__device__ __forceinline__ G2Affine g2_base() {
  const uint32_t x0[8] = ZKB_G2_BASE_X0, x1[8] = ZKB_G2_BASE_X1, y0[8] = ZKB_G2_BASE_Y0, y1[8] = ZKB_G2_BASE_Y1;
  G2Affine b;
#pragma unroll
  for (int i = 0; i < 8; i++) { b.x.c0.v[i] = x0[i]; b.x.c1.v[i] = x1[i]; b.y.c0.v[i] = y0[i]; b.y.c1.v[i] = y1[i]; }
  return b;
}

template <class F> __device__ __forceinline__ Affine<F> base_point();
template <> __device__ __forceinline__ Affine<Fq> base_point<Fq>() { return g1_base(); }
template <> __device__ __forceinline__ Affine<Fq2> base_point<Fq2>() { return g2_base(); }

template <class F>
__global__ void __launch_bounds__(128) k_fixed_base(Affine<F>* __restrict__ out, const Fr* __restrict__ scalars_mont, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr k = from_mont(scalars_mont[i]);
  out[i] = to_affine(scalar_mul(base_point<F>(), k.v));
}

template <class F>
__global__ void k_to_affine(Affine<F>* out, const XYZZ<F>* in, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = to_affine(in[i]);
}
// plain sum of affine points (fold of per-GPU partial results; `Sum for G1Local`, fr.rs:191-198)
template <class F>
__global__ void k_sum_affine(const Affine<F>* pts, size_t n, XYZZ<F>* out) {
  extern __shared__ uint4 smem_raw[];
  XYZZ<F>* sh = reinterpret_cast<XYZZ<F>*>(smem_raw);
  XYZZ<F> acc = XYZZ<F>::inf();
  for (size_t i = threadIdx.x; i < n; i += blockDim.x) acc = madd(acc, pts[i]);
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (uint32_t off = blockDim.x >> 1; off > 0; off >>= 1) {
    if (threadIdx.x < off) sh[threadIdx.x] = add(sh[threadIdx.x], sh[threadIdx.x + off]);
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = sh[0];
}

template <class F>
static int fixed_base_impl(zkb_ctx* ctx, Affine<F>* out, const Fr* scalars_mont, size_t n, cudaStream_t st) {
  if (!n) return ZKB_OK;
  ZKB_LAUNCH(ctx, k_fixed_base<F>, cdiv(n, 128), 128, 0, st, out, scalars_mont, n);
  return ZKB_OK;
}
template <class F>
static int to_affine_impl(zkb_ctx* ctx, Affine<F>* out, const XYZZ<F>* in, size_t n, cudaStream_t st) {
  if (!n) return ZKB_OK;
  ZKB_LAUNCH(ctx, k_to_affine<F>, cdiv(n, 64), 64, 0, st, out, in, n);
  return ZKB_OK;
}
template <class F>
static int sum_affine_impl(zkb_ctx* ctx, const Affine<F>* pts, size_t n, XYZZ<F>* d_out, cudaStream_t st) {
  ZKB_LAUNCH(ctx, k_sum_affine<F>, 1, 64, 64 * sizeof(XYZZ<F>), st, pts, n, d_out);
  return ZKB_OK;
}

}  // namespace zkb
