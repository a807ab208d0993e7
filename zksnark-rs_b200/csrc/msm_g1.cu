// G1 (Fq) instantiation of the MSM / point kernels (see msm_impl.cuh) and the field-independent
// record sort (signed-digit recoding + counting sort by bucket).
// Measured on B200 (profiles/r01_notes.md): for the G1 mixed addition the plain CIOS product with five
// 128-thread blocks per SM (96 registers) beats the wide-product / lazy-reduction code (fewer multiplies
// but a longer dependency chain): 0.968 vs 0.957 of the modmul peak.
#define ZKB_NO_LAZY 1
#define ZKB_ACC_MIN_BLOCKS 5
#define ZKB_AFF_MIN_BLOCKS 5  // k_affine_level<Fq>: 92 registers without spills
#include "msm_impl.cuh"
namespace zkb {

__device__ __forceinline__ uint32_t get_bits(const uint32_t k[8], int pos, int c) {
  // bits [pos, pos+c) of the 256-bit integer k (c <= 24)
  int w = pos >> 5, o = pos & 31;
  if (w >= 8) return 0;
  uint64_t lo = k[w];
  uint64_t hi = (w + 1 < 8) ? k[w + 1] : 0;
  uint64_t v = (lo | (hi << 32)) >> o;
  return (uint32_t)v & ((1u << c) - 1);
}

// scalars are canonical residues (< r < 2^254).  hist / cursor point at this job's bucket set.
__global__ void k_digits_count(const Fr* __restrict__ scalars, size_t n, DigitPlan pl, uint32_t* __restrict__ hist) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr k = scalars[i];
  uint32_t carry = 0;
  for (int j = 0; j < pl.W; j++) {
    uint32_t d = get_bits(k.v, j * pl.c, pl.c) + carry;
    carry = d > pl.nb;
    uint32_t mag = carry ? ((1u << pl.c) - d) : d;
    if (mag && (pl.ww == 1 || j % pl.ww == pl.wr)) atomicAdd(&hist[mag - 1], 1u);
  }
}

__global__ void k_digits_scatter(const Fr* __restrict__ scalars, size_t n, size_t stride, DigitPlan pl,
                                 uint32_t* __restrict__ cursor, uint32_t* __restrict__ sorted) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr k = scalars[i];
  uint32_t carry = 0;
  for (int j = 0; j < pl.W; j++) {
    uint32_t d = get_bits(k.v, j * pl.c, pl.c) + carry;
    carry = d > pl.nb;
    uint32_t mag = carry ? ((1u << pl.c) - d) : d;
    if (mag && (pl.ww == 1 || j % pl.ww == pl.wr)) {
      uint32_t pos = atomicAdd(&cursor[mag - 1], 1u);
      sorted[pos] = (uint32_t)((size_t)j * stride + i) | (carry << 31);
    }
  }
}

// exclusive scan of uint32 (three small kernels; total <= 2^24 entries)
static const int SCAN_B = 1024;  // elements per block (256 threads x 4)

__global__ void k_scan_block(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, uint32_t* __restrict__ sums, size_t n) {
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

__global__ void __launch_bounds__(256) k_scan_sums(uint32_t* sums, size_t nblocks, uint32_t* total) {
  // single block of 256 threads (it has to find room on an SM that bucket accumulations of another proof fill: a
  // 1024-thread block waited ~0.5 ms for that), four entries per thread, sequential over chunks of 1024
  __shared__ uint32_t sh[256];
  __shared__ uint32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (size_t base = 0; base < nblocks; base += 1024) {
    const size_t i0 = base + threadIdx.x * 4;
    uint32_t v[4], tot = 0;
#pragma unroll
    for (int q = 0; q < 4; q++) {
      v[q] = i0 + q < nblocks ? sums[i0 + q] : 0;
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
    uint32_t excl = carry + sh[threadIdx.x] - tot;
#pragma unroll
    for (int q = 0; q < 4; q++) {
      if (i0 + q < nblocks) sums[i0 + q] = excl;
      excl += v[q];
    }
    __syncthreads();
    if (threadIdx.x == 255) carry += sh[255];
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}

__global__ void k_scan_add(uint32_t* __restrict__ out, uint32_t* __restrict__ out2, const uint32_t* __restrict__ sums, size_t n,
                           const uint32_t* total) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    uint32_t v = out[i] + sums[i / SCAN_B];
    out[i] = v;
    out2[i] = v;
  }
  if (i == n) out[n] = *total;  // offsets[n] = total
}

int msm_sort_records(zkb_ctx* ctx, const MsmJob* jobs, int njobs, size_t stride, DigitPlan pl, uint32_t* hist, uint32_t* offs,
                     uint32_t* cursor, uint32_t* sums, uint32_t* sorted, cudaStream_t st) {
  const size_t nbk = (size_t)njobs * pl.nb;
  const size_t nscan_blocks = (nbk + SCAN_B - 1) / SCAN_B;
  ZKB_CUDA(ctx, cudaMemsetAsync(hist, 0, nbk * 4, st));
  for (int j = 0; j < njobs; j++)
    if (jobs[j].n)
      ZKB_LAUNCH(ctx, k_digits_count, cdiv(jobs[j].n, 256), 256, 0, st, jobs[j].scalars, jobs[j].n, pl, hist + (size_t)j * pl.nb);
  ZKB_LAUNCH(ctx, k_scan_block, (unsigned)nscan_blocks, 256, 0, st, hist, offs, sums, nbk);
  ZKB_LAUNCH(ctx, k_scan_sums, 1, 256, 0, st, sums, nscan_blocks, sums + nscan_blocks);
  ZKB_LAUNCH(ctx, k_scan_add, cdiv(nbk + 1, 256), 256, 0, st, offs, cursor, sums, nbk, sums + nscan_blocks);
  for (int j = 0; j < njobs; j++)
    if (jobs[j].n)
      ZKB_LAUNCH(ctx, k_digits_scatter, cdiv(jobs[j].n, 256), 256, 0, st, jobs[j].scalars, jobs[j].n, stride, pl,
                 cursor + (size_t)j * pl.nb, sorted);
  return ZKB_OK;
}

// ---- offsets of the pair-tree levels (affine_level.cuh): level l >= 1 holds ceil(k_b / 2^l) elements of bucket b, so all
// levels scan the level-0 offsets independently: the three scan kernels once, blockIdx.y = level - 1 -------------------
__global__ void k_lvl_scan_block(const uint32_t* __restrict__ offs0, uint32_t* __restrict__ lvl_offs, uint32_t* __restrict__ lvl_sums,
                                 size_t n, size_t nblocks) {
  __shared__ uint32_t sh[256];
  const int l = blockIdx.y + 1;
  uint32_t* out = lvl_offs + (size_t)blockIdx.y * (n + 1);
  uint32_t* sums = lvl_sums + (size_t)blockIdx.y * (nblocks + 1);
  size_t base = (size_t)blockIdx.x * SCAN_B + threadIdx.x * 4;
  uint32_t v[4], tot = 0;
#pragma unroll
  for (int q = 0; q < 4; q++) {
    v[q] = base + q < n ? affine_level_count(offs0[base + q + 1] - offs0[base + q], l) : 0;
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
__global__ void __launch_bounds__(256) k_lvl_scan_sums(uint32_t* lvl_sums, size_t nblocks) {
  __shared__ uint32_t sh[256];
  __shared__ uint32_t carry;
  uint32_t* sums = lvl_sums + (size_t)blockIdx.x * (nblocks + 1);
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (size_t base = 0; base < nblocks; base += 1024) {
    const size_t i0 = base + threadIdx.x * 4;
    uint32_t v[4], tot = 0;
#pragma unroll
    for (int q = 0; q < 4; q++) {
      v[q] = i0 + q < nblocks ? sums[i0 + q] : 0;
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
    uint32_t excl = carry + sh[threadIdx.x] - tot;
#pragma unroll
    for (int q = 0; q < 4; q++) {
      if (i0 + q < nblocks) sums[i0 + q] = excl;
      excl += v[q];
    }
    __syncthreads();
    if (threadIdx.x == 255) carry += sh[255];
    __syncthreads();
  }
  if (threadIdx.x == 0) sums[nblocks] = carry;  // the level's element count
}
__global__ void k_lvl_scan_add(uint32_t* __restrict__ lvl_offs, const uint32_t* __restrict__ lvl_sums, size_t n, size_t nblocks) {
  uint32_t* out = lvl_offs + (size_t)blockIdx.y * (n + 1);
  const uint32_t* sums = lvl_sums + (size_t)blockIdx.y * (nblocks + 1);
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] += sums[i / SCAN_B];
  if (i == n) out[n] = sums[nblocks];
}
static int msm_level_offsets(zkb_ctx* ctx, const MsmPlan& P, cudaStream_t st) {
  if (!P.aff_levels) return ZKB_OK;
  const size_t nscan_blocks = (P.nbk + SCAN_B - 1) / SCAN_B;
  ZKB_LAUNCH(ctx, k_lvl_scan_block, dim3((unsigned)nscan_blocks, P.aff_levels), 256, 0, st, P.offs, P.aff_offs, P.aff_sums, P.nbk, nscan_blocks);
  ZKB_LAUNCH(ctx, k_lvl_scan_sums, P.aff_levels, 256, 0, st, P.aff_sums, nscan_blocks);
  ZKB_LAUNCH(ctx, k_lvl_scan_add, dim3(cdiv(P.nbk + 1, 256), P.aff_levels), 256, 0, st, P.aff_offs, P.aff_sums, P.nbk, nscan_blocks);
  return ZKB_OK;
}

int msm_pick_c(size_t n) { return pick_c(n); }

template <> int MsmLaunch<Fq>::accumulate(zkb_ctx* ctx, const G1Affine* tab, const uint32_t* offs, const uint32_t* sorted,
                                          uint32_t nbk, size_t nacc, ChunkPlan ch, G1XYZZ* buckets, G1XYZZ* heads, cudaStream_t st,
                                          int pk) {
  return launch_accumulate<Fq>(ctx, tab, offs, sorted, nbk, nacc, ch, buckets, heads, st, pk);
}
template <> int MsmLaunch<Fq>::accumulate_affine(zkb_ctx* ctx, const MsmPlan& P, cudaStream_t st, int pk) {
  return launch_accumulate_affine<Fq>(ctx, P, st, pk);
}
template <> int MsmLaunch<Fq>::fix_heads(zkb_ctx* ctx, const uint32_t* offs, uint32_t nbk, ChunkPlan ch, G1XYZZ* buckets,
                                         const G1XYZZ* heads, cudaStream_t st) {
  return launch_fix_heads<Fq>(ctx, offs, nbk, ch, buckets, heads, st);
}
template <> int MsmLaunch<Fq>::reduce(zkb_ctx* ctx, const G1XYZZ* buckets, uint32_t nb, int njobs, G1XYZZ* lvlS, G1XYZZ* lvlA,
                                      G1XYZZ* d_out, cudaStream_t st, int tail) {
  return launch_reduce<Fq>(ctx, buckets, nb, njobs, lvlS, lvlA, d_out, st, tail);
}
template <> int MsmLaunch<Fq>::expand_table(zkb_ctx* ctx, G1Affine* tab, size_t stride, size_t n, int c, cudaStream_t st) {
  return launch_expand_table<Fq>(ctx, tab, stride, n, c, st);
}
template <> int MsmLaunch<Fq>::set_inf(zkb_ctx* ctx, G1XYZZ* out, int n, cudaStream_t st) { return launch_set_inf<Fq>(ctx, out, n, st); }

int msm_sort(zkb_ctx* ctx, const MsmPlan& P, cudaStream_t st) {
  if (P.empty) return ZKB_OK;
  DigitPlan pl = make_plan(P.c);
  pl.wr = P.win_rank;
  pl.ww = P.win_world;
  ZKB_TRY(msm_sort_records(ctx, P.jobs, P.njobs, P.stride, pl, P.hist, P.offs, P.cursor, P.sums, P.sorted, st));
  return msm_level_offsets(ctx, P, st);
}
int msm_g1(zkb_ctx* ctx, const G1Affine* tab, size_t stride, int c, const MsmJob* jobs, int njobs, G1XYZZ* d_out, int slot,
           cudaStream_t st, int win_rank, int win_world) {
  MsmPlan P;
  ZKB_TRY(msm_prepare(ctx, ctx->scratch, slot, 1, tab, stride, c, jobs, njobs, d_out, &P));
  P.win_rank = win_rank; P.win_world = win_world;
  ZKB_TRY(msm_sort(ctx, P, st));
  ZKB_TRY(msm_accumulate(ctx, P, st));
  return msm_tail(ctx, P, st);
}
int expand_table_g1(zkb_ctx* ctx, G1Affine* tab, size_t stride, size_t n, int c, cudaStream_t st) {
  return MsmLaunch<Fq>::expand_table(ctx, tab, stride, n, c, st);
}
int fixed_base_g1(zkb_ctx* ctx, G1Affine* out, const Fr* scalars_mont, size_t n, cudaStream_t st) {
  return fixed_base_impl<Fq>(ctx, out, scalars_mont, n, st);
}
int xyzz_to_affine_g1(zkb_ctx* ctx, G1Affine* out, const G1XYZZ* in, size_t n, cudaStream_t st) {
  return to_affine_impl<Fq>(ctx, out, in, n, st);
}
int check_points_g1(zkb_ctx* ctx, const G1Affine* pts, size_t n, int* d_bad, cudaStream_t st) {
  return check_points_impl<Fq>(ctx, pts, n, false, d_bad, st);
}
int sum_affine_g1(zkb_ctx* ctx, const G1Affine* pts, size_t n, G1XYZZ* d_out, cudaStream_t st) {
  return sum_affine_impl<Fq>(ctx, pts, n, d_out, st);
}
__global__ void k_fq_to_mont(Fq* d, size_t n, int to) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  d[i] = to ? to_mont(d[i]) : from_mont(d[i]);
}
int fq_to_mont(zkb_ctx* ctx, Fq* d, size_t n, bool to, cudaStream_t st) {
  if (!n) return ZKB_OK;
  ZKB_LAUNCH(ctx, k_fq_to_mont, cdiv(n, 256), 256, 0, st, d, n, to ? 1 : 0);
  return ZKB_OK;
}
}  // namespace zkb
