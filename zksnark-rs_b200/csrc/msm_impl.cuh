// Pippenger bucket multi-scalar multiplication over BN254 G1 / G2 for sm_100a, fixed-base form.
//
// Replaces the reference's MSM pattern `coeffs.zip(points).map(exp_encrypted_g*).sum()`
// (/root/reference/src/groth16/mod.rs:255-272, 279-290), i.e. n independent MSB-first
// double-and-add scalar multiplications (fr.rs:114-119) folded sequentially (fr.rs:191-223).
// Group addition is exact, so any summation order gives the same affine result.
//
// The bases of every MSM on the prove() path are CRS vectors: fixed for the life of the CRS.  So
// each base vector is expanded ONCE (at setup / upload) into a table
//     T[j][i] = 2^(c*j) * P_i,   j < W = floor(254/c) + 1,   affine,
// and a scalar k = sum_j d_j 2^(c*j) (signed digits d_j in [-2^(c-1), 2^(c-1)]) contributes the
// W records (T[j][i], d_j).  All windows then share ONE set of 2^(c-1) buckets: there is no
// per-window reduction and no Horner combine over windows, and c can be larger (fewer records).
//
// A call may carry several jobs over the same table (different scalar vectors over prefixes of
// the table): their bucket sets are laid side by side and every stage runs once for all of them.
//
// Pipeline (one stream, no host round trips):
//   1. k_digits_count    signed-digit recoding, histogram of bucket sizes           (L2 atomics)
//   2. scan              exclusive prefix sum of the bucket sizes
//   3. k_digits_scatter  counting sort of the (table index, sign) records by bucket (L2 atomics)
//   4. k_accumulate_chunks  one thread per S sorted records: XYZZ mixed additions, flush at
//                        bucket boundaries                                          (THE hot kernel)
//   5. k_fix_heads       fold pieces of buckets that span chunks
//   6. k_bucket_level    sum_b (b+1) B_b by a hierarchy of running sums, 16 -> 1 per level
//
// The kernels are templates on the coordinate field; the launches go through MsmLaunch<F> so that
// the heavy Fq2 instantiations can live in separate translation units (parallel nvcc jobs).
#pragma once
#include <stdlib.h>
#include "common.cuh"
#include "affine_level.cuh"

namespace zkb {

struct DigitPlan {
  int c;        // window bits
  int W;        // number of windows = floor(254 / c) + 1
  uint32_t nb;  // buckets = 2^(c-1); bucket b holds digit magnitude b+1
  int wr, ww;   // window sharding: digit j emits a record only if j % ww == wr (ww = 1: all)
};

__host__ __device__ inline DigitPlan make_plan(int c) {
  DigitPlan p;
  p.c = c;
  p.W = 254 / c + 1;
  p.nb = 1u << (c - 1);
  p.wr = 0;
  p.ww = 1;
  return p;
}

#ifndef ZKB_ACC_MIN_BLOCKS
#define ZKB_ACC_MIN_BLOCKS 1
#endif
#ifndef ZKB_ACC_THREADS
#define ZKB_ACC_THREADS 128
#endif
static const int MSM_MAX_JOBS = 4;
// Chunk plan for max_recs records: long chunks for the first ~7/8, short ones (a quarter as long) for the rest, so that
// the grid drains in a fraction of a long chunk's duration.  Long-chunk size by record count (measured, r01).  A plan that
// sizes the long chunks to fill the machine a whole number of times (S1 = 7/8 records / (resident threads x rounds)) was
// measured in round 2 and is NOT better: 2^20 on one GPU G2 7.55 -> 8.68 ms, a rank of an 8-GPU proof G1 1.81 -> 1.92 ms
// (profiles/r02_chunk_plan.txt) -- blocks do not run in lock-step rounds (bucket flushes and head pieces make chunk
// times uneven and the scheduler backfills), so full rounds followed by the short chunks only add a partial round.
// ZKB_ACC_S1 / ZKB_ACC_FRAC: developer switches for sweeps.
template <class F>
static inline ChunkPlan chunk_plan(size_t max_recs, int sm_count) {
  (void)sm_count;
  const bool g1 = sizeof(F) == sizeof(Fq);
  uint32_t S1 = g1 ? (max_recs >= ((size_t)1 << 25) ? 128 : (max_recs >= ((size_t)1 << 23) ? 64 : 32)) : (max_recs >= ((size_t)1 << 23) ? 128 : 32);
  if (const char* e = getenv(g1 ? "ZKB_ACC_S1" : "ZKB_ACC_S1_G2")) {
    int v = atoi(e);
    if (v >= 4 && v <= 1024) S1 = (uint32_t)v;
  }
  ChunkPlan ch;
  ch.S1 = S1;
  ch.S2 = S1 >= 32 ? S1 / 4 : S1;
  double frac = 0.875;
  if (const char* e = getenv("ZKB_ACC_FRAC")) frac = atof(e);
  ch.T1 = (uint32_t)((double)max_recs * frac / S1);
  if (max_recs < ((size_t)1 << 20) || ch.S2 == ch.S1) ch.T1 = (uint32_t)((max_recs + S1 - 1) / S1);  // one chunk size
  return ch;
}

// window size for a table over n points (single bucket set): N*W mixed adds (10 modmul) + the
// bucket hierarchy (2 full adds of 14 modmul per bucket, weighted x2 for its lower parallelism)
static inline int pick_c(size_t n) {
  if (const char* e = getenv("ZKB_MSM_C")) {
    int c = atoi(e);
    if (c >= 2 && c <= 23) return c;
  }
  int best = 2;
  double best_cost = 1e300;
  for (int c = 2; c <= 22; c++) {
    double W = 254 / c + 1;
    double cost = W * 10.0 * (double)(n ? n : 1) + 56.0 * (double)((size_t)1 << (c - 1));
    if (cost < best_cost) { best_cost = cost; best = c; }
  }
  return best;
}

// launch indirection (definitions: msm_g1.cu for Fq; msm_g2_*.cu for Fq2)
template <class F>
struct MsmLaunch {
  static int accumulate(zkb_ctx* ctx, const Affine<F>* tab, const uint32_t* offs, const uint32_t* sorted, uint32_t nbk,
                        size_t nacc, ChunkPlan ch, XYZZ<F>* buckets, XYZZ<F>* heads, cudaStream_t st, int prof_kind);
  // the batched-affine pair tree (P.aff_levels levels) followed by the XYZZ chain over the last level
  static int accumulate_affine(zkb_ctx* ctx, const MsmPlan& P, cudaStream_t st, int prof_kind);
  static int fix_heads(zkb_ctx* ctx, const uint32_t* offs, uint32_t nbk, ChunkPlan ch, XYZZ<F>* buckets, const XYZZ<F>* heads,
                       cudaStream_t st);
  // full hierarchy: buckets[njobs][nb] -> d_out[njobs]; lvlS / lvlA hold the intermediate levels
  static int reduce(zkb_ctx* ctx, const XYZZ<F>* buckets, uint32_t nb, int njobs, XYZZ<F>* lvlS, XYZZ<F>* lvlA, XYZZ<F>* d_out,
                    cudaStream_t st, int tail);
  static int expand_table(zkb_ctx* ctx, Affine<F>* tab, size_t stride, size_t n, int c, cudaStream_t st);
  static int set_inf(zkb_ctx* ctx, XYZZ<F>* out, int n, cudaStream_t st);
};

// digit / scan stages (field independent; defined in msm_g1.cu)
int msm_sort_records(zkb_ctx* ctx, const MsmJob* jobs, int njobs, size_t stride, DigitPlan pl, uint32_t* hist, uint32_t* offs,
                     uint32_t* cursor, uint32_t* sums, uint32_t* sorted, cudaStream_t st);

static inline size_t msm_level_elems(uint32_t nb, int njobs) {
  // lvlS and lvlA are contiguous: 2 * (nb/2 + 128) elements per job.  The v1 plan (fan-in >= 4) needs nb/4 + nb/16 + ...
  // < nb/3 (+ one per level) in each; the quad plan (launch_reduce_quad) at most 2 nb/4 + 3 nb/16 + 4 nb/64 + ... < 0.78 nb
  // (+ a few per level) in total, checked there.
  return ((size_t)nb / 2 + 128) * njobs;
}

// Levels of the batched-affine pair tree in front of the XYZZ chain (affine_level.cuh); 0 = the chain alone.
// ZKB_AFF_G1 / ZKB_AFF_G2 = levels (0..8): developer switches.  Measured on B200 (profiles/r02_affine_tree_v*.jsonl,
// r02_notes.md): bit-exact, and NOT faster than the chain -- G2 at 2^20: 7.55 + 0.62 ms (chain + head folding) against
// 7.99 + 0.21 ms with five levels; G1 slower (3.4 against 2.4 ms) -- so the default is 0 for both groups.
static inline int affine_levels_policy(int group, size_t max_recs, size_t nbk) {
  int L = 0;
  if (const char* e = getenv(group == 1 ? "ZKB_AFF_G1" : "ZKB_AFF_G2")) L = atoi(e);
  if (L < 0) L = 0;
  if (L > MSM_AFF_MAX_LEVELS) L = MSM_AFF_MAX_LEVELS;
  if (max_recs >= ((size_t)1 << 31) - nbk) L = 0;  // source words keep bit 31 for the sign
  while (L > 0 && (max_recs >> L) < nbk / 4) L--;  // a level is worth a launch while buckets still hold several elements
  return L;
}
static const int MSM_AFF_BATCH = 16;  // additions per inversion (32 measured: within 2 %; the host test also runs 4 and 32)
// Work items are handed out 32 at a time (one per lane); a warp takes up to K such groups and leaves.  K = 1: a block lives
// ~100 us, the grid is many waves long (no quantisation loss in the last wave, small levels still fill the machine) and
// the latency-class kernels of other proofs in flight find room.  Measured: K = 4 leaves the G2 levels at 3.25 waves (the
// last one a quarter full) and the late levels below one wave.  ZKB_AFF_K: developer switch.
static inline unsigned affine_groups_per_warp() {
  if (const char* e = getenv("ZKB_AFF_K")) {
    int k = atoi(e);
    if (k >= 1 && k <= (1 << 20)) return (unsigned)k;
  }
  return 1;
}
static inline unsigned affine_level_grid(size_t max_out, int batch, unsigned K) {
  const size_t items = (max_out + batch - 1) / batch, groups = (items + 31) / 32;
  const size_t blocks = (groups + 4 * (size_t)K - 1) / (4 * (size_t)K);
  return blocks ? (unsigned)blocks : 1u;
}

// phases of one MSM call (see MsmPlan in common.cuh); F-typed views of the plan's buffers
template <class F>
static int msm_prepare_t(zkb_ctx* ctx, DevBuf* slots, int slot, const Affine<F>* tab, size_t stride, int c, const MsmJob* jobs,
                         int njobs, XYZZ<F>* d_out, MsmPlan* P) {
  if (njobs < 1 || njobs > MSM_MAX_JOBS) return set_err(ctx, ZKB_ERR_ARG, "msm: %d jobs", njobs);
  if (c < 2 || c > 23) return set_err(ctx, ZKB_ERR_ARG, "msm: window_bits %d out of range [2,23]", c);
  DigitPlan pl = make_plan(c);
  size_t total_n = 0;
  for (int j = 0; j < njobs; j++) {
    if (jobs[j].n > stride) return set_err(ctx, ZKB_ERR_ARG, "msm: more scalars than bases");
    total_n += jobs[j].n;
    P->jobs[j] = jobs[j];
  }
  if ((size_t)pl.W * stride >= ((size_t)1 << 31) || (size_t)pl.W * total_n >= ((size_t)1 << 32) - 64)
    return set_err(ctx, ZKB_ERR_ARG, "msm: too many records for 32-bit indices");
  P->group = sizeof(F) == sizeof(Fq) ? 1 : 2;
  P->tab = tab; P->stride = stride; P->c = c; P->njobs = njobs; P->d_out = d_out;
  P->empty = total_n == 0;
  if (P->empty) return ZKB_OK;
  P->nbk = (size_t)njobs * pl.nb;
  const size_t nscan_blocks = (P->nbk + 1023) / 1024;
  P->max_recs = (size_t)pl.W * total_n;
  P->ch = chunk_plan<F>(P->max_recs, ctx->sm_count);
  P->nacc = P->ch.count(P->max_recs);
  P->lvl_elems = msm_level_elems(pl.nb, njobs);
  // pair tree: sizes of the levels (sum of ceil(k_b / 2) <= (sum k_b + nbk) / 2), chunk plan of the chain behind them
  P->aff_levels = affine_levels_policy(P->group, P->max_recs, P->nbk);
  P->aff_batch = MSM_AFF_BATCH;
  size_t nheads = P->nacc, aff_bytes = 0;
  if (P->aff_levels) {
    P->aff_max[0] = P->max_recs;
    for (int l = 1; l <= P->aff_levels; l++) P->aff_max[l] = (P->aff_max[l - 1] + P->nbk) / 2;
    P->aff_k = affine_groups_per_warp();
    P->aff_blocks = affine_level_grid(P->aff_max[1], P->aff_batch, P->aff_k);  // level 1 is the longest
    // chain behind the tree: few elements are left, so short chunks (a chunk is a serial chain of additions: what counts
    // is that the grid still fills the machine a few times over, not the number of head pieces)
    const size_t fin = P->aff_max[P->aff_levels];
    P->ch_fin = chunk_plan<F>(fin, ctx->sm_count);
    if (fin < ((size_t)1 << 23)) {
      uint32_t S = 32;
      while (S > 4 && fin / S < (size_t)ctx->sm_count * 1024) S >>= 1;
      P->ch_fin.S1 = P->ch_fin.S2 = S;
      P->ch_fin.T1 = (uint32_t)((fin + S - 1) / S);
    }
    P->nacc_fin = P->ch_fin.count(fin);
    if (P->nacc_fin > nheads) nheads = P->nacc_fin;
    aff_bytes = (P->aff_max[1] + (P->aff_levels >= 2 ? P->aff_max[2] : 0)) * sizeof(Affine<F>) +
                (size_t)P->aff_blocks * 128 * P->aff_batch * sizeof(F);
  }
  void* p;
  // u32 scratch: hist[nbk] | offs[nbk+1] | cursor[nbk] | sums[nscan_blocks+1] | pair tree: offsets[L][nbk+1] | sums[L][nscan_blocks+1] | counters[L]
  size_t u32_words = P->nbk * 3 + 1 + nscan_blocks + 8;
  const size_t aff_words = (size_t)P->aff_levels * (P->nbk + 1 + nscan_blocks + 1) + MSM_AFF_MAX_LEVELS;
  ZKB_TRY(scratch_get_in(ctx, slots, slot + 0, (u32_words + aff_words) * 4, &p));
  P->hist = (uint32_t*)p;
  P->offs = P->hist + P->nbk;
  P->cursor = P->offs + P->nbk + 1;
  P->sums = P->cursor + P->nbk;
  P->aff_offs = P->hist + u32_words;
  P->aff_sums = P->aff_offs + (size_t)P->aff_levels * (P->nbk + 1);
  P->aff_counters = P->aff_sums + (size_t)P->aff_levels * (nscan_blocks + 1);
  ZKB_TRY(scratch_get_in(ctx, slots, slot + 1, P->max_recs * 4, &p));
  P->sorted = (uint32_t*)p;
  ZKB_TRY(scratch_get_in(ctx, slots, slot + 2, (P->nbk + nheads + 2 * P->lvl_elems) * sizeof(XYZZ<F>) + aff_bytes, &p));
  XYZZ<F>* buckets = (XYZZ<F>*)p;
  P->buckets = buckets;
  P->heads = buckets + P->nbk;
  P->lvlS = buckets + P->nbk + nheads;
  P->lvlA = buckets + P->nbk + nheads + P->lvl_elems;
  if (P->aff_levels) {
    Affine<F>* a = reinterpret_cast<Affine<F>*>(buckets + P->nbk + nheads + 2 * P->lvl_elems);
    P->aff_buf[0] = a;
    P->aff_buf[1] = a + P->aff_max[1];
    P->aff_prefix = a + P->aff_max[1] + (P->aff_levels >= 2 ? P->aff_max[2] : 0);
  }
  return ZKB_OK;
}

template <class F>
static int msm_accumulate_t(zkb_ctx* ctx, const MsmPlan& P, cudaStream_t st) {
  if (P.empty) return ZKB_OK;
  const int prof_kind = P.group == 1 ? PK_ACC_G1 : PK_ACC_G2;
  ZKB_CUDA(ctx, cudaMemsetAsync(P.buckets, 0, P.nbk * sizeof(XYZZ<F>), st));  // all-zero XYZZ = identity
  if (ctx->profile) ctx->prof_units[prof_kind] += P.max_recs;
  if (P.aff_levels) {
    ZKB_CUDA(ctx, cudaMemsetAsync(P.aff_counters, 0, MSM_AFF_MAX_LEVELS * 4, st));
    return MsmLaunch<F>::accumulate_affine(ctx, P, st, prof_kind);
  }
  return MsmLaunch<F>::accumulate(ctx, (const Affine<F>*)P.tab, P.offs, P.sorted, (uint32_t)P.nbk, P.nacc, P.ch, (XYZZ<F>*)P.buckets,
                                  (XYZZ<F>*)P.heads, st, prof_kind);
}

template <class F>
static int msm_tail_t(zkb_ctx* ctx, const MsmPlan& P, cudaStream_t st) {
  if (P.empty) return MsmLaunch<F>::set_inf(ctx, (XYZZ<F>*)P.d_out, P.njobs, st);
  ZKB_TRY(MsmLaunch<F>::fix_heads(ctx, P.offs_fin(), (uint32_t)P.nbk, P.ch_tail(), (XYZZ<F>*)P.buckets, (const XYZZ<F>*)P.heads, st));
  return MsmLaunch<F>::reduce(ctx, (const XYZZ<F>*)P.buckets, make_plan(P.c).nb, P.njobs, (XYZZ<F>*)P.lvlS, (XYZZ<F>*)P.lvlA,
                              (XYZZ<F>*)P.d_out, st, P.tail);
}

// ================================================================================================
// kernels
// ================================================================================================
#if defined(__CUDACC__)

// Bucket accumulation, load-balanced: the sorted record array is cut into chunks of S records and
// every thread sums exactly one chunk with XYZZ mixed additions, flushing at bucket boundaries.
// A bucket that begins inside the chunk is written to buckets[g]; the leading piece of a bucket
// that began in an earlier chunk goes to heads[t] and is folded in by k_fix_heads.  buckets[] is
// zero-filled (= identity) beforehand, empty buckets are never touched.
// Gather of one table entry.  A G1 entry is 64 bytes and 64-byte aligned, but the L2 fills 128-byte lines by default
// (ncu: 134.7 B of DRAM reads per record, profiles/r02_l2_fetch_granularity.txt; cudaLimitMaxL2FetchGranularity changed
// nothing): the loads carry the L2::64B prefetch-size qualifier instead (-DZKB_ACC_NO_L2_HINT builds the plain loads).
template <class F>
__device__ __forceinline__ Affine<F> ld_table_entry(const Affine<F>* p) {
#if !defined(ZKB_ACC_NO_L2_HINT)
  if (sizeof(Affine<F>) == 64) {
    Affine<F> r;
    uint32_t* w = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
    for (int k = 0; k < 4; k++)
      asm volatile("ld.global.L2::64B.v4.u32 {%0, %1, %2, %3}, [%4];"
                   : "=r"(w[4 * k]), "=r"(w[4 * k + 1]), "=r"(w[4 * k + 2]), "=r"(w[4 * k + 3])
                   : "l"(reinterpret_cast<const char*>(p) + 16 * k));
    return r;
  }
#endif
  return *p;
}

template <class F>
__global__ void __launch_bounds__(ZKB_ACC_THREADS, ZKB_ACC_MIN_BLOCKS) k_accumulate_chunks(const Affine<F>* __restrict__ pts, const uint32_t* __restrict__ offs,
                                                           const uint32_t* __restrict__ sorted, uint32_t nbk, size_t nchunks,
                                                           ChunkPlan ch, XYZZ<F>* __restrict__ buckets, XYZZ<F>* __restrict__ heads) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nchunks) return;
  const uint32_t total = offs[nbk];
  const uint32_t start = ch.start((uint32_t)t);
  if (start >= total) return;
  const uint32_t S = ch.len((uint32_t)t);
  const uint32_t end = (total - start > S) ? start + S : total;
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
#if defined(ZKB_ACC_PREFETCH) && ZKB_ACC_PREFETCH > 0
    // developer variant: pull the table entry of a later record towards the SM while this addition runs
    if (p + ZKB_ACC_PREFETCH < end)
      asm volatile("prefetch.global.L1 [%0];" ::"l"(pts + (sorted[p + ZKB_ACC_PREFETCH] & 0x7fffffffu)));
#endif
    uint32_t rec = sorted[p];
    Affine<F> P = ld_table_entry(pts + (rec & 0x7fffffffu));
    if (rec >> 31) P = neg(P);
    acc = madd(acc, P);
  }
  if (is_head) heads[t] = acc; else buckets[g] = acc;
}

// ---- the batched-affine pair tree in front of the chain (affine_level.cuh) -------------------------------------------
// One level: every warp takes 32 consecutive work items at a time from an atomic counter, one item (B output elements, one
// inversion) per lane, up to `groups` times.  The running products of a thread's item live in a global scratch laid out
// [element][thread] (coalesced 32-byte / 64-byte rows; the rows of the resident blocks stay in the L2).
#ifndef ZKB_AFF_MIN_BLOCKS
#define ZKB_AFF_MIN_BLOCKS 1
#endif
#ifndef ZKB_AFF_DEFAULT_VAR
#define ZKB_AFF_DEFAULT_VAR 1
#endif
// ---- staged variant of the level body (device only) -------------------------------------------------------------------
// ncu of the plain body (profiles/r02_affine_level_g2_v2.json): the multiplier is 55 % busy, a third of the warp time is
// long-scoreboard stalls -- operand loads and the local-memory source words -- which the two warps per scheduler the
// 220 registers of the Fq2 instantiation leave cannot cover.  Here the operands of the NEXT additions are copied global
// -> shared asynchronously (cp.async, no registers) while the current one is computed, and the source words live in
// shared memory: per thread 2 B words + max(4 stages x (x0, x1), 2 stages x (P0, P1, prefix)), laid out [word][thread]
// (conflict-free 16-byte accesses).  Every thread only reads what it copied itself: no block barrier anywhere.
namespace affdev {
__device__ __forceinline__ void cp16(uint32_t dst, const void* src, bool hint64) {
  if (hint64) asm volatile("cp.async.cg.shared.global.L2::64B [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
  else asm volatile("cp.async.cg.shared.global.L2::128B [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
template <class T> __device__ __forceinline__ T lds(const uint4* base, int word0) {  // sizeof(T) / 16 words, stride 128
  T r;
  uint4* d = reinterpret_cast<uint4*>(&r);
#pragma unroll
  for (int w = 0; w < (int)(sizeof(T) / 16); w++) d[w] = base[(word0 + w) * 128];
  return r;
}
}  // namespace affdev
template <class F, int B>
constexpr size_t affine_staged_smem() {
  // source words [2B][128] u32, then the stage words [NW][128] uint4
  return (size_t)2 * B * 128 * 4 + (size_t)2 * (2 * sizeof(Affine<F>) / 16 + sizeof(F) / 16) * 128 * 16;
}
template <class F, int B, bool FIRST>
__device__ __forceinline__ void affine_level_item_staged(uint32_t item, const Affine<F>* __restrict__ pts, const uint32_t* __restrict__ sorted,
                                                         const uint32_t* __restrict__ offs_in, const uint32_t* __restrict__ offs_out, uint32_t nbk,
                                                         Affine<F>* __restrict__ out, F* __restrict__ prefix, size_t pstride, uint32_t* sw, uint4* stg) {
  using namespace affdev;
  constexpr int WP = sizeof(Affine<F>) / 16, WF = sizeof(F) / 16;
  constexpr int SB = 2 * WP + WF;      // words of one backward stage: P0 | P1 | prefix
  constexpr int NSF = 4, SF = 2 * WF;  // forward: 4 stages of x0 | x1  (4 * 2 WF <= 2 * SB)
  constexpr bool H = FIRST && sizeof(Affine<F>) == 64;
  const uint32_t total = offs_out[nbk];
  const uint32_t start = item * (uint32_t)B;
  if (start >= total) return;
  const uint32_t cnt = total - start < (uint32_t)B ? total - start : (uint32_t)B;
  uint32_t lo = 0, hi = nbk;
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (offs_out[mid] <= start) lo = mid + 1; else hi = mid;
  }
  uint32_t g = lo - 1;
  uint32_t o_lo = offs_out[g], o_hi = offs_out[g + 1], i_lo = offs_in[g], i_hi = offs_in[g + 1];
  uint32_t pairmask = 0;
  // sw[j * 128] = s0[j], sw[(B + j) * 128] = s1[j]   (sw, stg already point at this thread's column)
  for (uint32_t j = 0; j < cnt; j++) {
    const uint32_t p = start + j;
    while (p >= o_hi) {
      g++;
      o_lo = o_hi; o_hi = offs_out[g + 1];
      i_lo = offs_in[g]; i_hi = offs_in[g + 1];
    }
    const uint32_t i0 = i_lo + 2 * (p - o_lo);
    const bool pair = i0 + 1 < i_hi;
    if (FIRST) {
      sw[j * 128] = sorted[i0];
      sw[(B + j) * 128] = pair ? sorted[i0 + 1] : 0;
    } else {
      sw[j * 128] = i0;
      sw[(B + j) * 128] = i0 + 1;
    }
    if (pair) pairmask |= 1u << j;
  }
  const uint32_t stg_addr = (uint32_t)__cvta_generic_to_shared(stg);
  // ---- forward: x0, x1 of element j in stage j % NSF, one commit group per element ---------------------------------
  auto issue_f = [&](uint32_t j) {
    if ((pairmask >> j) & 1u) {
      const char* a0 = reinterpret_cast<const char*>(pts + (sw[j * 128] & aff::IDX));
      const char* a1 = reinterpret_cast<const char*>(pts + (sw[(B + j) * 128] & aff::IDX));
      const uint32_t d = stg_addr + (uint32_t)((j % NSF) * SF) * 128 * 16;
#pragma unroll
      for (int w = 0; w < WF; w++) {
        cp16(d + w * 128 * 16, a0 + 16 * w, H);
        cp16(d + (WF + w) * 128 * 16, a1 + 16 * w, H);
      }
    }
    commit();
  };
  for (uint32_t j = 0; j < (uint32_t)(NSF - 1); j++) {
    if (j < cnt) issue_f(j); else commit();
  }
  F acc = F::one();
  for (uint32_t j = 0; j < cnt; j++) {
    if (j + NSF - 1 < cnt) issue_f(j + NSF - 1); else commit();
    wait<NSF - 1>();  // groups 0 .. j have landed
    if (!((pairmask >> j) & 1u)) continue;
    const int st = (int)(j % NSF) * SF;
    const F x0 = lds<F>(stg, st), x1 = lds<F>(stg, st + WF);
    F d = x1 - x0;
    if (d.is_zero() || x0.is_zero() || x1.is_zero()) {  // rare: identity operand, P + P, P - P
      const Affine<F> P0 = aff::load_point<F, H>(pts, sw[j * 128]), P1 = aff::load_point<F, H>(pts, sw[(B + j) * 128]);
      if (pair_kind(P0, P1, d) >= 2) continue;
    }
    prefix[j * pstride] = acc;
    acc = acc * d;
  }
  wait<0>();
  // ---- backward: P0 | P1 | prefix of element j in stage j & 1 ---------------------------------------------------------
  auto issue_b = [&](uint32_t j) {
    const uint32_t d = stg_addr + (uint32_t)((j & 1u) * SB) * 128 * 16;
    const char* a0 = reinterpret_cast<const char*>(pts + (sw[j * 128] & aff::IDX));
#pragma unroll
    for (int w = 0; w < WP; w++) cp16(d + w * 128 * 16, a0 + 16 * w, H);
    if ((pairmask >> j) & 1u) {
      const char* a1 = reinterpret_cast<const char*>(pts + (sw[(B + j) * 128] & aff::IDX));
      const char* pf = reinterpret_cast<const char*>(prefix + j * pstride);
#pragma unroll
      for (int w = 0; w < WP; w++) cp16(d + (WP + w) * 128 * 16, a1 + 16 * w, H);
#pragma unroll
      for (int w = 0; w < WF; w++) cp16(d + (2 * WP + w) * 128 * 16, pf + 16 * w, false);
    }
    commit();
  };
  issue_b(cnt - 1);  // lands during the inversion
  F inv = inverse(acc);
  for (uint32_t j = cnt; j-- > 0;) {
    if (j > 0) { issue_b(j - 1); wait<1>(); } else wait<0>();
    const int st = (int)(j & 1u) * SB;
    Affine<F> P0 = lds<Affine<F>>(stg, st);
    if (sw[j * 128] >> 31) P0 = neg(P0);
    if (!((pairmask >> j) & 1u)) {
      out[start + j] = P0;
      continue;
    }
    Affine<F> P1 = lds<Affine<F>>(stg, st + WP);
    if (sw[(B + j) * 128] >> 31) P1 = neg(P1);
    F d;
    const int kind = pair_kind(P0, P1, d);
    Affine<F> R;
    if (kind >= 2) {
      R = kind == 2 ? P1 : (kind == 3 ? P0 : Affine<F>::inf());
    } else {
      const F dinv = inv * lds<F>(stg, st + 2 * WP);
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

// VAR 0: plain body (affine_level.cuh, the one the CPU test runs); 1: staged body.  (A 168-register build of the plain Fq2
// body -- three blocks per SM instead of two, ~220 bytes of spills -- was 4 % faster than the 220-register one and is gone.)
template <class F, int B, bool FIRST, int VAR>
__global__ void __launch_bounds__(128, ZKB_AFF_MIN_BLOCKS)
    k_affine_level(const Affine<F>* __restrict__ pts, const uint32_t* __restrict__ sorted, const uint32_t* __restrict__ offs_in,
                   const uint32_t* __restrict__ offs_out, uint32_t nbk, Affine<F>* __restrict__ out, F* __restrict__ prefix_base,
                   uint32_t* __restrict__ counter, unsigned groups) {
  extern __shared__ uint4 aff_smem[];
  const uint32_t total = offs_out[nbk];
  const uint32_t n_items = total / B + (total % B ? 1u : 0u);
  const size_t pstride = (size_t)gridDim.x * blockDim.x;
  F* prefix = prefix_base + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t lane = threadIdx.x & 31;
  for (unsigned k = 0; k < groups; k++) {
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(counter, 32u);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (base >= n_items) break;
    if (base + lane < n_items) {
      if (VAR == 1)
        affine_level_item_staged<F, B, FIRST>(base + lane, pts, sorted, offs_in, offs_out, nbk, out, prefix, pstride,
                                              reinterpret_cast<uint32_t*>(aff_smem) + threadIdx.x, aff_smem + 2 * B * 128 / 4 + threadIdx.x);
      else
        affine_level_item<F, B, FIRST>(base + lane, pts, sorted, offs_in, offs_out, nbk, out, prefix, pstride);
    }
    __syncwarp();
  }
}
// The chain over the elements of the last level: k_accumulate_chunks with the points read in place.
template <class F>
__global__ void __launch_bounds__(ZKB_ACC_THREADS, ZKB_ACC_MIN_BLOCKS) k_accumulate_points(const Affine<F>* __restrict__ pts, const uint32_t* __restrict__ offs,
                                                                                           uint32_t nbk, size_t nchunks, ChunkPlan ch,
                                                                                           XYZZ<F>* __restrict__ buckets, XYZZ<F>* __restrict__ heads) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nchunks) return;
  const uint32_t total = offs[nbk];
  const uint32_t start = ch.start((uint32_t)t);
  if (start >= total) return;
  const uint32_t S = ch.len((uint32_t)t);
  const uint32_t end = (total - start > S) ? start + S : total;
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
    acc = madd(acc, pts[p]);
  }
  if (is_head) heads[t] = acc; else buckets[g] = acc;
}

// ---- variant: the accumulator in SHARED memory (G2) ----------------------------------------------------------------
// k_accumulate_chunks<Fq2> needs 230 registers: two blocks of 128 threads per SM, 8 warps, and the multiplier pipe idles
// ~20 % of the time on fixed-latency dependencies that two warps per scheduler cannot cover.  A third block needs
// <= 168 registers.  Here the 64-register XYZZ accumulator lives in shared memory (32 KB per block, 16-byte word k of
// thread t at [k][t]: conflict-free LDS.128 / STS.128) and the mixed addition loads each coordinate where it is used.
template <class F>
struct SmAcc {
  uint4* base;  // &smem[threadIdx.x]
  static constexpr int WORDS = sizeof(F) / 16;
  __device__ __forceinline__ F ld(int coord) const {
    F r;
    uint32_t* d = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
    for (int k = 0; k < WORDS; k++) {
      const uint32_t addr = (uint32_t)__cvta_generic_to_shared(base + (coord * WORDS + k) * ZKB_ACC_THREADS);
      asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(d[4 * k]), "=r"(d[4 * k + 1]), "=r"(d[4 * k + 2]), "=r"(d[4 * k + 3]) : "r"(addr));
    }
    return r;
  }
  __device__ __forceinline__ void st(int coord, const F& v) const {
    const uint32_t* d = reinterpret_cast<const uint32_t*>(&v);
#pragma unroll
    for (int k = 0; k < WORDS; k++) {
      const uint32_t addr = (uint32_t)__cvta_generic_to_shared(base + (coord * WORDS + k) * ZKB_ACC_THREADS);
      asm volatile("st.shared.v4.u32 [%4], {%0, %1, %2, %3};" ::"r"(d[4 * k]), "r"(d[4 * k + 1]), "r"(d[4 * k + 2]), "r"(d[4 * k + 3]), "r"(addr) : "memory");
    }
  }
  __device__ __forceinline__ XYZZ<F> get() const { XYZZ<F> r; r.x = ld(0); r.y = ld(1); r.zz = ld(2); r.zzz = ld(3); return r; }
  __device__ __forceinline__ void put(const XYZZ<F>& v) const { st(0, v.x); st(1, v.y); st(2, v.zz); st(3, v.zzz); }
};
// acc += p (madd-2008-s, complete); `inf` tracks acc == identity in a register
template <class F>
__device__ __forceinline__ void madd_sm(const SmAcc<F>& a, bool& inf, const Affine<F>& p) {
  if (p.is_inf()) return;
  if (inf) { a.st(0, p.x); a.st(1, p.y); a.st(2, F::one()); a.st(3, F::one()); inf = false; return; }
  const F u2 = p.x * a.ld(2);
  const F s2 = p.y * a.ld(3);
  const F pp_ = u2 - a.ld(0);
  const F rr = s2 - a.ld(1);
  if (pp_.is_zero()) {
    if (rr.is_zero()) a.put(dbl_affine(p)); else inf = true;
    return;
  }
  const F pp = sqr(pp_);
  const F ppp = pp_ * pp;
  const F q = a.ld(0) * pp;
  const F x3 = sqr(rr) - ppp - dbl(q);
  a.st(0, x3);
  a.st(1, mul_sub_mul(rr, q - x3, a.ld(1), ppp));
  a.st(2, a.ld(2) * pp);
  a.st(3, a.ld(3) * ppp);
}
template <class F>
__global__ void __launch_bounds__(ZKB_ACC_THREADS, 3) k_accumulate_chunks_sm(const Affine<F>* __restrict__ pts, const uint32_t* __restrict__ offs,
                                                                             const uint32_t* __restrict__ sorted, uint32_t nbk, size_t nchunks,
                                                                             ChunkPlan ch, XYZZ<F>* __restrict__ buckets, XYZZ<F>* __restrict__ heads) {
  __shared__ uint4 sm[sizeof(XYZZ<F>) / 16 * ZKB_ACC_THREADS];
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nchunks) return;
  const uint32_t total = offs[nbk];
  const uint32_t start = ch.start((uint32_t)t);
  if (start >= total) return;
  const uint32_t S = ch.len((uint32_t)t);
  const uint32_t end = (total - start > S) ? start + S : total;
  uint32_t lo = 0, hi = nbk;
  while (lo < hi) {
    uint32_t mid = (lo + hi) >> 1;
    if (offs[mid] <= start) lo = mid + 1; else hi = mid;
  }
  uint32_t g = lo - 1;
  bool is_head = offs[g] < start;
  uint32_t bend = offs[g + 1];
  SmAcc<F> acc;
  acc.base = sm + threadIdx.x;
  bool inf = true;
  for (uint32_t p = start; p < end; p++) {
    if (p == bend) {
      XYZZ<F>* dst = is_head ? heads + t : buckets + g;
      *dst = inf ? XYZZ<F>::inf() : acc.get();
      is_head = false;
      inf = true;
      do { g++; bend = offs[g + 1]; } while (bend <= p);
    }
    uint32_t rec = sorted[p];
    Affine<F> P = ld_table_entry(pts + (rec & 0x7fffffffu));
    if (rec >> 31) P = neg(P);
    madd_sm(acc, inf, P);
  }
  XYZZ<F>* dst = is_head ? heads + t : buckets + g;
  *dst = inf ? XYZZ<F>::inf() : acc.get();
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
// lies strictly inside bucket g are t with offs[g] < start(t) < offs[g+1] (computed from the offsets
// and the chunk plan, so no search).  Short runs (the common case) are summed by the owning thread; long runs (skewed
// scalars: one huge bucket) are summed by the whole warp, lanes striding over the run followed by
// a shuffle tree.
template <class F>
__global__ void __launch_bounds__(128) k_fix_heads(const uint32_t* __restrict__ offs, uint32_t nbk, ChunkPlan ch,
                                                   XYZZ<F>* __restrict__ buckets, const XYZZ<F>* __restrict__ heads) {
  const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  uint32_t t0 = 1, t1 = 0;
  if (g < nbk) {
    const uint32_t lo = offs[g], hi = offs[g + 1];
    if (hi > lo) { t0 = ch.first_at_or_after(lo + 1); t1 = ch.first_at_or_after(hi) - 1; }
  }
  const uint32_t cnt = t1 >= t0 ? t1 - t0 + 1 : 0;
  const bool big = cnt > 64;
  XYZZ<F> acc = XYZZ<F>::inf();
  if (!big)
    for (uint32_t t = t0; t <= t1; t++) acc = add_ool(acc, heads[t]);
  unsigned todo = __ballot_sync(0xffffffffu, big);
  while (todo) {
    const int src = __ffs(todo) - 1;
    todo &= todo - 1;
    const uint32_t b0 = __shfl_sync(0xffffffffu, t0, src), b1 = __shfl_sync(0xffffffffu, t1, src);
    XYZZ<F> part = XYZZ<F>::inf();
    for (uint32_t t = b0 + lane; t <= b1; t += 32) part = add_ool(part, heads[t]);
    for (int off = 16; off > 0; off >>= 1) {
      XYZZ<F> o = shfl_down_xyzz(part, off);
      part = add_ool(part, o);
    }
    XYZZ<F> tot = shfl_xyzz(part, 0);
    if (lane == src) acc = tot;
  }
  if (cnt) buckets[g] = add_ool(buckets[g], acc);
}

// Bucket reduction  R = sum_b (b+1) B_b = F(B) + G(B),  F(X) = sum_b b X_b,  G(X) = sum_b X_b.
// Cut X into chunks of L: with S_t = sum_i X_{tL+i} and T_t = sum_i i X_{tL+i},
//     F(X) = sum_t T_t + L * F(S),    G(X) = G(S).
// Level k (k = 0, 1, ...) therefore maps (S^k, A^k) -> (S^{k+1}, A^{k+1}) with
//     A^{k+1}_t = sum_i A^k_{tL+i} + L^k * T^{k+1}_t        (A^0 = 0),
// and when one element is left R = A + S.  grid.y = job (bucket sets are contiguous, n_in each).
template <class F>
__global__ void __launch_bounds__(128) k_bucket_level(const XYZZ<F>* __restrict__ S_in, const XYZZ<F>* __restrict__ A_in,
                                                      uint32_t n_in, uint32_t n_out, uint32_t L, int shift,
                                                      XYZZ<F>* __restrict__ S_out, XYZZ<F>* __restrict__ A_out, size_t in_stride,
                                                      size_t out_stride) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_out) return;
  const size_t job = blockIdx.y;
  const XYZZ<F>* s = S_in + job * in_stride + (size_t)t * L;
  const uint32_t cnt = n_in - t * L < L ? n_in - t * L : L;
  XYZZ<F> run = XYZZ<F>::inf(), acc = XYZZ<F>::inf();
  for (int i = (int)cnt - 1; i >= 1; i--) {
    run = add_ool(run, s[i]);
    acc = add_ool(acc, run);
  }
  run = add_ool(run, s[0]);
  for (int q = 0; q < shift; q++) acc = dbl_ool(acc);
  if (A_in) {
    const XYZZ<F>* a = A_in + job * in_stride + (size_t)t * L;
    for (uint32_t i = 0; i < cnt; i++) acc = add_ool(acc, a[i]);
  }
  S_out[job * out_stride + t] = run;
  A_out[job * out_stride + t] = acc;
}

// plain L:1 sums of the vectors 1 .. nvec-1 of a level (the partial sums of D_0 .. D_{nvec-2}, see below) beside a
// thread-per-chunk level of the S vector: grid.z = vector - 1
template <class F>
__global__ void __launch_bounds__(128) k_sum_level(const XYZZ<F>* __restrict__ in, uint32_t n_in, uint32_t n_out, uint32_t L,
                                                   size_t in_stride, XYZZ<F>* __restrict__ out, size_t out_stride) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_out) return;
  const size_t job = blockIdx.y, v = blockIdx.z + 1;
  const XYZZ<F>* s = in + job * in_stride + v * n_in + (size_t)t * L;
  const uint32_t cnt = n_in - t * L < L ? n_in - t * L : L;
  XYZZ<F> acc = s[0];
  for (uint32_t i = 1; i < cnt; i++) acc = add_ool(acc, s[i]);
  out[job * out_stride + v * n_out + t] = acc;
}

// ---- the hierarchy on quads (ec.cuh: qadd / qdbl) -- the default ----------------------------------------------------
// With b = sum_k d_k W_k (mixed radix, W_0 = 1, W_{k+1} = W_k L_k):   R = sum_b (b+1) B_b = G + sum_k W_k D_k,
// G = sum_b B_b,  D_k = sum_b d_k(b) B_b.  Level k cuts the running vector S^k (S^0 = B) into chunks of L_k = 8, one WARP
// per chunk, one QUAD per element: a 3-step suffix scan gives Suf_i (so S^{k+1}_t = Suf_0) and a 3-step tree over
// Suf_1..7 gives the chunk's share T_t = sum_i i X_i of D_k.  The T vectors are NOT folded into a weighted accumulator
// (that costs log2 W_k dependent doublings at every level: 45 of the 87 chained operations at 2^16 buckets); they ride
// along as plain vectors that later levels only tree-sum 8:1 (blockIdx.z = vector: 0 = S, v >= 1 = the partial sums of
// D_{v-1}), and ONE warp evaluates G + D_0 + W_1 (D_1 + ...) by Horner at the end: log2(nb) doublings in all.
// in: per job `nvec` vectors of n_in elements (job stride in_stride); out: per job nvec + 1 vectors of n_out elements
// (S', D_0' .. D_{nvec-2}', and the new T), job stride (nvec + 1) * n_out.
template <class F>
__global__ void __launch_bounds__(128) k_level_quad(const XYZZ<F>* __restrict__ in, uint32_t n_in, uint32_t n_out, int nvec,
                                                    size_t in_stride, XYZZ<F>* __restrict__ out) {
  const uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (t >= n_out) return;  // whole warps leave together
  const int lane = threadIdx.x & 31, quad = lane >> 2;
  const size_t job = blockIdx.y;
  const int v = blockIdx.z;
  const uint32_t e = t * 8 + quad;
  const XYZZ<F>* src = in + job * in_stride + (size_t)v * n_in;
  XYZZ<F>* dst = out + job * (size_t)(nvec + 1) * n_out;
  XYZZ<F> x = e < n_in ? src[e] : XYZZ<F>::inf();
  if (v == 0) {
    for (int off = 1; off < 8; off <<= 1) {
      XYZZ<F> y = shfl_down_xyzz(x, 4 * off);
      if (quad + off >= 8) y = XYZZ<F>::inf();
      x = qadd(x, y);
    }
    XYZZ<F> w = quad ? x : XYZZ<F>::inf();
    for (int off = 4; off > 0; off >>= 1) {
      XYZZ<F> y = shfl_down_xyzz(w, 4 * off);
      if (quad + off >= 8) y = XYZZ<F>::inf();
      w = qadd(w, y);
    }
    if (lane == 0) {
      dst[t] = x;
      dst[(size_t)nvec * n_out + t] = w;
    }
  } else {
    for (int off = 4; off > 0; off >>= 1) {
      XYZZ<F> y = shfl_down_xyzz(x, 4 * off);
      if (quad + off >= 8) y = XYZZ<F>::inf();
      x = qadd(x, y);
    }
    if (lane == 0) dst[(size_t)v * n_out + t] = x;
  }
}

struct HornerPlan {
  int nd;        // number of D vectors
  int sh[16];    // sh[k] = log2(W_{k+1} / W_k): doublings between D_{k+1} and D_k
};
// R = G + D_0 + W_1 (D_1 + (W_2 / W_1) (D_2 + ...)): one warp per job, every quad computes the same
template <class F>
__global__ void __launch_bounds__(32) k_horner_final(const XYZZ<F>* __restrict__ in, size_t in_stride, HornerPlan hp, XYZZ<F>* __restrict__ out) {
  const XYZZ<F>* v = in + blockIdx.x * in_stride;
  XYZZ<F> acc = XYZZ<F>::inf();
  for (int k = hp.nd - 1; k >= 0; k--) {
    acc = qadd(acc, v[1 + k]);
    if (k > 0)
      for (int q = 0; q < hp.sh[k - 1]; q++) acc = qdbl(acc);
  }
  acc = qadd(acc, v[0]);
  if (threadIdx.x == 0) out[blockIdx.x] = acc;
}

// The same level map with L = 32 and one WARP per chunk (lane i holds X_{32t+i}): a shuffle suffix
// scan gives Suf_i = sum_{k>=i} X_k (so S_t = Suf_0 and T_t = sum_{i>=1} Suf_i), then one shuffle
// tree sums 2^shift Suf_i + A_i over the lanes.  11 + shift dependent additions per level instead of
// ~3L: used once the level is too small to fill the SMs with one thread per chunk.
template <class F>
__global__ void __launch_bounds__(128) k_bucket_level_warp(const XYZZ<F>* __restrict__ S_in, const XYZZ<F>* __restrict__ A_in,
                                                           uint32_t n_in, uint32_t n_out, int shift, XYZZ<F>* __restrict__ S_out,
                                                           XYZZ<F>* __restrict__ A_out) {
  const uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (t >= n_out) return;
  const size_t job = blockIdx.y;
  const uint32_t e = t * 32 + lane;
  XYZZ<F> x = e < n_in ? S_in[job * n_in + e] : XYZZ<F>::inf();
  for (int off = 1; off < 32; off <<= 1) {
    XYZZ<F> y = shfl_down_xyzz(x, off);
    if (lane + off < 32) x = add_ool(x, y);
  }
  XYZZ<F> v = lane ? x : XYZZ<F>::inf();
  for (int q = 0; q < shift; q++) v = dbl_ool(v);
  if (A_in && e < n_in) v = add_ool(v, A_in[job * n_in + e]);
  for (int off = 16; off > 0; off >>= 1) {
    XYZZ<F> y = shfl_down_xyzz(v, off);
    v = add_ool(v, y);
  }
  if (lane == 0) {
    S_out[job * n_out + t] = x;
    A_out[job * n_out + t] = v;
  }
}

template <class F>
__global__ void k_bucket_final(const XYZZ<F>* __restrict__ S, const XYZZ<F>* __restrict__ A, int njobs, XYZZ<F>* __restrict__ out) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= njobs) return;
  out[j] = A ? add_ool(S[j], A[j]) : S[j];
}

template <class F>
__global__ void k_set_inf(XYZZ<F>* out, int n) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) out[j] = XYZZ<F>::inf();
}

// table expansion: T[j][i] = 2^(c*j) * T[0][i]
template <class F>
__global__ void __launch_bounds__(128) k_expand_table(Affine<F>* __restrict__ tab, size_t stride, size_t n, int c, int W) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  XYZZ<F> acc = to_xyzz(tab[i]);
  for (int j = 1; j < W; j++) {
    for (int q = 0; q < c; q++) acc = dbl(acc);
    tab[(size_t)j * stride + i] = to_affine(acc);
  }
}

// ---- launch bodies (instantiated by the .cu that owns the kernel) --------------------------------
template <class F>
static int launch_accumulate(zkb_ctx* ctx, const Affine<F>* tab, const uint32_t* offs, const uint32_t* sorted, uint32_t nbk,
                             size_t nacc, ChunkPlan ch, XYZZ<F>* buckets, XYZZ<F>* heads, cudaStream_t st, int prof_kind) {
#if defined(ZKB_ACC_SM_VARIANT)
  static const int sm_mode = getenv("ZKB_ACC_SM") ? atoi(getenv("ZKB_ACC_SM")) : ZKB_ACC_SM_VARIANT - 1;  // developer switch
  if (sm_mode) {
    ZKB_LAUNCH_K(ctx, prof_kind, k_accumulate_chunks_sm<F>, cdiv(nacc, ZKB_ACC_THREADS), ZKB_ACC_THREADS, 0, st, tab, offs, sorted, nbk, nacc, ch, buckets, heads);
    return ZKB_OK;
  }
#endif
  // ZKB_ACC_PAD_G1 / _G2 (bytes, <= 48 K): unused dynamic shared memory per block, i.e. fewer accumulation blocks per SM, so
  // that blocks of the latency-class kernels of other proofs in flight can be resident beside them (developer switch)
  static const size_t pad = [] { const char* e = getenv(sizeof(F) == sizeof(Fq) ? "ZKB_ACC_PAD_G1" : "ZKB_ACC_PAD_G2"); long v = e ? atol(e) : 0; return (size_t)(v > 0 && v <= 49152 ? v : 0); }();
  ZKB_LAUNCH_K(ctx, prof_kind, k_accumulate_chunks<F>, cdiv(nacc, ZKB_ACC_THREADS), ZKB_ACC_THREADS, pad, st, tab, offs, sorted, nbk, nacc, ch, buckets, heads);
  return ZKB_OK;
}
template <class F, int B, int VAR>
static int launch_affine_levels(zkb_ctx* ctx, const MsmPlan& P, cudaStream_t st, int prof_kind) {
  const uint32_t nbk = (uint32_t)P.nbk;
  const size_t smem = VAR == 1 ? affine_staged_smem<F, B>() : 0;
  if (VAR == 1) {
    static bool once = false;  // > 48 KB of dynamic shared memory needs the opt-in, once per kernel
    if (!once) {
      ZKB_CUDA(ctx, cudaFuncSetAttribute(k_affine_level<F, B, true, VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      ZKB_CUDA(ctx, cudaFuncSetAttribute(k_affine_level<F, B, false, VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      once = true;
    }
  }
  for (int l = 0; l < P.aff_levels; l++) {  // level l + 1 from level l
    const uint32_t* offs_in = l == 0 ? P.offs : P.aff_offs + (size_t)(l - 1) * (P.nbk + 1);
    const uint32_t* offs_out = P.aff_offs + (size_t)l * (P.nbk + 1);
    Affine<F>* out = (Affine<F>*)P.aff_buf[l & 1];
    unsigned blocks = affine_level_grid(P.aff_max[l + 1], B, P.aff_k);
    if (blocks > P.aff_blocks) blocks = P.aff_blocks;  // the prefix scratch has aff_blocks * 128 columns
    if (l == 0)
      ZKB_LAUNCH_K(ctx, prof_kind, (k_affine_level<F, B, true, VAR>), blocks, 128, smem, st, (const Affine<F>*)P.tab, P.sorted, offs_in, offs_out, nbk, out,
                   (F*)P.aff_prefix, P.aff_counters + l, P.aff_k);
    else
      ZKB_LAUNCH_K(ctx, prof_kind, (k_affine_level<F, B, false, VAR>), blocks, 128, smem, st, (const Affine<F>*)P.aff_buf[(l - 1) & 1], (const uint32_t*)nullptr,
                   offs_in, offs_out, nbk, out, (F*)P.aff_prefix, P.aff_counters + l, P.aff_k);
  }
  return ZKB_OK;
}
// ZKB_AFF_VAR: 0 plain level body, 1 staged (cp.async + shared memory; the default); developer switch
template <class F>
static int launch_accumulate_affine(zkb_ctx* ctx, const MsmPlan& P, cudaStream_t st, int prof_kind) {
  int var = ZKB_AFF_DEFAULT_VAR;
  if (const char* e = getenv("ZKB_AFF_VAR")) var = atoi(e) ? 1 : 0;
  if (var) ZKB_TRY((launch_affine_levels<F, MSM_AFF_BATCH, 1>(ctx, P, st, prof_kind)));
  else ZKB_TRY((launch_affine_levels<F, MSM_AFF_BATCH, 0>(ctx, P, st, prof_kind)));
  ZKB_LAUNCH_K(ctx, prof_kind, k_accumulate_points<F>, cdiv(P.nacc_fin, ZKB_ACC_THREADS), ZKB_ACC_THREADS, 0, st,
               (const Affine<F>*)P.aff_buf[(P.aff_levels - 1) & 1], P.offs_fin(), (uint32_t)P.nbk, P.nacc_fin, P.ch_fin, (XYZZ<F>*)P.buckets,
               (XYZZ<F>*)P.heads);
  return ZKB_OK;
}
template <class F>
static int launch_fix_heads(zkb_ctx* ctx, const uint32_t* offs, uint32_t nbk, ChunkPlan ch, XYZZ<F>* buckets, const XYZZ<F>* heads,
                            cudaStream_t st) {
  ZKB_LAUNCH(ctx, k_fix_heads<F>, cdiv(nbk, 128), 128, 0, st, offs, nbk, ch, buckets, heads);
  return ZKB_OK;
}
// level plan shared by the scratch sizing and the launches: first level one thread per chunk of L
// (16 when that still gives >= 32K threads, else 4), then warp levels (32 per warp)
static inline uint32_t msm_first_L(uint32_t nb, int njobs) { return (size_t)nb * njobs / 16 >= 32768 ? 16 : 4; }

// Quad plan.  While the vector is too long for the quad levels to run one warp per scheduler (they spend four lanes per
// addition: 4x the multiplier work of a thread), levels with one THREAD per chunk of L (16 from 2^19 elements over all
// jobs: throughput-bound there; else 4) -- their A output with shift 0 is exactly T, and k_sum_level folds the D vectors
// alongside; then quad levels 8:1 down to one element; then Horner.  Developer switches: ZKB_TAIL=1 selects the previous
// plan (launch_reduce_v1), ZKB_TAIL_QMAX the element count (all jobs) from which the quad levels take over.
template <class F>
static int launch_reduce_quad(zkb_ctx* ctx, const XYZZ<F>* buckets, uint32_t nb, int njobs, XYZZ<F>* lvl, size_t lvl_cap, XYZZ<F>* d_out,
                              cudaStream_t st, int tail) {
  // measured (profiles/r02_tail_plans.txt): quads from 8192 elements give the shortest chain (2^16 single proof 2.77 ms
  // vs 3.04 from 2048); with other proofs in flight 2048 costs less multiplier time (2^16 batch 2.06 vs 2.13 ms / proof)
  static const size_t qmax_env = getenv("ZKB_TAIL_QMAX") ? (size_t)atoll(getenv("ZKB_TAIL_QMAX")) : 0;
  const size_t qmax = qmax_env ? qmax_env : (tail == 2 ? 2048 : 8192);
  HornerPlan hp;
  hp.nd = 0;
  const XYZZ<F>* in = buckets;
  size_t in_stride = nb, used = 0;
  uint32_t m = nb;
  int nvec = 1;
  XYZZ<F>* cur = lvl;
  while (m > 1) {
    const bool quad = (size_t)m * njobs <= qmax;
    const uint32_t L = quad ? 8 : ((size_t)m * njobs >= ((size_t)1 << 19) ? 16 : 4);
    const uint32_t mo = (m + L - 1) / L;
    const size_t out_stride = (size_t)(nvec + 1) * mo, need = out_stride * njobs;
    if (used + need > lvl_cap || hp.nd >= 16) return set_err(ctx, ZKB_ERR_ALLOC, "msm reduce: level scratch too small");
    if (quad) {
      dim3 grid(cdiv((size_t)mo * 32, 128), njobs, nvec);
      ZKB_LAUNCH(ctx, k_level_quad<F>, grid, 128, 0, st, in, m, mo, nvec, in_stride, cur);
    } else {
      dim3 grid(cdiv(mo, 128), njobs);
      ZKB_LAUNCH(ctx, k_bucket_level<F>, grid, 128, 0, st, in, (const XYZZ<F>*)nullptr, m, mo, L, 0, cur, cur + (size_t)nvec * mo, in_stride,
                 out_stride);
      if (nvec > 1) {
        dim3 grid2(cdiv(mo, 128), njobs, nvec - 1);
        ZKB_LAUNCH(ctx, k_sum_level<F>, grid2, 128, 0, st, in, m, mo, L, in_stride, cur, out_stride);
      }
    }
    hp.sh[hp.nd++] = quad ? 3 : (L == 16 ? 4 : 2);
    in = cur; in_stride = out_stride; nvec++; m = mo;
    cur += need; used += need;
  }
  ZKB_LAUNCH(ctx, k_horner_final<F>, njobs, 32, 0, st, in, in_stride, hp, d_out);
  return ZKB_OK;
}

template <class F>
static int launch_reduce_v1(zkb_ctx* ctx, const XYZZ<F>* buckets, uint32_t nb, int njobs, XYZZ<F>* lvlS, XYZZ<F>* lvlA, XYZZ<F>* d_out,
                            cudaStream_t st) {
  const XYZZ<F>*Sin = buckets, *Ain = nullptr;
  XYZZ<F>*So = lvlS, *Ao = lvlA;
  int shift = 0;
  uint32_t m = nb;
  if (m > 1) {
    const uint32_t L = msm_first_L(nb, njobs);
    uint32_t mo = (m + L - 1) / L;
    dim3 grid(cdiv(mo, 128), njobs);
    ZKB_LAUNCH(ctx, k_bucket_level<F>, grid, 128, 0, st, Sin, Ain, m, mo, L, shift, So, Ao, (size_t)m, (size_t)mo);
    Sin = So; Ain = Ao;
    So += (size_t)mo * njobs; Ao += (size_t)mo * njobs;
    shift += L == 16 ? 4 : 2;
    m = mo;
  }
  while (m > 1) {
    uint32_t mo = (m + 31) / 32;
    dim3 grid(cdiv((size_t)mo * 32, 128), njobs);
    ZKB_LAUNCH(ctx, k_bucket_level_warp<F>, grid, 128, 0, st, Sin, Ain, m, mo, shift, So, Ao);
    Sin = So; Ain = Ao;
    So += (size_t)mo * njobs; Ao += (size_t)mo * njobs;
    shift += 5;
    m = mo;
  }
  ZKB_LAUNCH(ctx, k_bucket_final<F>, 1, 32, 0, st, Sin, Ain, njobs, d_out);
  return ZKB_OK;
}
template <class F>
static int launch_reduce(zkb_ctx* ctx, const XYZZ<F>* buckets, uint32_t nb, int njobs, XYZZ<F>* lvlS, XYZZ<F>* lvlA, XYZZ<F>* d_out,
                         cudaStream_t st, int tail) {
  static const int forced = getenv("ZKB_TAIL") ? atoi(getenv("ZKB_TAIL")) : -1;  // developer switch: 0 quad, 1 v1
  if ((forced >= 0 ? forced : tail) == 1) return launch_reduce_v1<F>(ctx, buckets, nb, njobs, lvlS, lvlA, d_out, st);
  // lvlS and lvlA are one contiguous region (msm_prepare_t)
  return launch_reduce_quad<F>(ctx, buckets, nb, njobs, lvlS, 2 * msm_level_elems(nb, njobs), d_out, st, tail);
}
template <class F>
static int launch_expand_table(zkb_ctx* ctx, Affine<F>* tab, size_t stride, size_t n, int c, cudaStream_t st) {
  if (!n) return ZKB_OK;
  DigitPlan pl = make_plan(c);
  if (pl.W > 1) ZKB_LAUNCH(ctx, k_expand_table<F>, cdiv(n, 128), 128, 0, st, tab, stride, n, c, pl.W);
  return ZKB_OK;
}
template <class F>
static int launch_set_inf(zkb_ctx* ctx, XYZZ<F>* out, int n, cudaStream_t st) {
  ZKB_LAUNCH(ctx, k_set_inf<F>, 1, 32, 0, st, out, n);
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

// ---- validation of points that arrive over the ABI -------------------------------------------------
// The reference can only hold group elements that crate `bn` constructed (fr.rs:102-123), so it never meets a point
// off the curve or -- on G2, whose curve has a cofactor != 1 -- outside the order-r subgroup.  Raw coordinates can:
// bit 0 of *bad: a point is neither the identity nor on y^2 = x^3 + b; bit 1: a G2 point with [r]P != O (the Miller
// loop / bucket sums are only meaningful on the r-torsion; EIP-197 mandates the same check).  G1 has cofactor 1.
template <class F> __device__ __forceinline__ F curve_b();
template <> __device__ __forceinline__ Fq curve_b<Fq>() {
  const uint32_t b[8] = ZKB_G1_B;
  Fq r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = b[i];
  return r;
}
template <> __device__ __forceinline__ Fq2 curve_b<Fq2>() {
  const uint32_t b0[8] = ZKB_G2_B0, b1[8] = ZKB_G2_B1;
  Fq2 r;
#pragma unroll
  for (int i = 0; i < 8; i++) { r.c0.v[i] = b0[i]; r.c1.v[i] = b1[i]; }
  return r;
}
template <class F>
__device__ __forceinline__ bool on_curve(const Affine<F>& p) {
  return p.is_inf() || sqr(p.y) == sqr(p.x) * p.x + curve_b<F>();
}
template <class F>
__device__ bool in_subgroup(const Affine<F>& p) {  // [r]P == O by double-and-add with the complete formulas
  const Fr m = Fr::modulus();
  return scalar_mul(p, m.v).is_inf();
}
template <class F>
__global__ void __launch_bounds__(128) k_check_points(const Affine<F>* __restrict__ pts, size_t n, int subgroup, int* __restrict__ bad) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Affine<F> p = pts[i];
  int f = 0;
  if (!on_curve(p)) f = 1;
  else if (subgroup && !p.is_inf() && !in_subgroup(p)) f = 2;
  if (f) atomicOr(bad, f);
}
template <class F>
static int check_points_impl(zkb_ctx* ctx, const Affine<F>* pts, size_t n, bool subgroup, int* d_bad, cudaStream_t st) {
  if (!n) return ZKB_OK;
  ZKB_LAUNCH(ctx, k_check_points<F>, cdiv(n, 128), 128, 0, st, pts, n, subgroup ? 1 : 0, d_bad);
  return ZKB_OK;
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
#endif  // __CUDACC__

}  // namespace zkb
