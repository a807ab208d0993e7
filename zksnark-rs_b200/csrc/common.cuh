// Internal definitions shared by the translation units of libzkb200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>
#include "../../include/zkb200.h"
#include "ec.cuh"

namespace zkb {

struct Ctx;

// device scratch arena entry
struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
};

}  // namespace zkb

// One proof in flight uses one lane: a latency-class stream (high priority: polynomial stage, record
// sort, head folding, bucket hierarchy, finish -- short or low-occupancy kernels) and a
// throughput-class stream (low priority: the bucket accumulations that fill every SM), tied together
// by events.  With two lanes, the latency-class work of proof i+1 and the tails of proof i run in
// the gaps of the accumulations instead of serially (zkb_prove_batch).
struct zkb_lane {
  cudaStream_t hi = nullptr, lo = nullptr;
  cudaStream_t hi2 = nullptr;    // second latency-class stream: the G1 tail runs beside the G2 tail (both low-occupancy)
  cudaEvent_t ev[5] = {};        // 0/1: G1 records sorted / accumulated; 2/3: same for G2; 4: G1 tail done
  zkb::DevBuf scratch[16];       // per-lane scratch (MSM buffers, polynomial workspace, outputs)
  void* h_proof = nullptr;       // pinned staging for the 256-byte result (+ the exchange status word of a sharded proof)
};

struct zkb_ctx {
  int device = 0;
  zkb_lane lanes[4];               // lane 0 always, the others created on first use (lane_get); zkb_prove_batch keeps `batch_lanes` proofs in flight
  int batch_lanes = 0;             // 0 = by problem size (prove.cu: batch_lane_count); ZKB_LANES=1..4 overrides (developer switch)
  cudaStream_t stream = nullptr;   // = lanes[0].hi: everything outside the prove pipeline runs here
  cudaStream_t stream2 = nullptr;  // = lanes[0].lo
  int sm_count = 148;
  uint64_t launches = 0;
  std::string err;
  // twiddle tables tw[log_n][inverse]: omega^k (k < n/2) in Montgomery form, built lazily
  zkb::Fr* tw[28][2] = {};
  // per-pass twiddle TILES twt[log_n][inverse][pass]: what one block of k_ntt_pass needs, contiguous (two 16-byte
  // planes), so the block fetches it with ONE bulk-async (TMA) copy into shared memory (ntt.cu); built with tw
  uint4* twt[28][2][4] = {};
  // reusable scratch (grown on demand, never shrunk) for the non-pipelined entry points
  zkb::DevBuf scratch[16];
  // optional per-kernel-class CUDA-event timing (zkb_profile): pairs recorded around tracked launches
  bool profile = false;
  bool trace = false;  // zkb_profile(ctx, 2): events around EVERY launch, dumped by zkb_trace_dump
  struct ProfRec { cudaEvent_t a, b; int kind; const char* name; int stream_id; };
  cudaEvent_t trace_base = nullptr;
  uint64_t prof_units[16] = {};  // work units per tracked class (NTT: elements x passes; ACC: records)
  std::vector<ProfRec> prof;
};

namespace zkb {

extern thread_local std::string g_err;

int set_err(zkb_ctx* ctx, int code, const char* fmt, ...);
// lane `idx` of the context, its streams / events / pinned staging created on first use (api.cu)
int lane_get(zkb_ctx* ctx, int idx, zkb_lane** out);

#define ZKB_CUDA(ctx, expr)                                                                     \
  do {                                                                                          \
    cudaError_t e__ = (expr);                                                                   \
    if (e__ != cudaSuccess)                                                                     \
      return ::zkb::set_err(ctx, ZKB_ERR_CUDA, "%s:%d: %s -> %s", __FILE__, __LINE__, #expr,    \
                            cudaGetErrorString(e__));                                           \
  } while (0)

#define ZKB_TRY(expr)          \
  do {                         \
    int rc__ = (expr);         \
    if (rc__ != ZKB_OK) return rc__; \
  } while (0)

// kernel launch bookkeeping: count + check
#define ZKB_LAUNCH(ctx, kernel, grid, block, smem, strm, ...)                                   \
  do {                                                                                          \
    if ((ctx)->trace) ::zkb::prof_begin(ctx, 0, strm, #kernel);                                 \
    kernel<<<(grid), (block), (smem), (strm)>>>(__VA_ARGS__);                                   \
    if ((ctx)->trace) ::zkb::prof_end(ctx, strm);                                               \
    (ctx)->launches++;                                                                          \
    cudaError_t e__ = cudaGetLastError();                                                       \
    if (e__ != cudaSuccess)                                                                     \
      return ::zkb::set_err(ctx, ZKB_ERR_CUDA, "%s:%d: launch %s -> %s", __FILE__, __LINE__,    \
                            #kernel, cudaGetErrorString(e__));                                  \
  } while (0)

// tracked kernel classes for zkb_profile
enum ProfKind { PK_NTT = 1, PK_ACC_G1 = 2, PK_ACC_G2 = 3, PK_SORT = 4, PK_REDUCE = 5, PK_POINTWISE = 6, PK_ASSEMBLE = 7, PK_MAX = 8 };
void prof_begin(zkb_ctx* ctx, int kind, cudaStream_t st, const char* name = "");
void prof_end(zkb_ctx* ctx, cudaStream_t st);

// tracked launch: CUDA events on the launching stream around the kernel when profiling is on
#define ZKB_LAUNCH_K(ctx, kind, kernel, grid, block, smem, strm, ...)                           \
  do {                                                                                          \
    if ((ctx)->profile) ::zkb::prof_begin(ctx, kind, strm, #kernel);                            \
    kernel<<<(grid), (block), (smem), (strm)>>>(__VA_ARGS__);                                   \
    if ((ctx)->profile) ::zkb::prof_end(ctx, strm);                                             \
    (ctx)->launches++;                                                                          \
    cudaError_t e__ = cudaGetLastError();                                                       \
    if (e__ != cudaSuccess)                                                                     \
      return ::zkb::set_err(ctx, ZKB_ERR_CUDA, "%s:%d: launch %s -> %s", __FILE__, __LINE__,    \
                            #kernel, cudaGetErrorString(e__));                                  \
  } while (0)

// grow-only scratch slot (ctx->scratch, or an explicit slot array such as a lane's)
int scratch_get(zkb_ctx* ctx, int slot, size_t bytes, void** out);
int scratch_get_in(zkb_ctx* ctx, DevBuf* slots, int slot, size_t bytes, void** out);
void prof_clear(zkb_ctx* ctx);
// host helpers (api.cu): canonical limbs / small integers -> Montgomery Fr; device Fq array -> canonical host limbs
Fr fr_from_limbs(const uint64_t* l);
Fr fr_from_u64(uint64_t x);
int download_fq(zkb_ctx* ctx, uint64_t* dst, const void* d_src, size_t n_fq);

static inline unsigned cdiv(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

// ---- ntt.cu -------------------------------------------------------------------------------------
// In-place transform, Montgomery form.  DIF: natural in -> bit-reversed out.
int ntt_dif(zkb_ctx* ctx, Fr* d, uint32_t log_n, bool inverse, cudaStream_t st);
// DIT: bit-reversed in -> natural out.
int ntt_dit(zkb_ctx* ctx, Fr* d, uint32_t log_n, bool inverse, cudaStream_t st);
// out[bitrev(i)] = in[i] * scale (scale may be null), out != in
int bitrev_permute(zkb_ctx* ctx, Fr* out, const Fr* in, uint32_t log_n, const Fr* h_scale, cudaStream_t st);
// d[i] *= base^i  (Montgomery form); if bitrev, position i holds logical index bitrev(i)
int scale_powers(zkb_ctx* ctx, Fr* d, uint32_t log_n, const Fr& base, const Fr& first, bool bitrev, cudaStream_t st);
int vec_to_mont(zkb_ctx* ctx, Fr* d, size_t n, bool to, cudaStream_t st);
int vec_mul(zkb_ctx* ctx, Fr* out, const Fr* a, const Fr* b, size_t n, cudaStream_t st);
// out[i] = first * base^i
int fill_powers(zkb_ctx* ctx, Fr* out, const Fr& base, const Fr& first, size_t n, cudaStream_t st);
int ntt_combine(zkb_ctx* ctx, const Fr* parts, uint32_t log_n, uint32_t log_g, bool inverse, uint64_t k0, uint64_t count, Fr* out,
                cudaStream_t st);
Fr host_omega(uint32_t log_n, bool inverse);  // Montgomery form
int get_twiddles(zkb_ctx* ctx, uint32_t log_n, bool inverse, Fr** out);

// ---- msm_*.cu ----------------------------------------------------------------------------------
// Fixed-base MSM over an expanded table T[j][i] = 2^(c*j) P_i (row stride `stride`, W = 254/c + 1
// rows).  Each job: d_out[j] (XYZZ, Montgomery, device) = sum_{i < n} scalars[i] * P_i with
// canonical scalars in device memory.  All jobs of one call share the sort / accumulate launches.
struct MsmJob {
  const Fr* scalars;
  size_t n;
};
int msm_pick_c(size_t n_points);
static inline int msm_windows(int c) { return 254 / c + 1; }
// An MSM call in three phases so that a pipeline can put them on different streams:
//   msm_sort (digits, scan, scatter) -> msm_accumulate (zero buckets + THE hot kernel) -> msm_tail
//   (head folding + bucket hierarchy -> d_out).  msm_prepare validates and carves the scratch
//   (3 consecutive slots of `slots`, starting at slot_base).
// Chunk t of the sorted record array starts at start(t): the first T1 chunks have S1 records, the
// rest S2 < S1.  Blocks are dispatched in index order, so the short chunks run last and the grid
// drains in a fraction of a long chunk's duration (with one chunk size the last wave idles about
// half a chunk time: ~9 % of the kernel at 5-6 waves), while most records still sit in long chunks,
// which keeps the number of head pieces (one per chunk boundary inside a bucket) small.
struct ChunkPlan {
  uint32_t S1 = 32, S2 = 32, T1 = 0;
  __host__ __device__ uint32_t R1() const { return T1 * S1; }
  __host__ __device__ uint32_t start(uint32_t t) const { return t < T1 ? t * S1 : R1() + (t - T1) * S2; }
  __host__ __device__ uint32_t len(uint32_t t) const { return t < T1 ? S1 : S2; }
  // smallest t with start(t) >= x
  __host__ __device__ uint32_t first_at_or_after(uint32_t x) const {
    return x <= R1() ? (x + S1 - 1) / S1 : T1 + (x - R1() + S2 - 1) / S2;
  }
  size_t count(size_t max_recs) const {
    size_t r1 = (size_t)T1 * S1;
    return max_recs <= r1 ? (max_recs + S1 - 1) / S1 : T1 + (max_recs - r1 + S2 - 1) / S2;
  }
};
static const int MSM_AFF_MAX_LEVELS = 8;
struct MsmPlan {
  int group = 1;  // 1: G1 (Fq), 2: G2 (Fq2)
  const void* tab = nullptr;
  size_t stride = 0;
  int c = 0, njobs = 0;
  ChunkPlan ch;  // how the sorted records are cut into per-thread chunks
  int win_rank = 0, win_world = 1;  // window sharding: only table rows j = win_rank (mod win_world) emit records
  int tail = 0;  // bucket hierarchy: 0 = quad plan (latency: the caller waits for this proof), 2 = quad plan with quads from 2048 elements
                 // (several small proofs in flight), 1 = v1 plan (least multiplier work: long accumulations of other proofs hide it)
  MsmJob jobs[4];
  bool empty = true;
  size_t nbk = 0, max_recs = 0, nacc = 0, lvl_elems = 0;
  uint32_t *hist = nullptr, *offs = nullptr, *cursor = nullptr, *sums = nullptr, *sorted = nullptr;
  void *buckets = nullptr, *heads = nullptr, *lvlS = nullptr, *lvlA = nullptr, *d_out = nullptr;
  // batched-affine pair tree in front of the XYZZ chain (affine_level.cuh; 0 levels = off): level l = 1 .. aff_levels has
  // the offsets aff_offs + (l - 1) * (nbk + 1) and at most aff_max[l] elements (aff_max[0] = max_recs), written to
  // aff_buf[(l - 1) & 1]; the chain then runs over the last level with the chunk plan ch_fin (nacc_fin chunks)
  int aff_levels = 0, aff_batch = 16;
  size_t aff_max[MSM_AFF_MAX_LEVELS + 1] = {};
  uint32_t *aff_offs = nullptr, *aff_sums = nullptr, *aff_counters = nullptr;
  void* aff_buf[2] = {nullptr, nullptr};
  void* aff_prefix = nullptr;
  unsigned aff_blocks = 0, aff_k = 1;  // grid of the longest level (columns of the prefix scratch / 128); groups of 32 items per warp
  ChunkPlan ch_fin;
  size_t nacc_fin = 0;
  const uint32_t* offs_fin() const { return aff_levels ? aff_offs + (size_t)(aff_levels - 1) * (nbk + 1) : offs; }
  const ChunkPlan& ch_tail() const { return aff_levels ? ch_fin : ch; }
};
int msm_prepare(zkb_ctx* ctx, DevBuf* slots, int slot_base, int group, const void* tab, size_t stride, int c, const MsmJob* jobs,
                int njobs, void* d_out, MsmPlan* plan);
int msm_sort(zkb_ctx* ctx, const MsmPlan& p, cudaStream_t st);
int msm_accumulate(zkb_ctx* ctx, const MsmPlan& p, cudaStream_t st);
int msm_tail(zkb_ctx* ctx, const MsmPlan& p, cudaStream_t st);
// all phases on one stream with ctx->scratch (standalone zkb_msm)
// win_rank / win_world: the partial sum over the table rows (windows) j = win_rank (mod win_world) only
int msm_g1(zkb_ctx* ctx, const G1Affine* tab, size_t stride, int c, const MsmJob* jobs, int njobs, G1XYZZ* d_out,
           int slot_base, cudaStream_t st, int win_rank = 0, int win_world = 1);
int msm_g2(zkb_ctx* ctx, const G2Affine* tab, size_t stride, int c, const MsmJob* jobs, int njobs, G2XYZZ* d_out,
           int slot_base, cudaStream_t st, int win_rank = 0, int win_world = 1);
// fill rows 1..W-1 of the table from row 0 (columns [0, n))
int expand_table_g1(zkb_ctx* ctx, G1Affine* tab, size_t stride, size_t n, int c, cudaStream_t st);
int expand_table_g2(zkb_ctx* ctx, G2Affine* tab, size_t stride, size_t n, int c, cudaStream_t st);
int fixed_base_g1(zkb_ctx* ctx, G1Affine* out, const Fr* scalars_mont, size_t n, cudaStream_t st);
int fixed_base_g2(zkb_ctx* ctx, G2Affine* out, const Fr* scalars_mont, size_t n, cudaStream_t st);
// canonical <-> Montgomery for arrays of Fq (point coordinates)
int fq_to_mont(zkb_ctx* ctx, Fq* d, size_t n, bool to, cudaStream_t st);
int xyzz_to_affine_g1(zkb_ctx* ctx, G1Affine* out, const G1XYZZ* in, size_t n, cudaStream_t st);
int xyzz_to_affine_g2(zkb_ctx* ctx, G2Affine* out, const G2XYZZ* in, size_t n, cudaStream_t st);
// points supplied over the ABI (Montgomery form, device): *d_bad |= 1 if one is off its curve, |= 2 if a G2 point is
// outside the order-r subgroup (checked only when `subgroup`); d_bad must be zeroed by the caller
int check_points_g1(zkb_ctx* ctx, const G1Affine* pts, size_t n, int* d_bad, cudaStream_t st);
int check_points_g2(zkb_ctx* ctx, const G2Affine* pts, size_t n, bool subgroup, int* d_bad, cudaStream_t st);
int sum_affine_g1(zkb_ctx* ctx, const G1Affine* pts, size_t n, G1XYZZ* d_out, cudaStream_t st);
int sum_affine_g2(zkb_ctx* ctx, const G2Affine* pts, size_t n, G2XYZZ* d_out, cudaStream_t st);

}  // namespace zkb

// ---- shard.cu: one proof over several GPUs, exchanges over NVLink peer memory -----------------------
// Every rank owns an exchange WINDOW in its HBM that all peers map (CUDA IPC between processes, the raw
// pointer inside one process).  Ranks only ever WRITE into peers' windows (kernel stores over NVLink) and
// READ their own; a transfer is: remote stores, __threadfence_system, then a remote store of the channel's
// epoch into the peer's flag word, which the consumer kernel polls.  Layout (identical on every rank):
//   [0, 2048)        flags[channel][step][src rank] (u32 epochs)     [2048]  status word (1 = a wait timed out)
//   [4096, 24576)    partial-sum slots[channel][src rank][64 u32]
//   [32768, ...)     per channel 6 * m_max Fr: R0 (3 vectors) | R1 (2) | R2 (1), each [vector][src rank][q]
static const int ZKB_COMM_MAX_WORLD = 16;
static const int ZKB_COMM_CHANNELS = 5;  // one per proof lane + one for stand-alone transforms
static const int ZKB_COMM_STEPS = 4;     // three all-to-alls of the polynomial stage + the partial sums
namespace zkb {
struct CommView {  // by-value kernel argument
  char* base[ZKB_COMM_MAX_WORLD];  // base[g]: rank g's window as mapped into this process (base[rank]: own)
  int rank, world, lg;
  unsigned long long timeout_ns;
};
struct ShardTables {  // per (comm, log_n): this rank's slices of the cross-rank twiddles and coset tables
  Fr *Tinv = nullptr, *Tfwd = nullptr, *cosS = nullptr, *QS = nullptr;
};
}  // namespace zkb
struct zkb_comm {
  zkb_ctx* ctx = nullptr;
  int rank = 0, world = 1, lg = 0;
  uint32_t max_log_n = 0;
  size_t m_max = 0, window_bytes = 0;
  char* window = nullptr;  // own
  bool connected = false;
  bool peer_is_ipc[ZKB_COMM_MAX_WORLD] = {};
  zkb::CommView view;
  uint32_t epoch[ZKB_COMM_CHANNELS] = {};
  zkb::ShardTables tabs[28];
};

namespace zkb {
// shard.cu internals used by prove.cu
static inline size_t comm_flag_off(int ch, int step, int src) { return (((size_t)ch * ZKB_COMM_STEPS + step) * ZKB_COMM_MAX_WORLD + src) * 4; }
static const size_t COMM_STATUS_OFF = 2048, COMM_SLOTS_OFF = 4096, COMM_DATA_OFF = 32768;
static inline size_t comm_slot_off(int ch, int src) { return COMM_SLOTS_OFF + ((size_t)ch * ZKB_COMM_MAX_WORLD + src) * 256; }
static inline size_t comm_region_off(const zkb_comm* c, int ch, int region /*0,1,2*/) {
  static const size_t first[3] = {0, 3, 5};
  return COMM_DATA_OFF + ((size_t)ch * 6 + first[region]) * c->m_max * sizeof(Fr);
}
int shard_check(zkb_ctx* ctx, const zkb_comm* c, uint32_t log_n);
// builds (once) the cross-rank tables and the local transforms' twiddles of a size-2^log_n sharded transform
int shard_prepare(zkb_ctx* ctx, zkb_comm* c, uint32_t log_n);
int matvec_launch(zkb_ctx* ctx, const zkb_qap* q, const Fr* wmont, size_t k0, size_t kstride, size_t count, Fr* A, Fr* B, Fr* AB,
                  cudaStream_t st);
int combine_partials_launch(zkb_ctx* ctx, const uint32_t* d_partials, int world, size_t count, uint32_t* d_out, cudaStream_t st);
// polynomial stage of a sharded proof on channel `ch`: ws = 7 * (n / world) Fr; on return (stream order) un, vn, hn
// (Montgomery, layout S: local index s = k1 * q + t <-> coefficient (rank * q + t) + (n / world) * k1) are ws + 3m, 4m, 6m
int shard_poly_stage(zkb_ctx* ctx, zkb_comm* c, int ch, uint32_t epoch, const zkb_qap* q, Fr* ws, const Fr* wmont, cudaStream_t st);
// partial sums: `partial` (64 u32, device) -> every rank's slot; then wait for all ranks' records and fold them into `out`
int shard_exchange_partials(zkb_ctx* ctx, zkb_comm* c, int ch, uint32_t epoch, const uint32_t* partial, uint32_t* out, int* d_status_copy,
                            cudaStream_t st);
}  // namespace zkb

struct zkb_bases {
  int group = 1;
  size_t n = 0;
  int c = 0;          // window bits of the expanded table (0: not expanded yet)
  void* d = nullptr;  // G1Affine* or G2Affine*, Montgomery form: row 0 = the points, rows 1..W-1 = 2^(c*j) multiples
};

static const uint64_t ZKB_GENERIC_MAX_N = 32768;  // the n x n Lagrange table is 32 GiB there (180 GB of HBM)
struct zkb_qap {
  uint64_t n = 0, m = 0, n_input = 0;
  uint32_t log_n = 0;
  // by-gate CSR (transposed from the by-wire rows at upload): for gate k, entries [gptr[k], gptr[k+1])
  uint32_t* d_gptr[3] = {};
  uint32_t* d_wire[3] = {};
  zkb::Fr* d_coeff[3] = {};  // Montgomery
  // by-wire CSR as uploaded (setup evaluates rows at x)
  uint32_t* d_rptr[3] = {};
  uint32_t* d_gate[3] = {};
  zkb::Fr* d_rcoeff[3] = {};  // Montgomery, in row order
  uint64_t nnz[3] = {};
  // host copy of the by-gate structure (offsets and wires, no coefficients): the witness planner levelises it (witness.cu)
  std::vector<uint32_t> h_gptr[3], h_wire[3];
  // coset tables in bit-reversed position order: P[i] = g^br(i) / n, Q[i] = g^-br(i) / (2n), g = omega_2n
  zkb::Fr* d_cosP = nullptr;
  zkb::Fr* d_cosQ = nullptr;
  // generic root domain (explicit pairwise-distinct roots, n <= ZKB_GENERIC_MAX_N; the reference's
  // ASTParser numbers its gates 1..=n, circuit/mod.rs:517): dense O(n^2) tables built at upload
  bool generic = false;
  std::vector<zkb::Fr> h_roots;  // host copy (Montgomery) for t(x) in setup
  zkb::Fr* d_roots = nullptr;   // n
  zkb::Fr* d_Lc = nullptr;      // n x n: Lc[k*n + i] = coefficient i of the Lagrange basis polynomial L_k
  zkb::Fr* d_dinv = nullptr;    // n: 1 / t'(r_k)
  zkb::Fr* d_tc = nullptr;      // n + 1 coefficients of t(x) = prod (x - r_k)
  zkb::Fr* d_ginv = nullptr;    // n - 1 coefficients of 1 / rev(t) mod x^(n-1)
  // (the polynomial workspace -- 8 vectors of n Fr and the witness, canonical + Montgomery -- lives in
  // the lane scratch, so proofs in flight on different lanes can share one QAP)
};

struct zkb_crs {
  uint64_t n = 0, n_sum_gamma = 0, n_sum_delta = 0;
  int rank = 0, world = 1;
  int layout = 0;  // 0: contiguous index ranges; 1: xi / xi_t in the sharded transform's output layout S (shard.cu)
  // shard ranges [lo, hi) into the logical vectors (layout 1: xi / xi_t are LOCAL ranges [0, count))
  uint64_t xi_lo = 0, xi_hi = 0, xit_lo = 0, xit_hi = 0, sd_lo = 0, sd_hi = 0;
  // G1 table, row 0 = [xi1 shard | alpha1 beta1 delta1 | xi_t shard | sum_delta shard], rows j>0 = 2^(c1*j) multiples
  zkb::G1Affine* g1 = nullptr;
  uint64_t g1_cnt = 0;  // row stride
  int c1 = 0;
  // G2 table, row 0 = [xi2 shard | beta2 delta2]
  zkb::G2Affine* g2 = nullptr;
  uint64_t g2_cnt = 0;
  int c2 = 0;
  zkb::G1Affine* sum_gamma = nullptr;  // not on the prove path; kept for download / verify
  zkb::G2Affine* gamma2 = nullptr;
  uint64_t nxi() const { return xi_hi - xi_lo; }
  uint64_t nxt() const { return xit_hi - xit_lo; }
  uint64_t nsd() const { return sd_hi - sd_lo; }
  // column offsets inside a table row
  uint64_t off_fixed() const { return nxi(); }
  uint64_t off_xit() const { return nxi() + 3; }
  uint64_t off_sd() const { return nxi() + 3 + nxt(); }
};
