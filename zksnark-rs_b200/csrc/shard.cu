// One proof over G = 2^lg GPUs: the polynomial stage with the NTT's OUTER DIMENSION sharded across the ranks and the
// exchanges done by the kernels themselves over NVLink peer memory (include/zkb200.h: zkb_comm_*, zkb_ntt_shard; the
// proof entry points are in prove.cu).
//
// Replaces, like ntt.cu, the polynomial arithmetic behind h = (u_sum * v_sum - w_sum) / t
// (/root/reference/src/groth16/mod.rs:233-253, 277; coefficient_poly.rs:93-157; field/mod.rs:428-469) -- the reference
// is single-threaded and has no device or process boundary; this split is new.
//
// n = G*m, m = G*q.  Two ownership layouts of a length-n vector:
//   D (decimated)       rank r owns x[r + G*j2], j2 < m                        local index j2
//   S (strided blocks)  rank r owns x[(r*q + t) + m*k1], t < q, k1 < G         local index s = k1*q + t
// and two distributed transforms X[k] = sum_j x[j] w^(jk), each with ONE all-to-all of n*32*(G-1)/G bytes in total
// and log G butterfly stages in registers (not an O(G) sum per output):
//   DIT-distributed  D -> S:  local size-m transform Y_r; rank d receives Y_g[d*q .. (d+1)*q) from every g; then
//                             X[k2 + m*k1] = sum_g (w^m)^(g k1) * (w^(g k2) Y_g[k2])      (twiddle + size-G transform)
//   DIF-distributed  S -> D:  Z[j2][k1] = w^(j2 k1) * sum_j1 (w^m)^(j1 k1) x[j2 + m*j1]   (size-G transform + twiddle);
//                             rank k1 receives Z[.][k1] = W_k1; then X[k1 + G*k2] = local size-m transform of W_k1
// The proof chains them: gate evaluations A, B, A.B (layout D: rank r evaluates gates r, r+G, ...) -iNTT-> u_sum,
// v_sum, c = p_lo + p_hi (S) -coset NTT-> u, v on the coset (D) -product, iNTT-> d (S) -> h (S).  Three exchanges per
// proof (3, 2 and 1 vectors of m elements per rank), fused with the kernels on both sides: k_shard_mid finishes three
// inverse transforms from what exchange 0 delivered, starts two forward ones and stores their results straight into the
// peers' windows.  u_sum, v_sum and h come out in layout S, so the CRS vectors xi / xi_t are sharded the same way
// (crs.cu, layout 1): any partition of an MSM's terms gives the same sum, and results are compared in affine form.
// tests/shard_model.py is the CPU model of exactly these steps (checked against the reference restatement).
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include "common.cuh"

namespace zkb {

// ---- flags -----------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ uint32_t ld_flag(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_flag(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// Threads g < world poll flags[ch][step][g] of the own window until every rank has published `epoch`.  A wait that outlives
// the timeout sets the sticky status word instead of hanging the GPU: the proof then comes out wrong and the host reports
// ZKB_ERR_COMM.  Waits are ALWAYS their own one-block kernel (k_comm_wait) in front of the consumer, never a prologue of the
// consumer's blocks: with several proofs in flight, a grid of waiting blocks could fill every SM while the kernels that
// would release it -- another lane's sends, here and on the peer -- find no SM to run on (seen on 2 GPUs: both ranks timed out).
__device__ __forceinline__ void comm_wait_all(const CommView& cv, size_t flag_off0, uint32_t epoch) {
  if ((int)threadIdx.x < cv.world) {
    char* own = cv.base[cv.rank];
    int* status = reinterpret_cast<int*>(own + COMM_STATUS_OFF);
    const uint32_t* f = reinterpret_cast<const uint32_t*>(own + flag_off0) + threadIdx.x;
    const unsigned long long t0 = globaltimer_ns();
    while ((int32_t)(ld_flag(f) - epoch) < 0) {
      if (*reinterpret_cast<volatile int*>(status) != 0) break;
      if (globaltimer_ns() - t0 > cv.timeout_ns) { atomicExch(status, 1); break; }
      __nanosleep(100);
    }
    __threadfence_system();
  }
  __syncthreads();
}
__global__ void k_comm_wait(CommView cv, size_t flag_off0, uint32_t epoch) { comm_wait_all(cv, flag_off0, epoch); }
// after the producer kernel (stream order): publish `epoch` in every rank's flags[ch][step][this rank]
__global__ void k_comm_signal(CommView cv, size_t flag_off, uint32_t epoch) {
  if ((int)threadIdx.x < cv.world) {
    __threadfence_system();
    st_flag(reinterpret_cast<uint32_t*>(cv.base[threadIdx.x] + flag_off), epoch);
  }
}

__device__ __forceinline__ Fr ld_fr_cg(const Fr* p) {  // L2-coherent loads of data a peer wrote
  Fr a;
  const uint4* q = reinterpret_cast<const uint4*>(p);
  *reinterpret_cast<uint4*>(&a.v[0]) = __ldcg(q);
  *reinterpret_cast<uint4*>(&a.v[4]) = __ldcg(q + 1);
  return a;
}

// ---- the pieces of a distributed transform ------------------------------------------------------------
struct SmallTw { Fr w[8]; };  // w[k] = (root of order G)^k, k < G/2

// in-register size-2^LG transform (decimation in frequency): on return x[p] = X[bitrev(p)]
template <int LG>
__device__ __forceinline__ void small_dft(Fr (&x)[1 << LG], const SmallTw& tw) {
#pragma unroll
  for (int s = LG - 1; s >= 0; s--) {
    const int half = 1 << s, stride = (1 << LG) >> (s + 1);
#pragma unroll
    for (int b = 0; b < (1 << LG); b += 2 * half)
#pragma unroll
      for (int i = 0; i < half; i++) {
        const Fr a = x[b + i], c = x[b + i + half];
        x[b + i] = a + c;
        const Fr d = a - c;
        x[b + i + half] = i == 0 ? d : d * tw.w[i * stride];
      }
  }
}
template <int LG> __host__ __device__ constexpr int brev(int p) {
  int r = 0;
  for (int i = 0; i < LG; i++) r |= ((p >> i) & 1) << (LG - 1 - i);
  return r;
}

// all-to-all, sender side: vector v (nvec of them, `vstride` apart; local size-m transform results, bit-reversed order
// when BR) element k2 = d*q + t goes to rank d's region at [v][this rank][t]
template <bool BR>
__global__ void k_shard_send(const Fr* __restrict__ Y, size_t vstride, int nvec, uint32_t log_m, uint32_t q, CommView cv,
                             size_t region_off) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t m = (size_t)1 << log_m;
  if (idx >= m * nvec) return;
  const size_t v = idx >> log_m, k2 = idx & (m - 1);
  const size_t src = BR ? (log_m ? (size_t)(__brevll((unsigned long long)k2) >> (64 - log_m)) : 0) : k2;
  const Fr val = Y[v * vstride + src];
  const size_t d = k2 / q, t = k2 - d * q;
  Fr* dst = reinterpret_cast<Fr*>(cv.base[d] + region_off) + (v * cv.world + cv.rank) * q + t;
  *reinterpret_cast<uint4*>(&dst->v[0]) = *reinterpret_cast<const uint4*>(&val.v[0]);
  *reinterpret_cast<uint4*>(&dst->v[4]) = *reinterpret_cast<const uint4*>(&val.v[4]);
  __threadfence_system();
}

// receiver side of a DIT-distributed transform for column t: y[p] = X[k2 + m*brev(p)], k2 = rank*q + t
template <int LG>
__device__ __forceinline__ void combine_column(const Fr* __restrict__ R /* [G][q] */, const Fr* __restrict__ T /* [G][q] */, uint32_t q,
                                               uint32_t t, const SmallTw& tw, Fr (&y)[1 << LG]) {
#pragma unroll
  for (int g = 0; g < (1 << LG); g++) y[g] = ld_fr_cg(R + (size_t)g * q + t) * T[(size_t)g * q + t];
  small_dft<LG>(y, tw);
}

struct MidArgs {
  size_t r0_off, r1_off;  // own window: exchange-0 region; peers' windows: exchange-1 region
  uint32_t q;
  const Fr *Tinv, *Tfwd, *cosS;
  Fr *un, *vn, *cS;  // layout S outputs (Montgomery): u_sum, v_sum, c = iNTT(A.B)
};
// grid.y = 0: A -> u_sum -> coset -> first half of the forward transform -> peers; 1: the same for B / v_sum;
// 2: A.B -> c (stays local)
template <int LG>
__global__ void __launch_bounds__(128) k_shard_mid(CommView cv, MidArgs a, SmallTw tw_inv, SmallTw tw_fwd) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= a.q) return;
  constexpr int G = 1 << LG;
  const int v = blockIdx.y;
  const Fr* R0 = reinterpret_cast<const Fr*>(cv.base[cv.rank] + a.r0_off) + (size_t)v * G * a.q;
  Fr y[G];
  combine_column<LG>(R0, a.Tinv, a.q, t, tw_inv, y);
  Fr* out = v == 0 ? a.un : (v == 1 ? a.vn : a.cS);
#pragma unroll
  for (int p = 0; p < G; p++) out[(size_t)brev<LG>(p) * a.q + t] = y[p];
  if (v == 2) return;
  Fr z[G];
#pragma unroll
  for (int p = 0; p < G; p++) z[brev<LG>(p)] = y[p] * a.cosS[(size_t)brev<LG>(p) * a.q + t];
  small_dft<LG>(z, tw_fwd);
#pragma unroll
  for (int p = 0; p < G; p++) {
    const int k1 = brev<LG>(p);
    const Fr val = z[p] * a.Tfwd[(size_t)k1 * a.q + t];
    Fr* dst = reinterpret_cast<Fr*>(cv.base[k1] + a.r1_off) + ((size_t)v * G + cv.rank) * a.q + t;
    *reinterpret_cast<uint4*>(&dst->v[0]) = *reinterpret_cast<const uint4*>(&val.v[0]);
    *reinterpret_cast<uint4*>(&dst->v[4]) = *reinterpret_cast<const uint4*>(&val.v[4]);
  }
  __threadfence_system();
}

// last exchange: d = iNTT(coset product) (layout S), h = c / 2 - d * g^-j / 2
template <int LG>
__global__ void __launch_bounds__(128) k_shard_fin(CommView cv, size_t r2_off, uint32_t q,
                                                   const Fr* __restrict__ Tinv, const Fr* __restrict__ QS, const Fr* __restrict__ cS, Fr inv2,
                                                   SmallTw tw_inv, Fr* __restrict__ hn) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= q) return;
  constexpr int G = 1 << LG;
  Fr y[G];
  combine_column<LG>(reinterpret_cast<const Fr*>(cv.base[cv.rank] + r2_off), Tinv, q, t, tw_inv, y);
#pragma unroll
  for (int p = 0; p < G; p++) {
    const size_t s = (size_t)brev<LG>(p) * q + t;
    hn[s] = cS[s] * inv2 - y[p] * QS[s];
  }
}

// stand-alone distributed transform, receiver side: out (layout S, canonical residues)
template <int LG>
__global__ void __launch_bounds__(128) k_shard_combine(CommView cv, size_t r_off, uint32_t q,
                                                       const Fr* __restrict__ T, SmallTw tw, Fr* __restrict__ out) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= q) return;
  constexpr int G = 1 << LG;
  Fr y[G];
  combine_column<LG>(reinterpret_cast<const Fr*>(cv.base[cv.rank] + r_off), T, q, t, tw, y);
#pragma unroll
  for (int p = 0; p < G; p++) out[(size_t)brev<LG>(p) * q + t] = from_mont(y[p]);
}

// idx = g*q + t (= local S index with k1 = g):  Tinv = w^-(g k2) / n, Tfwd = w^(g k2), k2 = rank*q + t;
// cosS = g2n^j, QS = g2n^-j / 2, j = k2 + m*g
__global__ void k_shard_tables(Fr* Tinv, Fr* Tfwd, Fr* cosS, Fr* QS, Fr w, Fr winv, Fr g2n, Fr g2ninv, Fr invn, Fr inv2, int rank,
                               uint32_t q, size_t m) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= m) return;
  const uint64_t g = idx / q, t = idx - g * q, k2 = (uint64_t)rank * q + t, j = k2 + m * g;
  Tinv[idx] = invn * pow_u64(winv, g * k2);
  Tfwd[idx] = pow_u64(w, g * k2);
  cosS[idx] = pow_u64(g2n, j);
  QS[idx] = inv2 * pow_u64(g2ninv, j);
}

__global__ void k_partial_send(const uint32_t* __restrict__ partial, CommView cv, size_t slot_off) {
  const int i = threadIdx.x & 63, d = threadIdx.x >> 6;
  for (int g = d; g < cv.world; g += blockDim.x >> 6)
    reinterpret_cast<uint32_t*>(cv.base[g] + slot_off)[i] = partial[i];
  __threadfence_system();
}

int combine_partials_launch(zkb_ctx* ctx, const uint32_t* d_partials, int world, size_t count, uint32_t* d_out, cudaStream_t st);

static SmallTw small_tw(uint32_t log_n, int lg, bool inverse) {
  SmallTw t;
  for (auto& x : t.w) x = Fr::one();
  if (lg >= 1) {
    const Fr wG = host_omega((uint32_t)lg, inverse);  // order G
    Fr acc = Fr::one();
    for (int k = 0; k < (1 << lg) / 2 && k < 8; k++) { t.w[k] = acc; acc = acc * wG; }
  }
  (void)log_n;
  return t;
}

static int shard_tables(zkb_ctx* ctx, zkb_comm* c, uint32_t log_n, ShardTables** out) {
  ShardTables& t = c->tabs[log_n];
  if (!t.Tinv) {
    const size_t n = (size_t)1 << log_n, m = n >> c->lg, q = m >> c->lg;
    Fr* p = nullptr;
    if (cudaMalloc(&p, 4 * m * sizeof(Fr)) != cudaSuccess) return set_err(ctx, ZKB_ERR_ALLOC, "shard tables: cudaMalloc failed");
    const Fr w = host_omega(log_n, false), winv = host_omega(log_n, true);
    const Fr g = host_omega(log_n + 1, false), ginv = host_omega(log_n + 1, true);
    const Fr invn = inverse(fr_from_u64(n)), inv2 = inverse(fr_from_u64(2));
    k_shard_tables<<<cdiv(m, 128), 128, 0, ctx->stream>>>(p, p + m, p + 2 * m, p + 3 * m, w, winv, g, ginv, invn, inv2, c->rank,
                                                          (uint32_t)q, m);
    ctx->launches++;
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
      cudaFree(p);
      return set_err(ctx, ZKB_ERR_CUDA, "shard tables: %s", cudaGetErrorString(cudaGetLastError()));
    }
    t.Tinv = p; t.Tfwd = p + m; t.cosS = p + 2 * m; t.QS = p + 3 * m;
  }
  *out = &t;
  return ZKB_OK;
}

int shard_prepare(zkb_ctx* ctx, zkb_comm* c, uint32_t log_n) {
  ShardTables* T;
  ZKB_TRY(shard_tables(ctx, c, log_n, &T));
  Fr* tw;
  ZKB_TRY(get_twiddles(ctx, log_n - c->lg, false, &tw));
  return get_twiddles(ctx, log_n - c->lg, true, &tw);
}

int shard_check(zkb_ctx* ctx, const zkb_comm* c, uint32_t log_n) {
  if (!c || !c->connected) return set_err(ctx, ZKB_ERR_ARG, "comm is not connected (zkb_comm_connect)");
  if (c->ctx != ctx) return set_err(ctx, ZKB_ERR_ARG, "comm belongs to another context");
  if (log_n > c->max_log_n) return set_err(ctx, ZKB_ERR_ARG, "comm was created for transforms up to 2^%u, got 2^%u", c->max_log_n, log_n);
  if (log_n < (uint32_t)(2 * c->lg) + 1 || log_n > 27)
    return set_err(ctx, ZKB_ERR_UNSUPPORTED, "sharded transform of size 2^%u over %d ranks needs n >= 2 * world^2", log_n, c->world);
  return ZKB_OK;
}

template <int LG>
static int launch_mid(zkb_ctx* ctx, const zkb_comm* c, const MidArgs& a, size_t flag0_off, uint32_t epoch, cudaStream_t st) {
  ZKB_LAUNCH(ctx, k_comm_wait, 1, 32, 0, st, c->view, flag0_off, epoch);
  dim3 grid(cdiv(a.q, 128), 3);
  ZKB_LAUNCH(ctx, k_shard_mid<LG>, grid, 128, 0, st, c->view, a, small_tw(0, LG, true), small_tw(0, LG, false));
  return ZKB_OK;
}
template <int LG>
static int launch_fin(zkb_ctx* ctx, const zkb_comm* c, size_t r2, size_t f2, uint32_t epoch, uint32_t q, const ShardTables& T, Fr* cS,
                      Fr* hn, cudaStream_t st) {
  ZKB_LAUNCH(ctx, k_comm_wait, 1, 32, 0, st, c->view, f2, epoch);
  ZKB_LAUNCH(ctx, k_shard_fin<LG>, cdiv(q, 128), 128, 0, st, c->view, r2, q, T.Tinv, T.QS, cS, inverse(fr_from_u64(2)),
             small_tw(0, LG, true), hn);
  return ZKB_OK;
}
template <int LG>
static int launch_combine(zkb_ctx* ctx, const zkb_comm* c, size_t r_off, size_t f_off, uint32_t epoch, uint32_t q, const Fr* T, bool inv,
                          Fr* out, cudaStream_t st) {
  ZKB_LAUNCH(ctx, k_comm_wait, 1, 32, 0, st, c->view, f_off, epoch);
  ZKB_LAUNCH(ctx, k_shard_combine<LG>, cdiv(q, 128), 128, 0, st, c->view, r_off, q, T, small_tw(0, LG, inv), out);
  return ZKB_OK;
}
#define ZKB_BY_LG(lg, call)                                                                                          \
  ((lg) == 0 ? call<0> : (lg) == 1 ? call<1> : (lg) == 2 ? call<2> : (lg) == 3 ? call<3> : call<4>)

static int signal(zkb_ctx* ctx, const zkb_comm* c, int ch, int step, uint32_t epoch, cudaStream_t st) {
  ZKB_LAUNCH(ctx, k_comm_signal, 1, 32, 0, st, c->view, comm_flag_off(ch, step, c->rank), epoch);
  return ZKB_OK;
}

// k_matvec of prove.cu over the gates k0, k0 + stride, ...
int matvec_launch(zkb_ctx* ctx, const zkb_qap* q, const Fr* wmont, size_t k0, size_t kstride, size_t count, Fr* A, Fr* B, Fr* AB,
                  cudaStream_t st);

int shard_poly_stage(zkb_ctx* ctx, zkb_comm* c, int ch, uint32_t epoch, const zkb_qap* qp, Fr* ws, const Fr* wmont, cudaStream_t st) {
  const uint32_t log_n = qp->log_n, log_m = log_n - c->lg;
  const size_t n = qp->n, m = n >> c->lg, q = m >> c->lg;
  ShardTables* T;
  ZKB_TRY(shard_tables(ctx, c, log_n, &T));
  Fr *A = ws, *B = ws + m, *AB = ws + 2 * m, *un = ws + 3 * m, *vn = ws + 4 * m, *cS = ws + 5 * m, *hn = ws + 6 * m;
  ZKB_TRY(matvec_launch(ctx, qp, wmont, (size_t)c->rank, (size_t)c->world, m, A, B, AB, st));
  ZKB_TRY(ntt_dif(ctx, A, log_m, true, st));  // m * Y_r, bit-reversed
  ZKB_TRY(ntt_dif(ctx, B, log_m, true, st));
  ZKB_TRY(ntt_dif(ctx, AB, log_m, true, st));
  ZKB_LAUNCH(ctx, k_shard_send<true>, cdiv(3 * m, 256), 256, 0, st, A, m, 3, log_m, (uint32_t)q, c->view, comm_region_off(c, ch, 0));
  ZKB_TRY(signal(ctx, c, ch, 0, epoch, st));
  MidArgs a;
  a.r0_off = comm_region_off(c, ch, 0); a.r1_off = comm_region_off(c, ch, 1);
  a.q = (uint32_t)q;
  a.Tinv = T->Tinv; a.Tfwd = T->Tfwd; a.cosS = T->cosS;
  a.un = un; a.vn = vn; a.cS = cS;
  ZKB_TRY(ZKB_BY_LG(c->lg, launch_mid)(ctx, c, a, comm_flag_off(ch, 0, 0), epoch, st));
  ZKB_TRY(signal(ctx, c, ch, 1, epoch, st));
  ZKB_LAUNCH(ctx, k_comm_wait, 1, 32, 0, st, c->view, comm_flag_off(ch, 1, 0), epoch);
  Fr* Wu = reinterpret_cast<Fr*>(c->window + comm_region_off(c, ch, 1));  // [vector][src rank][q] = W in natural order
  Fr* Wv = Wu + (size_t)c->world * q;
  ZKB_TRY(ntt_dif(ctx, Wu, log_m, false, st));  // u_sum on the coset points of this rank, bit-reversed local order
  ZKB_TRY(ntt_dif(ctx, Wv, log_m, false, st));
  ZKB_TRY(vec_mul(ctx, Wv, Wu, Wv, m, st));
  ZKB_TRY(ntt_dit(ctx, Wv, log_m, true, st));   // natural order
  ZKB_LAUNCH(ctx, k_shard_send<false>, cdiv(m, 256), 256, 0, st, Wv, m, 1, log_m, (uint32_t)q, c->view, comm_region_off(c, ch, 2));
  ZKB_TRY(signal(ctx, c, ch, 2, epoch, st));
  ZKB_TRY(ZKB_BY_LG(c->lg, launch_fin)(ctx, c, comm_region_off(c, ch, 2), comm_flag_off(ch, 2, 0), epoch, (uint32_t)q, *T, cS, hn, st));
  return ZKB_OK;
}

int shard_exchange_partials(zkb_ctx* ctx, zkb_comm* c, int ch, uint32_t epoch, const uint32_t* partial, uint32_t* out, int* d_status_copy,
                            cudaStream_t st) {
  ZKB_LAUNCH(ctx, k_partial_send, 1, 256, 0, st, partial, c->view, comm_slot_off(ch, c->rank));
  ZKB_TRY(signal(ctx, c, ch, 3, epoch, st));
  ZKB_LAUNCH(ctx, k_comm_wait, 1, 32, 0, st, c->view, comm_flag_off(ch, 3, 0), epoch);
  ZKB_TRY(combine_partials_launch(ctx, reinterpret_cast<const uint32_t*>(c->window + comm_slot_off(ch, 0)), c->world, 1, out, st));
  if (d_status_copy)
    ZKB_CUDA(ctx, cudaMemcpyAsync(d_status_copy, c->window + COMM_STATUS_OFF, sizeof(int), cudaMemcpyDeviceToDevice, st));
  return ZKB_OK;
}

}  // namespace zkb

using namespace zkb;

// ---- handles -------------------------------------------------------------------------------------------
struct CommHandle {  // ZKB_COMM_HANDLE_BYTES = 128
  uint32_t magic, rank, world, device;
  uint64_t pid, ptr, bytes, max_log_n;
  cudaIpcMemHandle_t ipc;  // 64 bytes
  uint8_t pad[16];
};
static_assert(sizeof(CommHandle) == 128, "handle layout");
static const uint32_t COMM_MAGIC = 0x7a6b6232u;

extern "C" {

int zkb_comm_create(zkb_ctx* ctx, int rank, int world, uint32_t max_log_n, zkb_comm** out, uint8_t* handle) {
  if (!ctx || !out || !handle) return set_err(ctx, ZKB_ERR_ARG, "zkb_comm_create: NULL argument");
  *out = nullptr;
  int lg = 0;
  while ((1 << lg) < world) lg++;
  if (world < 1 || world > ZKB_COMM_MAX_WORLD || (1 << lg) != world || rank < 0 || rank >= world)
    return set_err(ctx, ZKB_ERR_ARG, "zkb_comm_create: world must be a power of two <= %d and 0 <= rank < world", ZKB_COMM_MAX_WORLD);
  if (max_log_n < (uint32_t)(2 * lg) + 1 || max_log_n > 27)
    return set_err(ctx, ZKB_ERR_ARG, "zkb_comm_create: max_log_n %u out of range [%d, 27]", max_log_n, 2 * lg + 1);
  ZKB_CUDA(ctx, cudaSetDevice(ctx->device));
  zkb_comm* c = new zkb_comm();
  c->ctx = ctx; c->rank = rank; c->world = world; c->lg = lg; c->max_log_n = max_log_n;
  c->m_max = ((size_t)1 << max_log_n) >> lg;
  c->window_bytes = COMM_DATA_OFF + (size_t)ZKB_COMM_CHANNELS * 6 * c->m_max * sizeof(Fr);
  if (cudaMalloc(&c->window, c->window_bytes) != cudaSuccess) {
    delete c;
    return set_err(ctx, ZKB_ERR_ALLOC, "zkb_comm_create: cudaMalloc of the %.1f MiB exchange window failed", c->window_bytes / 1048576.0);
  }
  cudaMemsetAsync(c->window, 0, COMM_DATA_OFF, ctx->stream);
  if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
    cudaFree(c->window);
    delete c;
    return set_err(ctx, ZKB_ERR_CUDA, "zkb_comm_create: %s", cudaGetErrorString(cudaGetLastError()));
  }
  CommHandle h;
  memset(&h, 0, sizeof h);
  h.magic = COMM_MAGIC; h.rank = (uint32_t)rank; h.world = (uint32_t)world; h.device = (uint32_t)ctx->device;
  h.pid = (uint64_t)getpid(); h.ptr = (uint64_t)(uintptr_t)c->window; h.bytes = c->window_bytes; h.max_log_n = max_log_n;
  if (world > 1 && cudaIpcGetMemHandle(&h.ipc, c->window) != cudaSuccess) {
    cudaGetLastError();  // peers inside this process still work through the raw pointer
    memset(&h.ipc, 0, sizeof h.ipc);
  }
  memcpy(handle, &h, sizeof h);
  for (auto& e : c->epoch) e = 0;
  memset(&c->view, 0, sizeof c->view);
  c->view.rank = rank; c->view.world = world; c->view.lg = lg;
  double tmo_ms = 20000.0;
  if (const char* e = getenv("ZKB_COMM_TIMEOUT_MS")) tmo_ms = atof(e);
  c->view.timeout_ns = (unsigned long long)(tmo_ms * 1e6);
  c->view.base[rank] = c->window;
  if (world == 1) c->connected = true;
  *out = c;
  return ZKB_OK;
}

int zkb_comm_connect(zkb_comm* c, const uint8_t* handles) {
  if (!c || !handles) return set_err(c ? c->ctx : nullptr, ZKB_ERR_ARG, "zkb_comm_connect: NULL argument");
  zkb_ctx* ctx = c->ctx;
  ZKB_CUDA(ctx, cudaSetDevice(ctx->device));
  for (int g = 0; g < c->world; g++) {
    CommHandle h;
    memcpy(&h, handles + (size_t)g * sizeof(CommHandle), sizeof h);
    if (h.magic != COMM_MAGIC || (int)h.rank != g || (int)h.world != c->world || h.bytes != c->window_bytes || h.max_log_n != c->max_log_n)
      return set_err(ctx, ZKB_ERR_ARG, "zkb_comm_connect: handle %d does not describe rank %d of this communicator", g, g);
    if (g == c->rank) continue;
    if (h.pid == (uint64_t)getpid()) {  // same process: the pointer itself (other device: peer access)
      if ((int)h.device != ctx->device) {
        cudaError_t e = cudaDeviceEnablePeerAccess((int)h.device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
          return set_err(ctx, ZKB_ERR_UNSUPPORTED, "zkb_comm_connect: no peer access from device %d to %u: %s", ctx->device, h.device,
                         cudaGetErrorString(e));
        cudaGetLastError();
      }
      c->view.base[g] = reinterpret_cast<char*>((uintptr_t)h.ptr);
    } else {
      void* p = nullptr;
      cudaError_t e = cudaIpcOpenMemHandle(&p, h.ipc, cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess)
        return set_err(ctx, ZKB_ERR_UNSUPPORTED, "zkb_comm_connect: cudaIpcOpenMemHandle of rank %d's window failed: %s", g,
                       cudaGetErrorString(e));
      c->view.base[g] = (char*)p;
      c->peer_is_ipc[g] = true;
    }
  }
  c->connected = true;
  return ZKB_OK;
}

void zkb_comm_destroy(zkb_comm* c) {
  if (!c) return;
  if (c->ctx) {
    cudaSetDevice(c->ctx->device);
    cudaDeviceSynchronize();
  }
  for (int g = 0; g < c->world; g++)
    if (c->peer_is_ipc[g] && c->view.base[g]) cudaIpcCloseMemHandle(c->view.base[g]);
  for (auto& t : c->tabs)
    if (t.Tinv) cudaFree(t.Tinv);
  cudaFree(c->window);
  delete c;
}

int zkb_comm_info(const zkb_comm* c, int* rank, int* world, int* status) {
  if (!c) return ZKB_ERR_ARG;
  if (rank) *rank = c->rank;
  if (world) *world = c->world;
  if (status) {
    *status = 0;
    cudaSetDevice(c->ctx->device);
    if (cudaMemcpy(status, c->window + COMM_STATUS_OFF, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess)
      return set_err(c->ctx, ZKB_ERR_CUDA, "zkb_comm_info: %s", cudaGetErrorString(cudaGetLastError()));
  }
  return ZKB_OK;
}

// One size-2^log_n transform over all ranks of `comm` (reference convention, field/mod.rs:508-537).  d_local holds this
// rank's m = n / world elements, canonical: in layout D (x[rank + world * i]) on entry, in layout S on return
// (index s = k1 * q + t <-> X[(rank * q + t) + m * k1], q = m / world).
int zkb_ntt_shard(zkb_ctx* ctx, zkb_comm* c, uint64_t* d_local, uint32_t log_n, int inverse_, int no_wait) {
  if (!ctx || !c || !d_local) return set_err(ctx, ZKB_ERR_ARG, "zkb_ntt_shard: NULL argument");
  ZKB_TRY(shard_check(ctx, c, log_n));
  ZKB_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const uint32_t log_m = log_n - c->lg;
  const size_t m = (size_t)1 << log_m, q = m >> c->lg;
  const int ch = ZKB_COMM_CHANNELS - 1;
  ShardTables* T;
  ZKB_TRY(shard_prepare(ctx, c, log_n));
  ZKB_TRY(shard_tables(ctx, c, log_n, &T));
  const uint32_t epoch = ++c->epoch[ch];
  Fr* d = (Fr*)d_local;
  ZKB_TRY(vec_to_mont(ctx, d, m, true, st));
  ZKB_TRY(ntt_dif(ctx, d, log_m, inverse_ != 0, st));
  // two receive buffers by epoch parity: a peer may start transform e+1 while this rank still reads the data of e, but
  // not e+2 (that needs this rank's send of e+1, which is stream-ordered after its reads of e)
  const size_t roff = comm_region_off(c, ch, 0) + (size_t)(epoch & 1) * 3 * c->m_max * sizeof(Fr);
  ZKB_LAUNCH(ctx, k_shard_send<true>, cdiv(m, 256), 256, 0, st, d, m, 1, log_m, (uint32_t)q, c->view, roff);
  ZKB_TRY(signal(ctx, c, ch, 0, epoch, st));
  ZKB_TRY(ZKB_BY_LG(c->lg, launch_combine)(ctx, c, roff, comm_flag_off(ch, 0, 0), epoch, (uint32_t)q,
                                           inverse_ ? T->Tinv : T->Tfwd, inverse_ != 0, d, st));
  if (!no_wait) {
    ZKB_CUDA(ctx, cudaStreamSynchronize(st));
    int status = 0;
    ZKB_TRY(zkb_comm_info(c, nullptr, nullptr, &status));
    if (status) return set_err(ctx, ZKB_ERR_COMM, "zkb_ntt_shard: a peer did not arrive within the exchange timeout");
  }
  return ZKB_OK;
}

}  // extern "C"
