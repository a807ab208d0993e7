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
// Twiddles omega^k (k < n/2) live in a per-size table (L2-resident: 16 MB at 2^20).  Twiddle TILES: beside the flat
// table every pass has a tiled copy in which the (2^B - 1) * C twiddles ONE block needs are contiguous (two 16-byte
// planes, stage-major), so a block stages them with one bulk-async copy (TMA: cp.async.bulk + mbarrier) that
// overlaps its data-tile loads, and the butterflies read them with conflict-free LDS.128 instead of 32-byte gathers
// from L2 (ZKB_NTT_TMA=0..3 selects the policy at run time, see ntt_tma_mode; all are bit-identical).
#include <stdlib.h>
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

// ---- twiddle tiles ---------------------------------------------------------------------------------
// Tile of group g (= column group Lbase / C of a pass) holds, for ls = 0..B-1, m0 < 2^ls, c < C:
//   slot ((2^ls - 1) + m0) * C + c  =  omega_n^(((m0 << lo) + g*C + c) << (log_n - 1 - lo - ls))
// as two planes of T uint4 (limbs 0..3 | limbs 4..7); slots >= T - C are padding.
__global__ void k_fill_tw_tiles(uint4* __restrict__ out, const Fr* __restrict__ tw, uint32_t log_n, uint32_t lo, uint32_t B,
                                uint32_t logC, size_t total) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const uint32_t logT = B + logC, T = 1u << logT, C = 1u << logC;
  const uint32_t slot = (uint32_t)(i & (T - 1));
  const uint64_t g = i >> logT;
  uint4 a = make_uint4(0, 0, 0, 0), b = a;
  if (slot < T - C) {
    const uint32_t idx = slot >> logC, c = slot & (C - 1);
    const uint32_t ls = 31 - __clz(idx + 1), m0 = idx + 1 - (1u << ls);
    const uint64_t j = ((uint64_t)m0 << lo) + (g << logC) + c;
    const Fr w = tw[j << (log_n - 1 - lo - ls)];
    a = *reinterpret_cast<const uint4*>(&w.v[0]);
    b = *reinterpret_cast<const uint4*>(&w.v[4]);
  }
  out[(g * 2) * T + slot] = a;
  out[(g * 2 + 1) * T + slot] = b;
}

template <bool DIT, bool TMA, bool SHF = false>
__global__ void k_ntt_pass(Fr* __restrict__ d, const Fr* __restrict__ tw, const uint4* __restrict__ tiles, uint32_t log_n, uint32_t hi,
                           uint32_t lo, uint32_t logC);

#ifndef ZKB_NTT_TMA_DEFAULT
#define ZKB_NTT_TMA_DEFAULT 3
#endif
// 0: 32-byte gathers from the flat table in every pass; 1: tiles in every pass; 2: tiles only in passes whose whole
// tile table is at most 4 MB (the low passes: few column groups, every block of a group re-reads the same tile from L2);
// 3 (default): tiles in every pass of transforms up to 2^18 (two passes), gathers above.  Measured on B200
// (profiles/r01_ntt_tma_modes.txt, ms per transform, modes 0 / 1 / 2): 2^12 .0431/.0409/.0409, 2^16 .0519/.0492/.0492,
// 2^18 .0749/.0716/.0727, 2^20 .2200/.2218/.2211, 2^22 .8308/.8425/.8386, 2^24 3.494/3.541/3.539 -- staging wins 4-5 %
// while the transform is latency-bound (two passes), and loses ~1 % once three passes keep the multiplier busy (the
// 64 KB of shared memory per block instead of 32 KB leaves less L1 for the data tiles).
static int ntt_tma_mode() {
  static const int v = getenv("ZKB_NTT_TMA") ? atoi(getenv("ZKB_NTT_TMA")) : ZKB_NTT_TMA_DEFAULT;
  return v;
}
static bool ntt_pass_uses_tiles(const Pass& p, uint32_t log_n) {
  const int mode = ntt_tma_mode();
  if (mode == 0) return false;
  if (mode == 3) return log_n <= 18;
  const size_t groups = ((size_t)1 << p.lo) >> p.logC;
  return mode == 1 || (groups << (p.hi - p.lo + p.logC)) * 32 <= ((size_t)4 << 20);
}

int get_twiddles(zkb_ctx* ctx, uint32_t log_n, bool inverse, Fr** out) {
  if (log_n < 1 || log_n > 27) return set_err(ctx, ZKB_ERR_ARG, "ntt size 2^%u unsupported", log_n);
  Fr*& slot = ctx->tw[log_n][inverse ? 1 : 0];
  if (!slot) {
    // Built into locals and published to ctx->tw / ctx->twt only after the fill kernels have completed: a failed
    // allocation or launch leaves the cache untouched (and frees what was allocated), so a later call cannot take a
    // half-built table for a finished one or race the fill on another lane's stream.
    const size_t half = (size_t)1 << (log_n - 1);
    Fr* t = nullptr;
    uint4* tiles[4] = {nullptr, nullptr, nullptr, nullptr};
    auto build = [&]() -> int {
      ZKB_CUDA(ctx, cudaMalloc(&t, half * sizeof(Fr)));
      ZKB_LAUNCH(ctx, k_fill_powers, cdiv(half, 256), 256, 0, ctx->stream, t, host_omega(log_n, inverse), Fr::one(), half);
      std::vector<Pass> ps = plan(log_n);
      for (size_t q = 0; q < ps.size() && q < 4; q++) {
        const Pass& p = ps[q];
        if (!ntt_pass_uses_tiles(p, log_n)) continue;
        const uint32_t logT = p.hi - p.lo + p.logC;
        const size_t groups = ((size_t)1 << p.lo) >> p.logC, total = groups << logT;
        ZKB_CUDA(ctx, cudaMalloc(&tiles[q], total * 32));
        // 64 KB of dynamic shared memory (data tile + twiddle tile) needs the opt-in, per device
        ZKB_CUDA(ctx, cudaFuncSetAttribute(k_ntt_pass<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
        ZKB_CUDA(ctx, cudaFuncSetAttribute(k_ntt_pass<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
        ZKB_LAUNCH(ctx, k_fill_tw_tiles, cdiv(total, 256), 256, 0, ctx->stream, tiles[q], t, log_n, p.lo, p.hi - p.lo, p.logC, total);
      }
      ZKB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      return ZKB_OK;
    };
    const int rc = build();
    if (rc != ZKB_OK) {
      cudaStreamSynchronize(ctx->stream);
      cudaFree(t);
      for (auto& p : tiles) cudaFree(p);
      return rc;
    }
    for (int q = 0; q < 4; q++) ctx->twt[log_n][inverse ? 1 : 0][q] = tiles[q];
    slot = t;
  }
  *out = slot;
  return ZKB_OK;
}

// ------------------------------------------------------------------------------------------------
// one pass = stages [lo, hi) of the size-2^log_n transform on 2^(hi-lo+logC)-element tiles.
// DIF (Gentleman-Sande) runs the stages downwards, DIT (Cooley-Tukey) upwards.
//
// Radix-4 rounds in registers: a thread owns the four tile elements e0 + {0, h, 2h, 3h} of two
// consecutive stages (half distances h and 2h), so a round trip through shared memory and a barrier
// buy TWO stages, and the two butterflies of a stage are independent multiplies in one thread.  The
// multiplication count is that of radix 2 (the "multiply by i" of a radix-4 butterfly is a generic
// Fr product): 4 per round.  Twiddles of a round: w_a for both butterflies of the lower stage s;
// w_b and w_c = w_b * omega^(n/4) (read from the same table) for the upper stage.  An odd last
// stage is a plain radix-2 step, two butterflies per thread.  Stage 0 (all twiddles 1) skips its
// products.
__device__ __forceinline__ Fr lds_fr(const uint4* p0, const uint4* p1, uint32_t e) {
  Fr a;
  *reinterpret_cast<uint4*>(&a.v[0]) = p0[e];
  *reinterpret_cast<uint4*>(&a.v[4]) = p1[e];
  return a;
}
__device__ __forceinline__ void sts_fr(uint4* p0, uint4* p1, uint32_t e, const Fr& x) {
  p0[e] = *reinterpret_cast<const uint4*>(&x.v[0]);
  p1[e] = *reinterpret_cast<const uint4*>(&x.v[4]);
}

// Measured on B200 (profiles/r01_notes.md): the two radix-4 groups of a thread unrolled side by side (two
// independent multiply chains per thread: the kernel stalls on fixed-latency dependencies, ncu "wait") at
// >= 3 blocks per SM: 0.236 -> 0.220 ms at 2^20; capping registers for 6-7 blocks per SM instead: 0.226-0.229.
#ifndef ZKB_NTT_MIN_BLOCKS
#define ZKB_NTT_MIN_BLOCKS 3
#endif
#ifndef ZKB_NTT_NO_UNROLL
#define ZKB_NTT_UNROLL2 1
#endif
// bulk-async (TMA) staging of one contiguous global range into shared memory, completion on an mbarrier
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // make the init visible to the async proxy
}
__device__ __forceinline__ void bulk_load(void* dst_smem, const void* src_gmem, uint32_t bytes, unsigned long long* bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
  } while (!ok);
}

// TMA = true: `tiles` is this pass's tiled twiddle table (k_fill_tw_tiles); thread 0 issues one bulk copy of the
// block's 32*T-byte tile into the upper half of shared memory before the data tile is loaded, every thread waits
// on the mbarrier after the data barrier, and twiddles are LDS.128 pairs.  TMA = false: 32-byte gathers from `tw`.
//
// SHF = true (DIF, the contiguous 1024-element pass only; ZKB_NTT_SHFL=1): the warp-shuffle variant the north star names.
// A thread keeps its radix-4 groups in registers over the last three rounds (stages 5..0); between rounds the four
// threads that differ in two LANE bits (q bits [2r, 2r+1], r = 1, 0) exchange elements with a 4x4 transpose made of
// XOR shuffles instead of the round trip through shared memory and its barrier.  (The earlier rounds regroup across
// warps and cannot use shuffles.)  Bit-identical output.  Measured on B200 (profiles/r02_ntt_shuffle.txt) and NOT the
// default: a 256-bit element is 8 SHFL plus the selects that emulate register indexing (64 instructions per exchanged
// element, three per group and transition) against two STS.128 + two LDS.128, and the pass is bound by the integer
// multiplier, not by shared memory or its barriers.
template <bool DIT, bool TMA, bool SHF>
__global__ void __launch_bounds__(128, ZKB_NTT_MIN_BLOCKS) k_ntt_pass(Fr* __restrict__ d, const Fr* __restrict__ tw,
                                                  const uint4* __restrict__ tiles, uint32_t log_n,
                                                  uint32_t hi, uint32_t lo, uint32_t logC) {
  extern __shared__ uint4 smem[];
  __shared__ alignas(8) unsigned long long tw_bar;
  const uint32_t B = hi - lo;
  const uint32_t logT = B + logC;
  const uint32_t T = 1u << logT;
  uint4* p0 = smem;       // limbs 0..3 of tile element e
  uint4* p1 = smem + T;   // limbs 4..7
  const uint32_t t = threadIdx.x;
  const uint32_t nthr = blockDim.x;  // T / 4 (T / 2 for tiles of two elements)
  const uint32_t C = 1u << logC;
  const uint64_t groups_per_H = ((uint64_t)1 << lo) >> logC;
  const uint64_t H = blockIdx.x / groups_per_H;
  const uint64_t Lbase = (blockIdx.x % groups_per_H) << logC;
  const uint64_t base = (H << hi) + Lbase;
  const uint4* g4 = reinterpret_cast<const uint4*>(d);
  uint4* gw4 = reinterpret_cast<uint4*>(d);
  const uint4* tw0 = smem + 2 * T;  // TMA: twiddle planes
  const uint4* tw1 = smem + 3 * T;

  if (TMA) {
    if (t == 0) mbar_init(&tw_bar, 1);
    __syncthreads();
    if (t == 0) bulk_load(smem + 2 * T, tiles + (size_t)(blockIdx.x % groups_per_H) * 2 * T, T * 32, &tw_bar);
  }
  for (uint32_t e = t; e < T; e += nthr) {
    uint64_t gi = base + ((uint64_t)(e >> logC) << lo) + (e & (C - 1));
    p0[e] = g4[gi * 2];
    p1[e] = g4[gi * 2 + 1];
  }
  __syncthreads();
  if (TMA) mbar_wait(&tw_bar, 0);

  // twiddle of the butterfly whose lower element is tile element e, in stage s = lo + ls
  auto twiddle = [&](uint32_t e, uint32_t ls) -> Fr {
    if (TMA) {  // slot ((2^ls - 1) + m0) * C + c, with m0 * C + c = e & (2^(ls + logC) - 1)
      const uint32_t slot = (((1u << ls) - 1) << logC) + (e & ((1u << (ls + logC)) - 1));
      return lds_fr(tw0, tw1, slot);
    }
    const uint32_t s = lo + ls;
    const uint64_t m0 = (e >> logC) & ((1u << ls) - 1);
    const uint64_t j = (m0 << lo) + Lbase + (e & (C - 1));
    const uint64_t jj = (s == 0) ? 0 : (j & (((uint64_t)1 << s) - 1));
    return tw[jj << (log_n - 1 - s)];
  };

  const uint32_t pairs = B >> 1;
  const bool odd = B & 1;
  // DIT: rounds (0,1), (2,3), ..., then the odd top stage.  DIF: the odd top stage first, then rounds downwards.
  for (uint32_t step = 0; step < pairs + (odd ? 1 : 0); step++) {
    const bool single = odd && (DIT ? step == pairs : step == 0);
    if (single) {
      const uint32_t ls = B - 1;
      const uint32_t sh = ls + logC;
      const uint32_t half = 1u << sh;
      for (uint32_t q = t; q < (T >> 1); q += nthr) {
        const uint32_t e0 = ((q >> sh) << (sh + 1)) | (q & (half - 1));
        const uint32_t e1 = e0 + half;
        const Fr w = twiddle(e0, ls);
        Fr a = lds_fr(p0, p1, e0), b = lds_fr(p0, p1, e1), x, y;
        if (DIT) {
          Fr wb = (lo + ls == 0) ? b : w * b;
          x = a + wb;
          y = a - wb;
        } else {
          x = a + b;
          y = (lo + ls == 0) ? a - b : (a - b) * w;
        }
        sts_fr(p0, p1, e0, x);
        sts_fr(p0, p1, e1, y);
      }
    } else {
      const uint32_t r = DIT ? step : (pairs - 1 - (odd ? step - 1 : step));
      if (SHF && !DIT && r == 2 && lo == 0 && logC == 0 && B == 10 && nthr == 128) {
        // rounds 2, 1, 0 in registers; thread t owns the groups q = t and q = t + 128 in every round
        Fr v[2][4];
        const uint32_t lane = t & 31;
#pragma unroll
        for (int rr = 2; rr >= 0; rr--) {
          const uint32_t sh = 2 * rr, h = 1u << sh;
#pragma unroll
          for (int gidx = 0; gidx < 2; gidx++) {
            const uint32_t q = t + 128 * gidx;
            const uint32_t e0 = ((q >> sh) << (sh + 2)) | (q & (h - 1));
            if (rr == 2) {
#pragma unroll
              for (int x = 0; x < 4; x++) v[gidx][x] = lds_fr(p0, p1, e0 + x * h);
            }
            const Fr wb = twiddle(e0, sh + 1), wc = twiddle(e0 + h, sh + 1);
            Fr a0 = v[gidx][0] + v[gidx][2], a2 = (v[gidx][0] - v[gidx][2]) * wb;
            Fr a1 = v[gidx][1] + v[gidx][3], a3 = (v[gidx][1] - v[gidx][3]) * wc;
            v[gidx][0] = a0 + a1;
            v[gidx][2] = a2 + a3;
            if (rr == 0) {  // stage 0: every twiddle is 1
              v[gidx][1] = a0 - a1;
              v[gidx][3] = a2 - a3;
            } else {
              const Fr wa = twiddle(e0, sh);
              v[gidx][1] = (a0 - a1) * wa;
              v[gidx][3] = (a2 - a3) * wa;
            }
            if (rr > 0) {
              // to round rr - 1: element (thread j, local x) -> (thread x, local j) among the four lanes that differ in
              // lane bits [2 (rr - 1), 2 (rr - 1) + 1]
              const int stride = 1 << (2 * (rr - 1));
              const int j = (lane >> (2 * (rr - 1))) & 3;
#pragma unroll
              for (int k = 1; k < 4; k++) {
                const int idx = j ^ k;
                const Fr send = sel4(idx, v[gidx][0], v[gidx][1], v[gidx][2], v[gidx][3]);
                Fr recv;
#pragma unroll
                for (int i = 0; i < 8; i++) recv.v[i] = __shfl_xor_sync(0xffffffffu, send.v[i], k * stride);
#pragma unroll
                for (int x = 0; x < 4; x++)
                  if (idx == x) v[gidx][x] = recv;
              }
            } else {
#pragma unroll
              for (int x = 0; x < 4; x++) sts_fr(p0, p1, e0 + x * h, v[gidx][x]);
            }
          }
        }
        __syncthreads();
        break;
      }
      const uint32_t ls = 2 * r;  // lower stage of the round
      const uint32_t sh = ls + logC;
      const uint32_t h = 1u << sh;
#ifdef ZKB_NTT_UNROLL2
#pragma unroll 2
#endif
      for (uint32_t q = t; q < (T >> 2); q += nthr) {
        const uint32_t e0 = ((q >> sh) << (sh + 2)) | (q & (h - 1));
        const uint32_t e1 = e0 + h, e2 = e0 + 2 * h, e3 = e0 + 3 * h;
        const bool unit = (lo + ls == 0);  // stage 0: every twiddle is 1
        const Fr wa = twiddle(e0, ls);
        const Fr wb = twiddle(e0, ls + 1);
        const Fr wc = twiddle(e1, ls + 1);
        Fr x0 = lds_fr(p0, p1, e0), x1 = lds_fr(p0, p1, e1), x2 = lds_fr(p0, p1, e2), x3 = lds_fr(p0, p1, e3);
        Fr y0, y1, y2, y3;
        if (DIT) {
          Fr t1 = unit ? x1 : wa * x1;
          Fr t3 = unit ? x3 : wa * x3;
          Fr a0 = x0 + t1, a1 = x0 - t1, a2 = x2 + t3, a3 = x2 - t3;
          Fr u2 = wb * a2, u3 = wc * a3;
          y0 = a0 + u2; y2 = a0 - u2;
          y1 = a1 + u3; y3 = a1 - u3;
        } else {
          Fr a0 = x0 + x2, a2 = (x0 - x2) * wb;
          Fr a1 = x1 + x3, a3 = (x1 - x3) * wc;
          y0 = a0 + a1;
          y1 = unit ? a0 - a1 : (a0 - a1) * wa;
          y2 = a2 + a3;
          y3 = unit ? a2 - a3 : (a2 - a3) * wa;
        }
        sts_fr(p0, p1, e0, y0);
        sts_fr(p0, p1, e1, y1);
        sts_fr(p0, p1, e2, y2);
        sts_fr(p0, p1, e3, y3);
      }
    }
    __syncthreads();
  }

  for (uint32_t e = t; e < T; e += nthr) {
    uint64_t gi = base + ((uint64_t)(e >> logC) << lo) + (e & (C - 1));
    gw4[gi * 2] = p0[e];
    gw4[gi * 2 + 1] = p1[e];
  }
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
    unsigned block = T >= 4 ? T / 4 : 1;
    if (block > 128) block = 128;  // two radix-4 groups per thread per round: four 128-thread blocks per SM beat two of 256
    static const int env_threads = getenv("ZKB_NTT_THREADS") ? atoi(getenv("ZKB_NTT_THREADS")) : 0;  // developer switch
    if (env_threads >= 32 && (unsigned)env_threads < block) block = (unsigned)env_threads;
    size_t smem = (size_t)T * 32;
    if (ctx->profile) ctx->prof_units[PK_NTT] += (uint64_t)1 << log_n;
    const int pi = DIT ? q : np - 1 - q;  // index of the pass in plan order (= index of its tile table)
    const uint4* tiles = pi < 4 ? ctx->twt[log_n][inverse ? 1 : 0][pi] : nullptr;  // null: this pass gathers from the flat table
    auto* k_ntt_pass_tiles = &k_ntt_pass<DIT, true>;     // named so that traces (zkb_trace_dump) tell the two apart
    auto* k_ntt_pass_gather = &k_ntt_pass<DIT, false>;
    static const int env_shfl = getenv("ZKB_NTT_SHFL") ? atoi(getenv("ZKB_NTT_SHFL")) : 0;  // developer switch (A/B; see k_ntt_pass)
    auto* k_ntt_pass_shuffle = &k_ntt_pass<false, false, true>;
    if (env_shfl && !DIT && p.lo == 0 && p.hi == 10 && block == 128) {
      ZKB_LAUNCH_K(ctx, PK_NTT, k_ntt_pass_shuffle, grid, block, smem, st, d, tw, tiles, log_n, p.hi, p.lo, p.logC);
    } else if (tiles) {
      ZKB_LAUNCH_K(ctx, PK_NTT, k_ntt_pass_tiles, grid, block, 2 * smem, st, d, tw, tiles, log_n, p.hi, p.lo, p.logC);
    } else {
      ZKB_LAUNCH_K(ctx, PK_NTT, k_ntt_pass_gather, grid, block, smem, st, d, tw, tiles, log_n, p.hi, p.lo, p.logC);
    }
  }
  return ZKB_OK;
}

int ntt_dif(zkb_ctx* ctx, Fr* d, uint32_t log_n, bool inverse, cudaStream_t st) {
  return run_ntt<false>(ctx, d, log_n, inverse, st);
}
int ntt_dit(zkb_ctx* ctx, Fr* d, uint32_t log_n, bool inverse, cudaStream_t st) {
  return run_ntt<true>(ctx, d, log_n, inverse, st);
}

// Outer-dimension sharding across GPUs: rank g transforms the decimated subsequence
// x[g], x[g+G], ... (a size-n/G transform, Y_g); after the all-gather of the Y_g every rank evaluates
// its slice of  X[k] = sum_g w^(g k) Y_g[k mod n/G].  Inputs and outputs are canonical residues:
// the twiddles are in Montgomery form, so the Montgomery product w * y is the canonical w y.
__global__ void k_ntt_combine(const Fr* __restrict__ parts, const Fr* __restrict__ tw, uint32_t log_n, uint32_t log_g, Fr scale,
                              int has_scale, uint64_t k0, uint64_t count, Fr* __restrict__ out) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const uint64_t n = (uint64_t)1 << log_n, sub = n >> log_g, k = k0 + i, kk = k & (sub - 1);
  Fr acc = parts[kk];
  for (uint64_t g = 1; g < ((uint64_t)1 << log_g); g++) {
    const uint64_t e = (g * k) & (n - 1);
    const Fr w = e < (n >> 1) ? tw[e] : neg(tw[e - (n >> 1)]);
    acc = acc + w * parts[g * sub + kk];
  }
  if (has_scale) acc = scale * acc;
  out[i] = acc;
}
int ntt_combine(zkb_ctx* ctx, const Fr* parts, uint32_t log_n, uint32_t log_g, bool inverse, uint64_t k0, uint64_t count, Fr* out,
                cudaStream_t st) {
  if (!count) return ZKB_OK;
  Fr scale = Fr::one();
  if (inverse) scale = inverse_fermat(fr_from_u64((uint64_t)1 << log_g));  // zkb_ntt_fr's inverse already divided by n/G
  Fr* tw = nullptr;
  if (log_g) ZKB_TRY(get_twiddles(ctx, log_n, inverse, &tw));
  ZKB_LAUNCH(ctx, k_ntt_combine, cdiv(count, 256), 256, 0, st, parts, tw, log_n, log_g, scale, inverse ? 1 : 0, k0, count, out);
  return ZKB_OK;
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
