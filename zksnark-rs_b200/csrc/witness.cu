// Witness generation on the device: replaces `weights()` + `evaluate()`
// (/root/reference/src/groth16/circuit/mod.rs:529-656) and the builder's memoised `Circuit::evaluate`
// (circuit/builder/mod.rs:535-580).
//
// The reference walks the program's `(= var (* lhs rhs))` assignments one after the other, each side a
// literal / variable / `+` list of literal-weighted variables (circuit/mod.rs:639-656).  In the QAP data
// model that is, per gate k:   a[out_k] = <u row of gate k, a> * <v row of gate k, a>   with out_k the
// single wire of gate k's w row -- the by-gate CSR rows that k_matvec (prove.cu) reads anyway.  Gates
// whose inputs are all known are independent, so the device evaluates the circuit LEVEL BY LEVEL:
//   plan (host, once per circuit and choice of free wires): level(k) = 1 + max level of the gates that
//     produce k's inputs (free wires and the unity wire: level 0); gates counting-sorted by level; the
//     u / v entries of the gates copied into a level-ordered CSR on the device (one contiguous read per
//     level instead of three dependent gathers per gate);
//   generate (device, per witness): scatter the free values, then
//     - levels wider than 512 gates: one launch each (one thread per gate), chained with PROGRAMMATIC DEPENDENT LAUNCH:
//       a level's grid is launched while its predecessor still runs, loads its gates' structure (offsets, wires,
//       coefficients -- none of which depend on the witness) and only then waits for the predecessor's stores
//       (griddepcontrol.wait), so the launch latency and the structure loads leave the chain (measured: 8.4 -> 6.5 us
//       per level; 2^20 gates in 4 levels 0.245 -> 0.156 ms);
//     - runs of consecutive narrower levels: ONE single-block launch, block barrier between levels: a depth-d stretch
//       costs d barriers instead of d launches.
//     Measured and NOT the default (ZKB_WIT_CLUSTER=1; profiles/r02_witness_v2_cluster_pdl.jsonl): runs of levels up to
//     4096 gates on a thread-block CLUSTER of 8 CTAs with a cluster barrier per level.  A level is a dependent chain (three
//     L2 accesses, five products of one thread), not multiplier throughput, so eight SMs do not shorten it and the cluster
//     barrier + L2-coherent loads cost more than the block barrier: 8.5 vs 7.0 us per 256-gate level, 16 vs 8.4 at 4096.
// Error behaviour mirrors the reference: assigning a wire twice, reading a wire no gate has produced
// ("Under constrained expression"; with ZKB_WITNESS_PROGRAM_ORDER also a wire only a LATER gate produces,
// which is what the sequential walk of circuit/mod.rs:598-621 rejects), a wire nothing assigns, the wrong
// number of values.
#include <string.h>
#include <algorithm>
#include "common.cuh"

struct zkb_witness_plan {
  uint64_t m = 0, n_gates = 0, n_free = 0, n_levels = 0, max_width = 0, nnz = 0;
  uint32_t* d_free = nullptr;   // n_free wire indices
  uint32_t* d_ptr = nullptr;    // 2 * n_gates + 1: [2g] start of the u entries of plan gate g, [2g+1] start of its v entries
  uint32_t* d_wire = nullptr;   // nnz
  zkb::Fr* d_coef = nullptr;    // nnz, Montgomery
  uint32_t* d_out = nullptr;    // n_gates: output wire
  zkb::Fr* d_winv = nullptr;    // n_gates: 1 / (w coefficient), only when some coefficient != 1
  uint32_t* d_lptr = nullptr;   // n_levels + 1: plan-gate range of every level
  struct Seg { uint32_t l0, l1; int kind; };  // 0: chain (one block), 1: cluster, 2: one wide level
  std::vector<Seg> segs;        // launch schedule
  std::vector<uint32_t> h_lptr;
};

namespace zkb {

static const uint32_t WIT_NARROW = 512;    // block size of the chain kernel
static const uint32_t WIT_CHAIN_MAX = 32;  // runs whose levels all have at most this many gates stay in one block
static const uint32_t WIT_CL = 8;          // CTAs per cluster (the portable maximum)
static const uint32_t WIT_CL_THREADS = 256;
static const uint32_t WIT_CL_MAX = 4096;   // levels up to this many gates run inside the cluster kernel

static const uint32_t WIT_UNIT = 0x80000000u;  // wire-index flag: the entry's coefficient is 1

__device__ __forceinline__ Fr ld_fr_l2(const Fr* p) {  // L2-coherent (no L1): the value may come from another SM of the cluster
  Fr a;
  const uint4* q = reinterpret_cast<const uint4*>(p);
  *reinterpret_cast<uint4*>(&a.v[0]) = __ldcg(q);
  *reinterpret_cast<uint4*>(&a.v[4]) = __ldcg(q + 1);
  return a;
}

__device__ __forceinline__ void wit_gate(const uint32_t* __restrict__ ptr, const uint32_t* __restrict__ wire,
                                         const Fr* __restrict__ coef, const uint32_t* __restrict__ out,
                                         const Fr* __restrict__ winv, Fr* a, uint32_t g) {
  const uint32_t pu = ptr[2 * g], pv = ptr[2 * g + 1], pe = ptr[2 * g + 2];
  Fr su = Fr::zero(), sv = Fr::zero();
  for (uint32_t p = pu; p < pe; p++) {
    const uint32_t w = wire[p];
    const Fr x = a[w & ~WIT_UNIT];
    const Fr t = (w & WIT_UNIT) ? x : coef[p] * x;
    if (p < pv) su = su + t; else sv = sv + t;
  }
  Fr r = su * sv;
  if (winv) r = r * winv[g];
  a[out[g]] = r;
}

// one wide level: plan gates [g0, g1), one thread each.  Launched with programmatic stream serialisation: everything
// before griddepcontrol.wait overlaps the previous level's grid; the witness is only touched after it.
__global__ void __launch_bounds__(128) k_wit_level(const uint32_t* __restrict__ ptr, const uint32_t* __restrict__ wire,
                                                   const Fr* __restrict__ coef, const uint32_t* __restrict__ out,
                                                   const Fr* __restrict__ winv, Fr* a, uint32_t g0, uint32_t g1) {
  asm volatile("griddepcontrol.launch_dependents;");  // the next level may start launching (it waits for our completion below)
  const uint32_t g = g0 + blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t pu = 0, pv = 0, pe = 0, o = 0, w0 = 0;
  Fr c0 = Fr::zero();
  if (g < g1) {
    pu = ptr[2 * g]; pv = ptr[2 * g + 1]; pe = ptr[2 * g + 2];
    o = out[g];
    if (pe > pu) { w0 = wire[pu]; c0 = coef[pu]; }  // first entry; the others follow from L2 while the first product runs
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (g >= g1) return;
  Fr su = Fr::zero(), sv = Fr::zero();
  for (uint32_t p = pu; p < pe; p++) {
    const uint32_t w = p == pu ? w0 : wire[p];
    const Fr x = ld_fr_l2(a + (w & ~WIT_UNIT));
    const Fr t = (w & WIT_UNIT) ? x : (p == pu ? c0 : coef[p]) * x;
    if (p < pv) su = su + t; else sv = sv + t;
  }
  Fr r = su * sv;
  if (winv) r = r * winv[g];
  a[o] = r;
}

// a run of levels [l0, l1), each at most WIT_CL_MAX gates, on ONE cluster of WIT_CL CTAs
__global__ void __cluster_dims__(WIT_CL, 1, 1) __launch_bounds__(WIT_CL_THREADS)
    k_wit_cluster(const uint32_t* __restrict__ lptr, const uint32_t* __restrict__ ptr, const uint32_t* __restrict__ wire,
                  const Fr* __restrict__ coef, const uint32_t* __restrict__ out, const Fr* __restrict__ winv, Fr* a, uint32_t l0, uint32_t l1) {
  const uint32_t tid = blockIdx.x * WIT_CL_THREADS + threadIdx.x, nthr = WIT_CL * WIT_CL_THREADS;
  uint32_t lo = lptr[l0];
  for (uint32_t l = l0; l < l1; l++) {
    const uint32_t hi = lptr[l + 1];
    for (uint32_t g = lo + tid; g < hi; g += nthr) {
      const uint32_t pu = ptr[2 * g], pv = ptr[2 * g + 1], pe = ptr[2 * g + 2];
      Fr su = Fr::zero(), sv = Fr::zero();
      for (uint32_t p = pu; p < pe; p++) {
        const uint32_t w = wire[p];
        const Fr x = ld_fr_l2(a + (w & ~WIT_UNIT));
        const Fr t = (w & WIT_UNIT) ? x : coef[p] * x;
        if (p < pv) su = su + t; else sv = sv + t;
      }
      Fr r = su * sv;
      if (winv) r = r * winv[g];
      a[out[g]] = r;
    }
    lo = hi;
    __threadfence();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
}

// a run of narrow levels [l0, l1) in ONE block: the stores of level l are visible to the block after the barrier.
// The structure of the NEXT level's gate (offsets, output wire, the first four wire indices) does not depend on the
// witness: it is loaded before the barrier, so that after it only a[wire] -> products -> store remain on the chain
// (per level of a depth-n chain: 3.0 -> 1.5 us with unit coefficients, profiles/r02_witness_v3.jsonl).
struct WitGate {
  uint32_t pu, pv, pe, o, w[4];
};
__device__ __forceinline__ WitGate wit_fetch(const uint32_t* __restrict__ ptr, const uint32_t* __restrict__ wire,
                                             const uint32_t* __restrict__ out, uint32_t g) {
  WitGate s;
  s.pu = ptr[2 * g]; s.pv = ptr[2 * g + 1]; s.pe = ptr[2 * g + 2];
  s.o = out[g];
#pragma unroll
  for (int i = 0; i < 4; i++) s.w[i] = s.pu + i < s.pe ? wire[s.pu + i] : 0u;
  return s;
}
__global__ void __launch_bounds__(WIT_NARROW) k_wit_run(const uint32_t* __restrict__ lptr, const uint32_t* __restrict__ ptr,
                                                        const uint32_t* __restrict__ wire, const Fr* __restrict__ coef,
                                                        const uint32_t* __restrict__ out, const Fr* __restrict__ winv, Fr* a,
                                                        uint32_t l0, uint32_t l1) {
  uint32_t lo = lptr[l0], hi = lptr[l0 + 1];
  uint32_t g = lo + threadIdx.x;
  bool have = g < hi;
  WitGate s = {};
  if (have) s = wit_fetch(ptr, wire, out, g);
  for (uint32_t l = l0; l < l1; l++) {
    if (have) {
      Fr su = Fr::zero(), sv = Fr::zero();
#pragma unroll
      for (int i = 0; i < 4; i++) {  // the prefetched entries (compile-time indices: s.w stays in registers)
        const uint32_t p = s.pu + i;
        if (p < s.pe) {
          const uint32_t w = s.w[i];
          const Fr x = a[w & ~WIT_UNIT];
          const Fr t = (w & WIT_UNIT) ? x : coef[p] * x;
          if (p < s.pv) su = su + t; else sv = sv + t;
        }
      }
      for (uint32_t p = s.pu + 4; p < s.pe; p++) {
        const uint32_t w = wire[p];
        const Fr x = a[w & ~WIT_UNIT];
        const Fr t = (w & WIT_UNIT) ? x : coef[p] * x;
        if (p < s.pv) su = su + t; else sv = sv + t;
      }
      Fr r = su * sv;
      if (winv) r = r * winv[g];
      a[s.o] = r;
    }
    if (l + 1 < l1) {  // next level's structure, before the barrier
      lo = hi;
      hi = lptr[l + 2];
      g = lo + threadIdx.x;
      have = g < hi;
      if (have) s = wit_fetch(ptr, wire, out, g);
    }
    __syncthreads();
  }
}

// a[] = 0 except the unity wire; the caller's values (canonical) go to the free wires
__global__ void k_wit_init(Fr* __restrict__ a, size_t m, const uint32_t* __restrict__ free_w, const Fr* __restrict__ vals, size_t n_free) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_free) a[free_w[i]] = to_mont(vals[i]);
  if (i == 0) a[0] = Fr::one();
}

// plan construction: entries of plan gate g <- the by-gate rows of QAP gate order[g]
__global__ void k_wit_gather(const uint32_t* __restrict__ order, const uint32_t* __restrict__ ptr, const uint32_t* __restrict__ gptr_u,
                             const uint32_t* __restrict__ wire_u, const Fr* __restrict__ coef_u, const uint32_t* __restrict__ gptr_v,
                             const uint32_t* __restrict__ wire_v, const Fr* __restrict__ coef_v, const uint32_t* __restrict__ gptr_w,
                             const Fr* __restrict__ coef_w, uint32_t* __restrict__ wire, Fr* __restrict__ coef, Fr* __restrict__ winv,
                             int* __restrict__ zero_flag, size_t n_gates) {
  const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_gates) return;
  const uint32_t k = order[g];
  uint32_t o = ptr[2 * g];
  // bit 31 of the wire index marks a coefficient of 1 (variables and unweighted sums: most entries of a parsed program,
  // circuit/mod.rs:300-480): the evaluation then skips the product
  for (uint32_t p = gptr_u[k], e = gptr_u[k + 1]; p < e; p++, o++) {
    const Fr c = coef_u[p];
    wire[o] = wire_u[p] | (c == Fr::one() ? WIT_UNIT : 0u);
    coef[o] = c;
  }
  for (uint32_t p = gptr_v[k], e = gptr_v[k + 1]; p < e; p++, o++) {
    const Fr c = coef_v[p];
    wire[o] = wire_v[p] | (c == Fr::one() ? WIT_UNIT : 0u);
    coef[o] = c;
  }
  const Fr cw = coef_w[gptr_w[k]];
  if (cw.is_zero()) *zero_flag = 1;  // 0 * out = U * V assigns nothing: the reference's evaluate has no such gate; refuse it
  if (winv) winv[g] = inverse(cw);
}

__global__ void k_wit_any_not_one(const Fr* __restrict__ c, size_t n, int* flag) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && !(c[i] == Fr::one())) *flag = 1;
}

}  // namespace zkb

using namespace zkb;

// Levelisation (host only): gp / gw = by-gate offsets and wires of u, v, w.  On success glevel[k] = level of gate k
// (1-based; NONE for gates that assign nothing) and outw[k] = its output wire.
static const uint32_t NONE = 0xffffffffu;
static int wit_levelise(zkb_ctx* ctx, uint64_t n, uint64_t m, const std::vector<uint32_t>* gp, const std::vector<uint32_t>* gw,
                        const uint32_t* free_wires, size_t n_free, bool program_order, std::vector<uint32_t>& glevel,
                        std::vector<uint32_t>& outw, uint64_t* n_gates_out) {
  // wire state: NONE = nothing assigns it (yet); otherwise the level at which its value exists
  std::vector<uint32_t> wlevel(m, NONE), producer(m, NONE);
  wlevel[0] = 0;  // the unity wire (`once(F::one())`, circuit/mod.rs:634)
  for (size_t i = 0; i < n_free; i++) {
    uint32_t w = free_wires[i];
    if (w == 0 || w >= m) return set_err(ctx, ZKB_ERR_ARG, "witness plan: free wire %u out of range [1, %llu)", w, (unsigned long long)m);
    if (wlevel[w] != NONE) return set_err(ctx, ZKB_ERR_ARG, "witness plan: free wire %u listed twice", w);
    wlevel[w] = 0;
  }
  // output wire of every gate: the single entry of its w row.  Gates without a w entry assign nothing (padding gates
  // 0 * 0 = 0 of a re-indexed QAP, or pure constraints) and are skipped.
  outw.assign(n, NONE);
  uint64_t n_gates = 0;
  for (uint64_t k = 0; k < n; k++) {
    uint32_t cnt = gp[2][k + 1] - gp[2][k];
    if (cnt == 0) continue;
    if (cnt != 1)
      return set_err(ctx, ZKB_ERR_UNSUPPORTED, "witness plan: gate %llu has %u wires in its w row; evaluation needs exactly one output "
                     "wire per gate (the form `(= var (* lhs rhs))` produces, circuit/mod.rs:300-340)", (unsigned long long)k, cnt);
    uint32_t w = gw[2][gp[2][k]];
    if (wlevel[w] != NONE || producer[w] != NONE)  // circuit/mod.rs:601-606
      return set_err(ctx, ZKB_ERR_ARG, "witness plan: gate %llu: Attempted to assign to an already assigned variable (wire %u)",
                     (unsigned long long)k, w);
    producer[w] = (uint32_t)k;
    outw[k] = w;
    n_gates++;
  }
  // levels
  glevel.assign(n, NONE);
  auto input_level = [&](uint64_t k, uint32_t w, uint32_t* lv) -> int {  // level at which wire w exists, for gate k
    if (wlevel[w] != NONE) { *lv = wlevel[w]; return ZKB_OK; }
    return set_err(ctx, ZKB_ERR_ARG, "witness plan: gate %llu: Under constrained expression (wire %u has no value%s)",
                   (unsigned long long)k, w, producer[w] != NONE ? " yet: a later gate assigns it" : "");
  };
  if (program_order) {
    for (uint64_t k = 0; k < n; k++) {
      if (outw[k] == NONE) continue;
      uint32_t lv = 0;
      for (int t = 0; t < 2; t++)
        for (uint32_t p = gp[t][k]; p < gp[t][k + 1]; p++) {
          uint32_t l;
          ZKB_TRY(input_level(k, gw[t][p], &l));
          lv = std::max(lv, l);
        }
      glevel[k] = lv + 1;
      wlevel[outw[k]] = lv + 1;
    }
  } else {
    // any topological order (the builder's demand-driven evaluate, builder/mod.rs:556-580): depth-first with an explicit stack
    std::vector<uint8_t> state(n, 0);  // 0 new, 1 on the stack, 2 done
    std::vector<uint64_t> stack;
    for (uint64_t k0 = 0; k0 < n; k0++) {
      if (outw[k0] == NONE || state[k0] == 2) continue;
      stack.push_back(k0);
      while (!stack.empty()) {
        uint64_t k = stack.back();
        if (state[k] == 2) { stack.pop_back(); continue; }
        state[k] = 1;
        uint32_t lv = 0;
        bool ready = true;
        for (int t = 0; t < 2 && ready; t++)
          for (uint32_t p = gp[t][k]; p < gp[t][k + 1]; p++) {
            uint32_t w = gw[t][p];
            if (wlevel[w] != NONE) { lv = std::max(lv, wlevel[w]); continue; }
            uint32_t pk = producer[w];
            if (pk == NONE) { uint32_t l; return input_level(k, w, &l); }
            if (state[pk] == 1)
              return set_err(ctx, ZKB_ERR_ARG, "witness plan: gates %llu and %u depend on each other (Under constrained expression)",
                             (unsigned long long)k, pk);
            stack.push_back(pk);
            ready = false;
            break;
          }
        if (!ready) continue;
        glevel[k] = lv + 1;
        wlevel[outw[k]] = lv + 1;
        state[k] = 2;
        stack.pop_back();
      }
    }
  }
  for (uint64_t w = 1; w < m; w++)
    if (wlevel[w] == NONE)  // `.expect("Every variable should have an assignment")`, circuit/mod.rs:630
      return set_err(ctx, ZKB_ERR_ARG, "witness plan: Every variable should have an assignment (wire %llu has none)", (unsigned long long)w);
  *n_gates_out = n_gates;
  return ZKB_OK;
}

extern "C" {

void zkb_witness_plan_free(zkb_ctx* ctx, zkb_witness_plan* p) {
  if (!p) return;
  if (ctx) cudaSetDevice(ctx->device);
  cudaFree(p->d_free); cudaFree(p->d_ptr); cudaFree(p->d_wire); cudaFree(p->d_coef);
  cudaFree(p->d_out); cudaFree(p->d_winv); cudaFree(p->d_lptr);
  delete p;
}

int zkb_witness_plan_create(zkb_ctx* ctx, const zkb_qap* q, const uint32_t* free_wires, size_t n_free, int flags,
                            zkb_witness_plan** out) {
  if (!ctx || !q || !out || (n_free && !free_wires)) return set_err(ctx, ZKB_ERR_ARG, "zkb_witness_plan_create: NULL argument");
  *out = nullptr;
  const uint64_t n = q->n, m = q->m;
  const std::vector<uint32_t>*gp = q->h_gptr, *gw = q->h_wire;
  if (gp[0].size() != n + 1 || gp[1].size() != n + 1 || gp[2].size() != n + 1)
    return set_err(ctx, ZKB_ERR_ARG, "zkb_witness_plan_create: the QAP carries no host copy of its by-gate rows");
  std::vector<uint32_t> glevel, outw;
  uint64_t n_gates = 0;
  ZKB_TRY(wit_levelise(ctx, n, m, gp, gw, free_wires, n_free, (flags & ZKB_WITNESS_PROGRAM_ORDER) != 0, glevel, outw, &n_gates));
  // counting sort of the gates by level
  uint32_t n_levels = 0;
  for (uint64_t k = 0; k < n; k++)
    if (glevel[k] != NONE) n_levels = std::max(n_levels, glevel[k]);
  std::vector<uint32_t> lptr(n_levels + 1, 0);
  for (uint64_t k = 0; k < n; k++)
    if (glevel[k] != NONE) lptr[glevel[k]]++;  // level l (1-based) -> slot l; prefix sums below shift it to [l-1]
  for (uint32_t l = 1; l <= n_levels; l++) lptr[l] += lptr[l - 1];
  // now lptr[l] = number of gates with level <= l, i.e. END of level l (1-based) = start of plan level index l
  std::vector<uint32_t> order(n_gates ? n_gates : 1), cur(lptr.begin(), lptr.end());
  for (uint64_t k = 0; k < n; k++)
    if (glevel[k] != NONE) order[cur[glevel[k] - 1]++] = (uint32_t)k;
  std::vector<uint32_t> ptr(2 * n_gates + 1), outs(n_gates ? n_gates : 1);
  uint64_t nnz = 0;
  for (uint64_t g = 0; g < n_gates; g++) {
    uint32_t k = order[g];
    ptr[2 * g] = (uint32_t)nnz;
    nnz += gp[0][k + 1] - gp[0][k];
    ptr[2 * g + 1] = (uint32_t)nnz;
    nnz += gp[1][k + 1] - gp[1][k];
    outs[g] = outw[k];
  }
  ptr[2 * n_gates] = (uint32_t)nnz;
  if (nnz >= ((uint64_t)1 << 32)) return set_err(ctx, ZKB_ERR_ARG, "witness plan: too many entries");

  ZKB_CUDA(ctx, cudaSetDevice(ctx->device));
  zkb_witness_plan* p = new zkb_witness_plan();
  p->m = m; p->n_gates = n_gates; p->n_free = n_free; p->n_levels = n_levels; p->nnz = nnz;
  p->h_lptr = lptr;
  for (uint32_t l = 0; l < n_levels; l++) p->max_width = std::max<uint64_t>(p->max_width, lptr[l + 1] - lptr[l]);
  // launch schedule: a wide level is its own launch; a run of narrower levels (up to 512 gates each) shares one
  // single-block launch.  ZKB_WIT_CLUSTER=1 (developer switch, A/B): runs of levels up to 4096 gates on the cluster kernel
  // unless every level of the run is at most WIT_CHAIN_MAX gates.
  static const bool use_cluster = getenv("ZKB_WIT_CLUSTER") && atoi(getenv("ZKB_WIT_CLUSTER")) != 0;
  const uint32_t run_max = use_cluster ? WIT_CL_MAX : WIT_NARROW;
  for (uint32_t l = 0; l < n_levels;) {
    if (lptr[l + 1] - lptr[l] > run_max) { p->segs.push_back({l, l + 1, 2}); l++; continue; }
    uint32_t e = l, widest = 0;
    while (e < n_levels && lptr[e + 1] - lptr[e] <= run_max) { widest = std::max(widest, lptr[e + 1] - lptr[e]); e++; }
    p->segs.push_back({l, e, use_cluster && widest > WIT_CHAIN_MAX ? 1 : 0});
    l = e;
  }
  cudaStream_t st = ctx->stream;
  auto fail = [&](int code) { zkb_witness_plan_free(ctx, p); return code; };
  void* vp;
  int rc = scratch_get(ctx, 12, (n_gates + 1) * 4 + sizeof(int), &vp);
  if (rc != ZKB_OK) return fail(rc);
  uint32_t* d_order = (uint32_t*)vp;
  int* d_flag = (int*)(d_order + n_gates + 1);
  if (cudaMalloc(&p->d_free, (n_free + 1) * 4) || cudaMalloc(&p->d_ptr, (2 * n_gates + 1) * 4) || cudaMalloc(&p->d_wire, (nnz + 1) * 4) ||
      cudaMalloc(&p->d_coef, (nnz + 1) * 32) || cudaMalloc(&p->d_out, (n_gates + 1) * 4) || cudaMalloc(&p->d_lptr, (n_levels + 1) * 4))
    return fail(set_err(ctx, ZKB_ERR_ALLOC, "witness plan: cudaMalloc failed"));
  if (n_free) cudaMemcpyAsync(p->d_free, free_wires, n_free * 4, cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(p->d_ptr, ptr.data(), (2 * n_gates + 1) * 4, cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(p->d_lptr, lptr.data(), (n_levels + 1) * 4, cudaMemcpyHostToDevice, st);
  if (n_gates) {
    cudaMemcpyAsync(p->d_out, outs.data(), n_gates * 4, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d_order, order.data(), n_gates * 4, cudaMemcpyHostToDevice, st);
  }
  cudaMemsetAsync(d_flag, 0, sizeof(int), st);
  int flag = 0;
  if (q->nnz[2]) {
    k_wit_any_not_one<<<cdiv(q->nnz[2], 256), 256, 0, st>>>(q->d_coeff[2], q->nnz[2], d_flag);
    ctx->launches++;
  }
  cudaMemcpyAsync(&flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, st);
  cudaError_t e = cudaStreamSynchronize(st);  // host vectors go out of scope; flag decides the allocation below
  if (e != cudaSuccess) return fail(set_err(ctx, ZKB_ERR_CUDA, "witness plan upload: %s", cudaGetErrorString(e)));
  if (flag && cudaMalloc(&p->d_winv, (n_gates + 1) * 32)) return fail(set_err(ctx, ZKB_ERR_ALLOC, "witness plan: cudaMalloc failed"));
  cudaMemsetAsync(d_flag, 0, sizeof(int), st);
  flag = 0;
  if (n_gates) {
    k_wit_gather<<<cdiv(n_gates, 128), 128, 0, st>>>(d_order, p->d_ptr, q->d_gptr[0], q->d_wire[0], q->d_coeff[0], q->d_gptr[1],
                                                     q->d_wire[1], q->d_coeff[1], q->d_gptr[2], q->d_coeff[2], p->d_wire, p->d_coef,
                                                     p->d_winv, d_flag, (size_t)n_gates);
    ctx->launches++;
  }
  cudaMemcpyAsync(&flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, st);
  e = cudaStreamSynchronize(st);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) return fail(set_err(ctx, ZKB_ERR_CUDA, "witness plan gather: %s", cudaGetErrorString(e)));
  if (flag) return fail(set_err(ctx, ZKB_ERR_DIV_ZERO, "witness plan: a gate's output wire has coefficient 0 in its w row"));
  *out = p;
  return ZKB_OK;
}

// Host only (no device is touched): the levels the planner would assign to the gates of `h` -- the seam the CPU tests
// check the host logic through.  gate_level: n entries (0 = the gate assigns nothing).
int zkb_witness_levels(const zkb_qap_host* h, const uint32_t* free_wires, size_t n_free, int flags, uint32_t* gate_level,
                       uint64_t* n_levels) {
  if (!h || !gate_level || (n_free && !free_wires)) return set_err(nullptr, ZKB_ERR_ARG, "zkb_witness_levels: NULL argument");
  const uint64_t n = h->n, m = h->m;
  if (n == 0 || m == 0 || m >= ((uint64_t)1 << 31)) return set_err(nullptr, ZKB_ERR_ARG, "zkb_witness_levels: bad dimensions");
  std::vector<uint32_t> gp[3], gw[3];
  for (int t = 0; t < 3; t++) {  // by-wire CSR -> by-gate (the transposition zkb_qap_upload does)
    const uint64_t* rp = h->row_ptr[t];
    if (!rp || (rp[m] && !h->gate[t])) return set_err(nullptr, ZKB_ERR_ARG, "zkb_witness_levels: NULL rows");
    gp[t].assign(n + 1, 0);
    gw[t].resize(rp[m]);
    for (uint64_t e = 0; e < rp[m]; e++) {
      if (h->gate[t][e] >= n) return set_err(nullptr, ZKB_ERR_ARG, "gate index out of range");
      gp[t][h->gate[t][e] + 1]++;
    }
    for (uint64_t k = 0; k < n; k++) gp[t][k + 1] += gp[t][k];
    std::vector<uint32_t> cur(gp[t].begin(), gp[t].end() - 1);
    for (uint64_t i = 0; i < m; i++)
      for (uint64_t e = rp[i]; e < rp[i + 1]; e++) gw[t][cur[h->gate[t][e]]++] = (uint32_t)i;
  }
  std::vector<uint32_t> glevel, outw;
  uint64_t n_gates = 0;
  ZKB_TRY(wit_levelise(nullptr, n, m, gp, gw, free_wires, n_free, (flags & ZKB_WITNESS_PROGRAM_ORDER) != 0, glevel, outw, &n_gates));
  uint32_t top = 0;
  for (uint64_t k = 0; k < n; k++) {
    gate_level[k] = glevel[k] == NONE ? 0 : glevel[k];
    top = std::max(top, gate_level[k]);
  }
  if (n_levels) *n_levels = top;
  return ZKB_OK;
}

int zkb_witness_plan_info(const zkb_witness_plan* p, uint64_t* n_gates, uint64_t* n_levels, uint64_t* max_width, uint64_t* n_launches) {
  if (!p) return set_err(nullptr, ZKB_ERR_ARG, "zkb_witness_plan_info: NULL plan");
  if (n_gates) *n_gates = p->n_gates;
  if (n_levels) *n_levels = p->n_levels;
  if (max_width) *max_width = p->max_width;
  if (n_launches) *n_launches = p->segs.size() + 2;  // + k_wit_init and the Montgomery -> canonical conversion
  return ZKB_OK;
}

int zkb_witness_generate(zkb_ctx* ctx, const zkb_witness_plan* p, const uint64_t* values, size_t n_values, int values_on_device,
                         uint64_t* weights_out, int out_on_device) {
  if (!ctx || !p || !weights_out || (p->n_free && !values)) return set_err(ctx, ZKB_ERR_ARG, "zkb_witness_generate: NULL argument");
  if (n_values != p->n_free)  // circuit/mod.rs:553-558
    return set_err(ctx, ZKB_ERR_ARG, "zkb_witness_generate: Wrong number of values supplied (%zu for %llu free wires)", n_values,
                   (unsigned long long)p->n_free);
  ZKB_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const size_t m = p->m;
  Fr* a;
  void* vp;
  if (out_on_device) a = (Fr*)weights_out;
  else { ZKB_TRY(scratch_get(ctx, 11, m * sizeof(Fr), &vp)); a = (Fr*)vp; }
  const Fr* d_vals = (const Fr*)values;
  if (!values_on_device && p->n_free) {
    ZKB_TRY(scratch_get(ctx, 12, p->n_free * sizeof(Fr), &vp));
    ZKB_CUDA(ctx, cudaMemcpyAsync(vp, values, p->n_free * sizeof(Fr), cudaMemcpyHostToDevice, st));
    d_vals = (const Fr*)vp;
  }
  ZKB_CUDA(ctx, cudaMemsetAsync(a, 0, m * sizeof(Fr), st));
  ZKB_LAUNCH(ctx, k_wit_init, cdiv(std::max<size_t>(p->n_free, 1), 256), 256, 0, st, a, m, p->d_free, d_vals, (size_t)p->n_free);
  for (const auto& s : p->segs) {
    if (s.kind == 2) {
      const uint32_t g0 = p->h_lptr[s.l0], g1 = p->h_lptr[s.l1];
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(cdiv(g1 - g0, 128));
      cfg.blockDim = dim3(128);
      cfg.stream = st;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      at[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = at;
      cfg.numAttrs = 1;
      if (ctx->trace) prof_begin(ctx, 0, st, "k_wit_level");
      ZKB_CUDA(ctx, cudaLaunchKernelEx(&cfg, k_wit_level, (const uint32_t*)p->d_ptr, (const uint32_t*)p->d_wire, (const Fr*)p->d_coef,
                                       (const uint32_t*)p->d_out, (const Fr*)p->d_winv, a, g0, g1));
      if (ctx->trace) prof_end(ctx, st);
      ctx->launches++;
    } else if (s.kind == 1) {
      ZKB_LAUNCH(ctx, k_wit_cluster, WIT_CL, WIT_CL_THREADS, 0, st, p->d_lptr, p->d_ptr, p->d_wire, p->d_coef, p->d_out, p->d_winv, a, s.l0, s.l1);
    } else {
      ZKB_LAUNCH(ctx, k_wit_run, 1, WIT_NARROW, 0, st, p->d_lptr, p->d_ptr, p->d_wire, p->d_coef, p->d_out, p->d_winv, a, s.l0, s.l1);
    }
  }
  ZKB_TRY(vec_to_mont(ctx, a, m, false, st));
  if (!out_on_device) ZKB_CUDA(ctx, cudaMemcpyAsync(weights_out, a, m * sizeof(Fr), cudaMemcpyDeviceToHost, st));
  ZKB_CUDA(ctx, cudaStreamSynchronize(st));
  return ZKB_OK;
}

}  // extern "C"
