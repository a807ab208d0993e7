// G2 bucket accumulation kernel (k_accumulate_chunks<Fq2>), see msm_impl.cuh.
// Fq2 products as calls to one shared out-of-line copy (operands in registers): the fully inlined G2
// mixed addition is ~100 KB of code and ran instruction-fetch bound (ncu: no_instruction 1.2 per issue,
// fmaheavy pipe 64 %); out of line it reaches 0.89 of the modmul peak (profiles/r01_*).
#define ZKB_FQ2_OOL 1
// -DZKB_FQ2_MSM_LAZY=2 builds rr (q - x3) - y1 ppp with two reductions instead of four (ff.cuh: fq2_msm_lazy, out of line).
// Measured on B200 (profiles/r02_notes.md): 4.5 % fewer multiplier instructions per addition, but the kernel gets SLOWER
// (2^20: 8.30 -> 8.67 ms; the six-product body is a third large out-of-line function and 64 registers of arguments per
// call), so the default keeps two Fq2 products and a subtraction.
#define ZKB_AFF_MIN_BLOCKS 2
#define ZKB_ACC_SM_VARIANT 1  // 1: built, off by default (ZKB_ACC_SM=1 selects it); 2: on by default
#include "msm_impl.cuh"
namespace zkb {
template <> int MsmLaunch<Fq2>::accumulate(zkb_ctx* ctx, const G2Affine* tab, const uint32_t* offs, const uint32_t* sorted,
                                           uint32_t nbk, size_t nacc, ChunkPlan ch, G2XYZZ* buckets, G2XYZZ* heads, cudaStream_t st,
                                           int pk) {
  return launch_accumulate<Fq2>(ctx, tab, offs, sorted, nbk, nacc, ch, buckets, heads, st, pk);
}
template <> int MsmLaunch<Fq2>::accumulate_affine(zkb_ctx* ctx, const MsmPlan& P, cudaStream_t st, int pk) {
  return launch_accumulate_affine<Fq2>(ctx, P, st, pk);
}
}  // namespace zkb
