// G2 bucket accumulation kernel (k_accumulate_chunks<Fq2>), see msm_impl.cuh.
#include "msm_impl.cuh"
namespace zkb {
template <> int MsmLaunch<Fq2>::accumulate(zkb_ctx* ctx, const G2Affine* tab, const uint32_t* offs, const uint32_t* sorted,
                                           uint32_t nbk, size_t nacc, G2XYZZ* buckets, G2XYZZ* heads, cudaStream_t st, int pk) {
  return launch_accumulate<Fq2>(ctx, tab, offs, sorted, nbk, nacc, buckets, heads, st, pk);
}
}  // namespace zkb
