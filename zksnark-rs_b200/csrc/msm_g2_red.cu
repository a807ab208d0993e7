// G2 head folding and bucket hierarchy kernels (k_fix_heads<Fq2>, k_bucket_level<Fq2>), see msm_impl.cuh.
#include "msm_impl.cuh"
namespace zkb {
template <> int MsmLaunch<Fq2>::fix_heads(zkb_ctx* ctx, const uint32_t* offs, uint32_t nbk, ChunkPlan ch, G2XYZZ* buckets,
                                          const G2XYZZ* heads, cudaStream_t st) {
  return launch_fix_heads<Fq2>(ctx, offs, nbk, ch, buckets, heads, st);
}
template <> int MsmLaunch<Fq2>::reduce(zkb_ctx* ctx, const G2XYZZ* buckets, uint32_t nb, int njobs, G2XYZZ* lvlS, G2XYZZ* lvlA,
                                       G2XYZZ* d_out, cudaStream_t st, int tail) {
  return launch_reduce<Fq2>(ctx, buckets, nb, njobs, lvlS, lvlA, d_out, st, tail);
}
}  // namespace zkb
