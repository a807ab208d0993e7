// G2 (Fq2) instantiation of the MSM / point kernels (see msm_impl.cuh).
#include "msm_impl.cuh"
namespace zkb {
int msm_g2(zkb_ctx* ctx, const G2Affine* pts, const Fr* scalars, bool mont, size_t n, int c, G2XYZZ* d_out, int slot,
           cudaStream_t st) {
  return msm_impl<Fq2>(ctx, pts, scalars, mont, n, c, d_out, slot, st);
}
int fixed_base_g2(zkb_ctx* ctx, G2Affine* out, const Fr* scalars_mont, size_t n, cudaStream_t st) {
  return fixed_base_impl<Fq2>(ctx, out, scalars_mont, n, st);
}
int xyzz_to_affine_g2(zkb_ctx* ctx, G2Affine* out, const G2XYZZ* in, size_t n, cudaStream_t st) {
  return to_affine_impl<Fq2>(ctx, out, in, n, st);
}
int sum_affine_g2(zkb_ctx* ctx, const G2Affine* pts, size_t n, G2XYZZ* d_out, cudaStream_t st) {
  return sum_affine_impl<Fq2>(ctx, pts, n, d_out, st);
}
}  // namespace zkb
