// G2 table expansion and fixed-base generation kernels, see msm_impl.cuh.
#include "msm_impl.cuh"
namespace zkb {
template <> int MsmLaunch<Fq2>::expand_table(zkb_ctx* ctx, G2Affine* tab, size_t stride, size_t n, int c, cudaStream_t st) {
  return launch_expand_table<Fq2>(ctx, tab, stride, n, c, st);
}
int check_points_g2(zkb_ctx* ctx, const G2Affine* pts, size_t n, bool subgroup, int* d_bad, cudaStream_t st) {
  return check_points_impl<Fq2>(ctx, pts, n, subgroup, d_bad, st);
}
int fixed_base_g2(zkb_ctx* ctx, G2Affine* out, const Fr* scalars_mont, size_t n, cudaStream_t st) {
  return fixed_base_impl<Fq2>(ctx, out, scalars_mont, n, st);
}
}  // namespace zkb
