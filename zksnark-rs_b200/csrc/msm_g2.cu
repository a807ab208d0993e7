// G2 (Fq2): MSM orchestration and the small point kernels.  The heavy Fq2 kernels are
// instantiated in msm_g2_acc.cu / msm_g2_red.cu / msm_g2_tab.cu (parallel nvcc jobs).
#include "msm_impl.cuh"
namespace zkb {
template <> int MsmLaunch<Fq2>::set_inf(zkb_ctx* ctx, G2XYZZ* out, int n, cudaStream_t st) { return launch_set_inf<Fq2>(ctx, out, n, st); }
int msm_g2(zkb_ctx* ctx, const G2Affine* tab, size_t stride, int c, const MsmJob* jobs, int njobs, G2XYZZ* d_out, int slot,
           cudaStream_t st) {
  return msm_run<Fq2>(ctx, tab, stride, c, jobs, njobs, d_out, slot, st, PK_ACC_G2);
}
int expand_table_g2(zkb_ctx* ctx, G2Affine* tab, size_t stride, size_t n, int c, cudaStream_t st) {
  return MsmLaunch<Fq2>::expand_table(ctx, tab, stride, n, c, st);
}
int xyzz_to_affine_g2(zkb_ctx* ctx, G2Affine* out, const G2XYZZ* in, size_t n, cudaStream_t st) {
  return to_affine_impl<Fq2>(ctx, out, in, n, st);
}
int sum_affine_g2(zkb_ctx* ctx, const G2Affine* pts, size_t n, G2XYZZ* d_out, cudaStream_t st) {
  return sum_affine_impl<Fq2>(ctx, pts, n, d_out, st);
}
}  // namespace zkb
