// G2 (Fq2): MSM orchestration and the small point kernels.  The heavy Fq2 kernels are
// instantiated in msm_g2_acc.cu / msm_g2_red.cu / msm_g2_tab.cu (parallel nvcc jobs).
#include "msm_impl.cuh"
namespace zkb {
template <> int MsmLaunch<Fq2>::set_inf(zkb_ctx* ctx, G2XYZZ* out, int n, cudaStream_t st) { return launch_set_inf<Fq2>(ctx, out, n, st); }
// group dispatch of the phased interface (the typed launches are defined in their own TUs)
int msm_prepare(zkb_ctx* ctx, DevBuf* slots, int slot_base, int group, const void* tab, size_t stride, int c, const MsmJob* jobs,
                int njobs, void* d_out, MsmPlan* plan) {
  if (group == 1) return msm_prepare_t<Fq>(ctx, slots, slot_base, (const G1Affine*)tab, stride, c, jobs, njobs, (G1XYZZ*)d_out, plan);
  return msm_prepare_t<Fq2>(ctx, slots, slot_base, (const G2Affine*)tab, stride, c, jobs, njobs, (G2XYZZ*)d_out, plan);
}
int msm_accumulate(zkb_ctx* ctx, const MsmPlan& P, cudaStream_t st) {
  return P.group == 1 ? msm_accumulate_t<Fq>(ctx, P, st) : msm_accumulate_t<Fq2>(ctx, P, st);
}
int msm_tail(zkb_ctx* ctx, const MsmPlan& P, cudaStream_t st) {
  return P.group == 1 ? msm_tail_t<Fq>(ctx, P, st) : msm_tail_t<Fq2>(ctx, P, st);
}
int msm_g2(zkb_ctx* ctx, const G2Affine* tab, size_t stride, int c, const MsmJob* jobs, int njobs, G2XYZZ* d_out, int slot,
           cudaStream_t st, int win_rank, int win_world) {
  MsmPlan P;
  ZKB_TRY(msm_prepare(ctx, ctx->scratch, slot, 2, tab, stride, c, jobs, njobs, d_out, &P));
  P.win_rank = win_rank; P.win_world = win_world;
  ZKB_TRY(msm_sort(ctx, P, st));
  ZKB_TRY(msm_accumulate(ctx, P, st));
  return msm_tail(ctx, P, st);
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
