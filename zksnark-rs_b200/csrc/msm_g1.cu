// G1 (Fq) instantiation of the MSM / point kernels (see msm_impl.cuh).
#include "msm_impl.cuh"
namespace zkb {
int msm_g1(zkb_ctx* ctx, const G1Affine* pts, const Fr* scalars, bool mont, size_t n, int c, G1XYZZ* d_out, int slot,
           cudaStream_t st) {
  return msm_impl<Fq>(ctx, pts, scalars, mont, n, c, d_out, slot, st);
}
int fixed_base_g1(zkb_ctx* ctx, G1Affine* out, const Fr* scalars_mont, size_t n, cudaStream_t st) {
  return fixed_base_impl<Fq>(ctx, out, scalars_mont, n, st);
}
int xyzz_to_affine_g1(zkb_ctx* ctx, G1Affine* out, const G1XYZZ* in, size_t n, cudaStream_t st) {
  return to_affine_impl<Fq>(ctx, out, in, n, st);
}
int sum_affine_g1(zkb_ctx* ctx, const G1Affine* pts, size_t n, G1XYZZ* d_out, cudaStream_t st) {
  return sum_affine_impl<Fq>(ctx, pts, n, d_out, st);
}
__global__ void k_fq_to_mont(Fq* d, size_t n, int to) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  d[i] = to ? to_mont(d[i]) : from_mont(d[i]);
}
int fq_to_mont(zkb_ctx* ctx, Fq* d, size_t n, bool to, cudaStream_t st) {
  if (!n) return ZKB_OK;
  ZKB_LAUNCH(ctx, k_fq_to_mont, cdiv(n, 256), 256, 0, st, d, n, to ? 1 : 0);
  return ZKB_OK;
}
}  // namespace zkb
