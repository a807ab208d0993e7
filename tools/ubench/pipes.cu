// Micro-benchmark: issue rates of the integer / FP pipes that a 256-bit modmul can use on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu ; run on the GPU box.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define CHAINS 8
template <int MODE>
__global__ void __launch_bounds__(256) k(uint64_t* out, int iters, uint32_t seed) {
  uint32_t a = threadIdx.x * 2654435761u + seed, b = blockIdx.x * 40503u + 1;
  uint64_t w[CHAINS];
  uint32_t x[CHAINS];
  double d[CHAINS];
  float f[CHAINS];
  const double da = (double)(a | 1), db = 1.0000001;
  const float fa = (float)(a & 0xffff), fb = 1.0001f;
#pragma unroll
  for (int c = 0; c < CHAINS; c++) { w[c] = a + c; x[c] = b + c; d[c] = c; f[c] = c; }
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int c = 0; c < CHAINS; c++) {
      if (MODE == 0 || MODE == 4 || MODE == 5 || MODE == 6 || MODE == 9) {
        uint32_t lo = (uint32_t)w[c];
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[c]) : "r"(lo), "r"(a));
      }
      if (MODE == 1) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(a), "r"(b));
      if (MODE == 8) asm volatile("mad.hi.u32 %0, %0, %1, %0;" : "+r"(x[c]) : "r"(a));
      if (MODE == 2 || MODE == 4 || MODE == 7 || MODE == 9) asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(d[c]) : "d"(db), "d"(da));
      if (MODE == 9) asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(d[c]) : "d"(db), "d"(da));
      if (MODE == 3 || MODE == 5) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[c]) : "f"(fb), "f"(fa));
      if (MODE == 6 || MODE == 7) asm volatile("xor.b32 %0, %0, %1;" : "+r"(x[c]) : "r"(a));
    }
  }
  uint64_t s = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; c++) s += w[c] + x[c] + (uint64_t)d[c] + (uint64_t)f[c];
  if (s == 0x1234567) out[threadIdx.x] = s;
}

template <int MODE>
double run(uint64_t* d_out, int sms, int iters, const char* name, int ops_per_iter_primary) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  int blocks = sms * 8;
  k<MODE><<<blocks, 256>>>(d_out, 10, 1);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 3; r++) {
    cudaEventRecord(e0);
    k<MODE><<<blocks, 256>>>(d_out, iters, r);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  double ops = (double)blocks * 256 * iters * CHAINS * ops_per_iter_primary;
  double per_s = ops / (best * 1e-3);
  printf("%-34s %8.3f ms  %8.2f Gop/s (primary op)  = %6.2f lanes/clk/SM @1.965GHz\n", name, best, per_s / 1e9,
         per_s / sms / 1.965e9);
  return per_s;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
  uint64_t* d; cudaMalloc(&d, 1 << 20);
  int it = 20000, sms = p.multiProcessorCount;
  run<0>(d, sms, it, "IMAD.WIDE.U32", 1);
  run<1>(d, sms, it, "IMAD (lo)", 1);
  run<8>(d, sms, it, "IMAD.HI", 1);
  run<2>(d, sms, it, "DFMA", 1);
  run<3>(d, sms, it, "FFMA", 1);
  run<4>(d, sms, it, "IMAD.WIDE + DFMA (1:1)", 1);
  run<5>(d, sms, it, "IMAD.WIDE + FFMA (1:1)", 1);
  run<6>(d, sms, it, "IMAD.WIDE + LOP (1:1)", 1);
  run<7>(d, sms, it, "DFMA + LOP (1:1)", 1);
  run<9>(d, sms, it, "IMAD.WIDE + 2 DFMA (1:2)", 1);
  return 0;
}
