// dfma_operands.cu — does the FP64 pipe issue slower when an operand comes from the constant bank / a uniform register /
// an immediate, or when DFMA/DMUL/DADD are mixed?  16 warps per SM (4 per SMSP), ILP 8.
#include <cuda_runtime.h>
#include <cstdio>
#include <algorithm>
struct K { double c[8]; };
__constant__ K kc;
template <int MODE>
__global__ void __launch_bounds__(128) k(double* out, int iters, double a, double b, const __grid_constant__ K prm) {
  double x[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) x[j] = threadIdx.x * 1e-9 + j;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (MODE == 0) x[j] = fma(x[j], a, b);                       // reg, reg, reg
      if (MODE == 1) x[j] = fma(x[j], kc.c[j], kc.c[(j + 1) & 7]);  // __constant__ operands
      if (MODE == 2) x[j] = fma(x[j], prm.c[j], prm.c[(j + 1) & 7]); // kernel-parameter operands
      if (MODE == 3) x[j] = fma(x[j], x[j], 0.5);                   // immediate
      if (MODE == 4) { if (j & 1) x[j] = x[j] * a; else x[j] = x[j] + b; }   // DMUL / DADD mix
      if (MODE == 5) { if (j % 3 == 0) x[j] = fma(x[j], a, b); else if (j % 3 == 1) x[j] = x[j] * kc.c[j]; else x[j] = x[j] + kc.c[j]; }
    }
  }
  double s = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += x[j];
  if (s == 123.456) out[0] = s;
}
template <int MODE> double run(double* out, int sms, const K& h) {
  const int iters = 1 << 14, blocks = sms * 4;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int r = 0; r < 4; ++r) { cudaEventRecord(e0); k<MODE><<<blocks, 128>>>(out, iters, 1.0000001, 1e-9, h); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); best = std::min(best, ms); }
  return 8.0 * iters * 128.0 * blocks / (best * 1e-3) / 1e12;   // T ops/s
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  K h; for (int j = 0; j < 8; ++j) h.c[j] = 1.0 + 1e-9 * j;
  cudaMemcpyToSymbol(kc, &h, sizeof h);
  double* out; cudaMalloc(&out, 64);
  printf("{\"Tops_reg\": %.2f, \"Tops_constant\": %.2f, \"Tops_param\": %.2f, \"Tops_imm\": %.2f, \"Tops_dmul_dadd\": %.2f, \"Tops_mix_const\": %.2f}\n",
         run<0>(out, p.multiProcessorCount, h), run<1>(out, p.multiProcessorCount, h), run<2>(out, p.multiProcessorCount, h),
         run<3>(out, p.multiProcessorCount, h), run<4>(out, p.multiProcessorCount, h), run<5>(out, p.multiProcessorCount, h));
  return 0;
}
