// dfma_mix.cu — can non-FP64 instructions issue "under" FP64 instructions (which occupy their pipe 2 cycles)?
// Per loop trip: 8 independent DFMAs plus K independent integer / FP32 / shared-memory instructions.
#include <cuda_runtime.h>
#include <cstdio>
#include <algorithm>
template <int KIND, int K>
__global__ void __launch_bounds__(128) k(double* out, int iters, double a, double b, int ia) {
  __shared__ double tab[256];
  tab[threadIdx.x & 255] = threadIdx.x; tab[(threadIdx.x + 128) & 255] = 1.0;
  __syncthreads();
  double x[8]; int n[8]; float f[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { x[j] = threadIdx.x * 1e-9 + j; n[j] = threadIdx.x + j; f[j] = threadIdx.x + j; }
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      x[j] = fma(x[j], a, b);
      if (j < K) {
        if (KIND == 0) n[j] = (n[j] ^ ia) + (n[j] >> 3);           // ALU: LOP3 / SHF / IADD
        if (KIND == 1) n[j] = n[j] * ia + i;                       // IMAD (FMA pipe)
        if (KIND == 2) f[j] = fmaf(f[j], 1.0001f, 0.5f);           // FFMA
        if (KIND == 3) x[j] += tab[(n[j] + i) & 255];              // LDS + DADD
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += x[j] + n[j] + f[j];
  if (s == 123.456) out[0] = s;
}
template <int KIND, int K> double run(double* out, int sms) {
  const int iters = 1 << 13, blocks = sms * 4;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int r = 0; r < 4; ++r) { cudaEventRecord(e0); k<KIND, K><<<blocks, 128>>>(out, iters, 1.0000001, 1e-9, 12345); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); best = std::min(best, ms); }
  // SM cycles per loop trip per SMSP-warp: time * clock / (iters * warps per SMSP)
  return best * 1e-3 * 1.965e9 / (iters * 4.0);
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  double* out; cudaMalloc(&out, 64);
  const int sms = p.multiProcessorCount;
  printf("{\"note\": \"cycles per trip (8 DFMA + K other) per warp slot at 4 warps/SMSP, assuming 1965 MHz\", \"dfma_only\": %.1f, "
         "\"alu_k4\": %.1f, \"alu_k8\": %.1f, \"imad_k4\": %.1f, \"imad_k8\": %.1f, \"ffma_k8\": %.1f, \"lds_dadd_k4\": %.1f, \"lds_dadd_k8\": %.1f}\n",
         run<0, 0>(out, sms), run<0, 4>(out, sms), run<0, 8>(out, sms), run<1, 4>(out, sms), run<1, 8>(out, sms), run<2, 8>(out, sms),
         run<3, 4>(out, sms), run<3, 8>(out, sms));
  return 0;
}
