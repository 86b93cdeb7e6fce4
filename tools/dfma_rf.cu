// dfma_rf.cu — FP64 issue rate vs number of distinct register-file source operands (operand-collector bandwidth).
#include <cuda_runtime.h>
#include <cstdio>
#include <algorithm>
template <int MODE>
__global__ void __launch_bounds__(128) k(double* out, int iters, double a0) {
  double acc[8], y[8], z[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { acc[j] = threadIdx.x * 1e-9 + j; y[j] = 1.0 + 1e-9 * (threadIdx.x + j); z[j] = 1.0 - 1e-9 * (threadIdx.x + 2 * j); }
  double a = a0;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (MODE == 0) acc[j] = fma(a, y[j], acc[j]);          // a reused, 2 new RF reads
      if (MODE == 1) acc[j] = fma(y[j], z[j], acc[j]);       // 3 distinct RF reads
      if (MODE == 2) acc[j] = fma(y[j], y[j], acc[j]);       // same register twice + acc
      if (MODE == 3) acc[j] = fma(acc[j], a, a);             // 1 new RF read, 2 reused
      if (MODE == 4) acc[j] = fma(y[j], z[(j + 1) & 7], acc[j]);   // 3 distinct, no pattern
    }
  }
  double s = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += acc[j] + y[j] + z[j];
  if (s == 123.456) out[0] = s;
}
template <int MODE> double run(double* out, int sms, int wps) {
  const int iters = 1 << 14, blocks = sms * wps;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int r = 0; r < 4; ++r) { cudaEventRecord(e0); k<MODE><<<blocks, 128>>>(out, iters, 1.0000001); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); best = std::min(best, ms); }
  return 8.0 * iters * 128.0 * blocks / (best * 1e-3) / 1e12;
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  double* out; cudaMalloc(&out, 64);
  for (int wps : {4, 8}) {
    printf("{\"warps_per_sm\": %d, \"axpy_a_reused\": %.2f, \"three_distinct\": %.2f, \"square_plus_acc\": %.2f, \"one_new_two_reused\": %.2f, \"three_distinct_b\": %.2f}\n",
           wps * 4, run<0>(out, p.multiProcessorCount, wps), run<1>(out, p.multiProcessorCount, wps), run<2>(out, p.multiProcessorCount, wps),
           run<3>(out, p.multiProcessorCount, wps), run<4>(out, p.multiProcessorCount, wps));
  }
  return 0;
}
