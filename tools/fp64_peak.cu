// fp64_peak.cu — microbenchmarks for the roofline denominators MEASURED_PEAKS.json does not carry:
//   (i)   register-resident DFMA throughput on all SMs           -> P_FP64 (vector pipe)
//   (ii)  mma.sync.m8n8k4.f64 (DMMA) throughput                   -> P_FP64 (tensor pipe)
//   (iii) read-only streaming bandwidth (sum-reduce of a 8 GiB buffer) for comparison with the copy figure
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu ; prints one JSON line.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <algorithm>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("{\"error\": \"%s at %d\"}\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__global__ void __launch_bounds__(256) k_dfma(double* out, int iters, double a, double b) {
  double x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 1e-9 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = fma(x[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i];
  if (s == 123.456) out[0] = s;
}

__global__ void __launch_bounds__(256) k_dmma(double* out, int iters) {
  double c[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) { c[i][0] = 0; c[i][1] = 0; }
  double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  if (s == 123.456) out[0] = s;
}

__global__ void __launch_bounds__(256) k_read(const double2* __restrict__ p, size_t n2, double* out) {
  double s = 0;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
  for (; i + 3 * st < n2; i += 4 * st) {
    double2 a = __ldcs(p + i), b = __ldcs(p + i + st), c = __ldcs(p + i + 2 * st), d = __ldcs(p + i + 3 * st);
    s += a.x + a.y + b.x + b.y + c.x + c.y + d.x + d.y;
  }
  for (; i < n2; i += st) { double2 a = __ldcs(p + i); s += a.x + a.y; }
  if (s == 123.456) out[0] = s;
}

template <class F> float best_ms(F launch, int reps) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); best = std::min(best, ms);
  }
  return best;
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  double* out; CK(cudaMalloc(&out, 64));
  const int iters = 1 << 14;
  const int blocks = sms * 8;
  float ms_dfma = best_ms([&] { k_dfma<<<blocks, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
  CK(cudaGetLastError());
  double dfma_tf = 2.0 * 16.0 * iters * 256.0 * blocks / (ms_dfma * 1e-3) / 1e12;
  float ms_dmma = best_ms([&] { k_dmma<<<blocks, 256>>>(out, iters); }, 5);
  CK(cudaGetLastError());
  double dmma_tf = 2.0 * 256.0 * 8.0 * iters * 8.0 /*warps per block*/ * blocks / (ms_dmma * 1e-3) / 1e12;
  // sustained DFMA (2 s back to back)
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int reps = std::max(1, (int)(2000.0f / ms_dfma));
  cudaEventRecord(e0);
  for (int r = 0; r < reps; ++r) k_dfma<<<blocks, 256>>>(out, iters, 1.0000001, 1e-9);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms_sus; cudaEventElapsedTime(&ms_sus, e0, e1);
  double dfma_sus_tf = 2.0 * 16.0 * iters * 256.0 * blocks * reps / (ms_sus * 1e-3) / 1e12;
  size_t bytes = (size_t)8 << 30;
  double2* buf; CK(cudaMalloc(&buf, bytes)); CK(cudaMemset(buf, 0, bytes));
  float ms_read = best_ms([&] { k_read<<<sms * 8, 256>>>(buf, bytes / 16, out); }, 10);
  CK(cudaGetLastError());
  double read_gbs = bytes / (ms_read * 1e-3) / 1e9;
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"dfma_tflops_burst\": %.2f, \"dfma_tflops_sustained_2s\": %.2f, \"dmma_m8n8k4_tflops\": %.2f, "
         "\"read_only_gbs\": %.1f, \"clock_khz_max\": %d}\n", prop.name, sms, dfma_tf, dfma_sus_tf, dmma_tf, read_gbs, prop.clockRate);
  return 0;
}
