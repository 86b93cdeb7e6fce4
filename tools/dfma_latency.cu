// dfma_latency.cu — dependent-issue latency of DFMA / DMUL / DADD and the ILP needed to saturate one SMSP.
#include <cuda_runtime.h>
#include <cstdio>
__global__ void k_lat(double* out, long long* cyc, int iters, double a, double b) {
  double x = threadIdx.x * 1e-9;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) { x = fma(x, a, b); x = fma(x, a, b); x = fma(x, a, b); x = fma(x, a, b); x = fma(x, a, b); x = fma(x, a, b); x = fma(x, a, b); x = fma(x, a, b); }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  if (x == 123.456) out[0] = x;
}
template <int ILP>
__global__ void k_ilp(double* out, long long* cyc, int iters, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int j = 0; j < ILP; ++j) x[j] = threadIdx.x * 1e-9 + j;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int j = 0; j < ILP; ++j) x[j] = fma(x[j], a, b);
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  double s = 0;
#pragma unroll
  for (int j = 0; j < ILP; ++j) s += x[j];
  if (s == 123.456) out[0] = s;
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 64); cudaMalloc(&cyc, 64);
  long long h;
  const int iters = 10000;
  k_lat<<<1, 32>>>(out, cyc, iters, 1.0000001, 1e-9); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("{\"dfma_dependent_latency_cycles\": %.2f", (double)h / (8.0 * iters));
#define RUN(ILP, WARPS) k_ilp<ILP><<<1, 32 * WARPS>>>(out, cyc, iters, 1.0000001, 1e-9); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); \
  printf(", \"cycles_per_dfma_ilp%d_warps%d\": %.3f", ILP, WARPS, (double)h / (4.0 * ILP * iters * ((WARPS + 3) / 4)));
  // warps are spread over the 4 SMSPs: WARPS/4 warps per SMSP; cycles per DFMA warp-instruction per SMSP
  RUN(1, 4) RUN(2, 4) RUN(4, 4) RUN(8, 4) RUN(1, 16) RUN(2, 16) RUN(4, 16) RUN(8, 16) RUN(1, 24) RUN(2, 24)
  printf("}\n");
  return 0;
}
