// dmma_mix.cu — what does it cost to mix DFMA/DADD instructions into a DMMA stream on the same SM sub-partition?
// Per loop trip and warp: 10 independent DMMAs (m8n8k4) + N independent DFMAs, arranged as
//   mode 0: DMMAs only            mode 1: DFMAs only
//   mode 2: 10 DMMAs then N DFMAs (two blocks)
//   mode 3: finely interleaved (1 DMMA, N/10 DFMAs, ...)
//   mode 4: warp-specialised: even warps run the DMMAs of two trips, odd warps the DFMAs of two trips
// 16 warps per SM (4 per sub-partition), all operands in registers, no memory traffic.
#include <cuda_runtime.h>
#include <cstdio>
#include <algorithm>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int MODE, int N>
__global__ void __launch_bounds__(256, 2) k(double* out, int iters, double a, double b) {
  double acc[10][2], x[8], fa[4];
#pragma unroll
  for (int i = 0; i < 10; ++i) { acc[i][0] = threadIdx.x * 1e-9; acc[i][1] = i; }
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-9 + i;
#pragma unroll
  for (int i = 0; i < 4; ++i) fa[i] = a + i * 1e-3 + threadIdx.x * 1e-6;
  const int warp = threadIdx.x >> 5;
  auto mm = [&](int i) { dmma(acc[i][0], acc[i][1], fa[i & 3], fa[(i >> 1) & 3]); };
  auto ff = [&](int j) { asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x[j & 7]) : "d"(a), "d"(b)); };
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 10; ++i) mm(i);
    } else if (MODE == 1) {
#pragma unroll
      for (int j = 0; j < N; ++j) ff(j);
    } else if (MODE == 2) {
#pragma unroll
      for (int i = 0; i < 10; ++i) mm(i);
#pragma unroll
      for (int j = 0; j < N; ++j) ff(j);
    } else if (MODE == 3) {
#pragma unroll
      for (int i = 0; i < 10; ++i) {
        mm(i);
#pragma unroll
        for (int j = 0; j < N / 10; ++j) ff(i * (N / 10) + j);
      }
    } else if (MODE == 5) {          // 10 DMMAs, then a DEPENDENT chain of N DFMAs
#pragma unroll
      for (int i = 0; i < 10; ++i) mm(i);
#pragma unroll
      for (int j = 0; j < N; ++j) ff(0);
    } else if (MODE == 6) {          // dependent chain steps spread between the DMMAs
#pragma unroll
      for (int i = 0; i < 10; ++i) {
        mm(i);
#pragma unroll
        for (int j = 0; j < N / 10; ++j) ff(0);
      }
    } else if (MODE == 7) {          // dependent chain of N DFMAs alone
#pragma unroll
      for (int j = 0; j < N; ++j) ff(0);
    } else {
      if (warp & 1) {
#pragma unroll
        for (int j = 0; j < 2 * N; ++j) ff(j);
      } else {
#pragma unroll
        for (int i = 0; i < 20; ++i) mm(i % 10);
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 10; ++i) s += acc[i][0] + acc[i][1];
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
  if (s == 123.456) out[0] = s;
}
template <int MODE, int N> double run(double* out, int sms) {
  const int iters = 1 << 12;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int r = 0; r < 4; ++r) { cudaEventRecord(e0); k<MODE, N><<<sms * 2, 256>>>(out, iters, 1.0000001, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); best = std::min(best, ms); }
  return best * 1e-3 * 1.965e9 / (iters * 4.0);    // SM cycles per trip per warp slot (4 warps per sub-partition), at 1965 MHz
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  double* out; cudaMalloc(&out, 64);
  const int sms = p.multiProcessorCount;
  printf("{\"note\": \"cycles per trip (10 DMMA + N DFMA) per warp at 4 warps/SMSP, assuming 1965 MHz\", "
         "\"dmma_only\": %.1f, \"dfma20_only\": %.1f, \"blocks_n20\": %.1f, \"interleaved_n20\": %.1f, \"specialised_n20\": %.1f, "
         "\"dfma10_only\": %.1f, \"blocks_n10\": %.1f, \"interleaved_n10\": %.1f, \"specialised_n10\": %.1f, "
         "\"dep10_alone\": %.1f, \"dmma_then_dep10\": %.1f, \"dmma_spread_dep10\": %.1f, \"dep20_alone\": %.1f, \"dmma_then_dep20\": %.1f, \"dmma_spread_dep20\": %.1f}\n",
         run<0, 20>(out, sms), run<1, 20>(out, sms), run<2, 20>(out, sms), run<3, 20>(out, sms), run<4, 20>(out, sms),
         run<1, 10>(out, sms), run<2, 10>(out, sms), run<3, 10>(out, sms), run<4, 10>(out, sms),
         run<7, 10>(out, sms), run<5, 10>(out, sms), run<6, 10>(out, sms), run<7, 20>(out, sms), run<5, 20>(out, sms), run<6, 20>(out, sms));
  return 0;
}
