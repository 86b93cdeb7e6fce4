// dmma_lds.cu — what does the 4x4-patch DMMA consumer loop of k_gram_ws sustain when its fragments come from shared
// memory (LDS.64 per fragment, pitch 8T+4) rather than from registers?  Variants: warps per block, random vs constant
// data, barrier per chunk or not.  Prints JSON lines.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <algorithm>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// MODE 0: fragments from LDS each slab; 1: fragments loaded once (registers); BAR: __syncthreads per chunk of 8 slabs
// MODE 2/3/4: as 0 + barrier, plus a small per-chunk phase before the barrier: 2 = 8 dependent DFMAs + shuffle sum,
// 3 = the same length of integer work, 4 = 8 dependent DFMAs only
template <int MODE, bool BAR>
__global__ void __launch_bounds__(512, 1) k(const double* __restrict__ src, double* out, int chunks, int PW, int krows) {
  extern __shared__ double sm[];
  for (int i = threadIdx.x; i < krows * PW; i += blockDim.x) sm[i] = src[i % 4096];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, rsel = lane & 3, csel = lane >> 2;
  const int offa = 32 * (warp % 4) + csel, offb = 32 * ((warp / 4) % 4 + 4) + csel;
  double acc[4][4][2] = {};
  double fa[4], fb[4];
  double extra = 0.0;
  for (int i = 0; i < 4; ++i) { fa[i] = sm[rsel * PW + offa + 8 * i]; fb[i] = sm[rsel * PW + offb + 8 * i]; }
  for (int c = 0; c < chunks; ++c) {
#pragma unroll 2
    for (int slab = 0; slab < krows / 4; ++slab) {
      if (MODE != 1) {
        const double* base = sm + (slab * 4 + rsel) * PW;
#pragma unroll
        for (int i = 0; i < 4; ++i) { fa[i] = base[offa + 8 * i]; fb[i] = base[offb + 8 * i]; }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma(acc[i][j][0], acc[i][j][1], fa[i], fb[j]);
    }
    if (MODE == 2 || MODE == 4) {
      double d = fa[0];
#pragma unroll
      for (int q = 0; q < 8; ++q) d = fma(d, fb[q & 3], sm[(c & 7) * PW + lane + 32 * q]);
      if (MODE == 2) for (int o = 16; o >= 1; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
      extra += d;
    }
    if (MODE == 3) {
      int d = lane + c;
#pragma unroll
      for (int q = 0; q < 8; ++q) d = d * 3 + ((const int*)sm)[(c & 7) * PW + lane + 32 * q];
      for (int o = 16; o >= 1; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
      extra += d;
    }
    if (BAR) __syncthreads();
  }
  if (extra == 1.2345) out[1] = extra;
  double s = 0;
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) s += acc[i][j][0] + acc[i][j][1];
  if (s == 123.456) out[0] = s;
}

template <int MODE, bool BAR>
void run(const char* name, int warps, const double* src, double* out, int sms, bool random) {
  const int PW = 260, krows = 32, chunks = 4000;
  const size_t smem = sizeof(double) * krows * PW;
  cudaFuncSetAttribute(k<MODE, BAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int r = 0; r < 4; ++r) {
    cudaEventRecord(e0);
    k<MODE, BAR><<<sms, warps * 32, smem>>>(src, out, chunks, PW, krows);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); best = std::min(best, ms);
  }
  const double flops = 2.0 * 256.0 * 16.0 * (krows / 4) * chunks * warps * sms;
  printf("{\"variant\": \"%s\", \"warps\": %d, \"random\": %d, \"ms\": %.3f, \"tflops\": %.2f}\n", name, warps, (int)random, best, flops / (best * 1e-3) / 1e12);
}

int main() {
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount;
  double* h = (double*)malloc(4096 * 8);
  double *src, *out; cudaMalloc(&src, 4096 * 8); cudaMalloc(&out, 8);
  for (int random = 0; random < 2; ++random) {
    for (int i = 0; i < 4096; ++i) h[i] = random ? (rand() / (double)RAND_MAX - 0.5) * 3.0 : 1.0 + (i % 7) * 0.125;
    cudaMemcpy(src, h, 4096 * 8, cudaMemcpyHostToDevice);
    for (int warps : {4, 8, 12, 16}) {
      run<1, false>("regs", warps, src, out, sms, random);
      run<0, false>("lds", warps, src, out, sms, random);
      run<0, true>("lds+bar", warps, src, out, sms, random);
      if (random) { run<2, true>("lds+bar+dfma_phase", warps, src, out, sms, random); run<4, true>("lds+bar+dfma_only", warps, src, out, sms, random);
                    run<3, true>("lds+bar+int_phase", warps, src, out, sms, random); }
    }
  }
  return 0;
}
