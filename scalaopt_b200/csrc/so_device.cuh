// so_device.cuh — device-side building blocks shared by the evaluation kernels (sm_100a).
//
//  * so_exp: fp64 exp with a 128-entry 2^(j/128) table in shared memory and a degree-5 polynomial:
//    10 FP64-pipe instructions instead of ~25 for libdevice exp.  The Gaussian-peak / exponential
//    families are FP64-issue-bound on B200 (64 DFMA/clk/SM vs 16 B/row at HBM speed), so the
//    transcendental is the first thing to shave.  |rel err| < 2^-52 on [-708, 709].
//  * deterministic two-level reduction: fixed thread->row mapping, butterfly shuffles inside a warp,
//    shared memory across warps in warp order, per-block partials in global memory, and the block that
//    draws the last ticket adds the partials in block order.  The result does not depend on scheduling.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace so {

constexpr int kExpTableSize = 128;

__device__ __forceinline__ double exp_slow(double x) { return exp(x); }

// Fill tab[0..128) with 2^(j/128).  Call from every thread of the block, then __syncthreads().
__device__ __forceinline__ void exp_table_init(double* tab) {
  for (int j = threadIdx.x; j < kExpTableSize; j += blockDim.x) tab[j] = exp2((double)j * (1.0 / 128.0));
}

__device__ __forceinline__ double so_exp(double x, const double* __restrict__ tab) {
  const double kInv = 184.66496523378731;            // 128 / ln 2
  const double kShift = 6755399441055744.0;          // 1.5 * 2^52: round-to-nearest-integer trick
  const double kLn2Hi = 0x1.62e42fefa39efp-8;         // ln2/128 rounded to double
  const double kLn2Lo = 0x1.abc9e3b39803fp-63;        // ln2/128 - kLn2Hi
  const int hx = __double2hiint(x) & 0x7fffffff;
  if (hx >= 0x40862000) {                             // |x| >= 708, inf or NaN: rare, take the library path
    return exp_slow(x);
  }
  double kd = fma(x, kInv, kShift);
  const int ki = __double2loint(kd);
  kd -= kShift;
  double r = fma(kd, -kLn2Hi, x);
  r = fma(kd, -kLn2Lo, r);
  // exp(r) - 1 = r (1 + r/2 + r^2/6 + r^3/24 + r^4/120),  |r| <= ln2/256
  double q = fma(r, 8.3333333333333332e-03, 4.1666666666666664e-02);
  q = fma(q, r, 1.6666666666666666e-01);
  q = fma(q, r, 0.5);
  q = fma(q, r, 1.0);
  const double pm1 = q * r;
  const double T = tab[ki & (kExpTableSize - 1)];
  const double res = fma(T, pm1, T);
  const int e = ki >> 7;                              // arithmetic shift: floor(k / 128)
  return __hiloint2double(__double2hiint(res) + e * 1048576, __double2loint(res));
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// streaming 128-bit / 64-bit loads: read once, do not keep in L1
__device__ __forceinline__ double2 ld_stream2(const double2* p) { return __ldcs(p); }
__device__ __forceinline__ double ld_stream(const double* p) { return __ldcs(p); }

// Reduce Q per-thread accumulators over the grid, deterministically.
//   acc[Q]      per-thread partial sums
//   smem        >= (blockDim.x/32) * Q doubles of shared memory
//   partials    gridDim.x * Q doubles of global workspace
//   out[q]      = scale[q] * sum_b partials[b][q], written by the last block to finish
//   ticket      global counter, zero before the launch, reset to zero by the last block
template <int Q>
__device__ __forceinline__ void grid_reduce(double (&acc)[Q], double* smem, double* __restrict__ partials,
                                            double* __restrict__ out, const double* __restrict__ scale,
                                            unsigned int* ticket) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
#pragma unroll
  for (int q = 0; q < Q; ++q) {
    const double s = warp_sum(acc[q]);
    if (lane == 0) smem[warp * Q + q] = s;
  }
  __syncthreads();
  for (int q = threadIdx.x; q < Q; q += blockDim.x) {
    double s = 0.0;
    for (int w = 0; w < nwarps; ++w) s += smem[w * Q + q];
    partials[(size_t)blockIdx.x * Q + q] = s;
  }
  __threadfence();
  __syncthreads();
  __shared__ unsigned int s_last;
  if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
  __syncthreads();
  if (s_last) {
    __threadfence();
    for (int q = threadIdx.x; q < Q; q += blockDim.x) {
      double s = 0.0;
      for (unsigned b = 0; b < gridDim.x; ++b) s += __ldcg(&partials[(size_t)b * Q + q]);
      out[q] = scale ? s * scale[q] : s;
    }
    if (threadIdx.x == 0) *ticket = 0u;
  }
}

// same, run-time Q (accumulators already reduced to one value per (block, q) in `block_vals` shared memory)
__device__ __forceinline__ void grid_reduce_block_vals(const double* block_vals, int Q, double* __restrict__ partials,
                                                       double* __restrict__ out, const double* __restrict__ scale,
                                                       unsigned int* ticket) {
  for (int q = threadIdx.x; q < Q; q += blockDim.x) partials[(size_t)blockIdx.x * Q + q] = block_vals[q];
  __threadfence();
  __syncthreads();
  __shared__ unsigned int s_last2;
  if (threadIdx.x == 0) s_last2 = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
  __syncthreads();
  if (s_last2) {
    __threadfence();
    for (int q = threadIdx.x; q < Q; q += blockDim.x) {
      double s = 0.0;
      for (unsigned b = 0; b < gridDim.x; ++b) s += __ldcg(&partials[(size_t)b * Q + q]);
      out[q] = scale ? s * scale[q] : s;
    }
    if (threadIdx.x == 0) *ticket = 0u;
  }
}

}  // namespace so
