// so_qr.cuh — per-thread Householder folds and the fixed merge tree of the on-device TSQR of [J | r]
// (what QR.apply, linalg/QR.scala:131-250, computes with 2n+1 passes over the data set).
// Used by the single-input families (so_scalar.cu, OpQr) and by the narrow linear-predictor families (so_wide_qr.cu).
#pragma once
#include <type_traits>
#include <utility>

#include "so_device.cuh"

namespace so {
namespace {

template <int C> __host__ __device__ constexpr int tri_idx(int j, int k) { return j * C - j * (j - 1) / 2 + (k - j); }

// 1/x and sqrt(x) for normal, positive x: MUFU seed + Newton steps, branch-free (the IEEE sequences the compiler emits
// for `/` and sqrt() carry slow-path calls; at 2 per column and 8 rows per fold they were a third of the kernel's code).
// Both are accurate to ~1 ulp, which is all a Householder reflector needs (its error enters as O(eps) orthogonality).
__device__ __forceinline__ double fast_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  r = fma(r, fma(-x, r, 1.0), r);
  r = fma(r, fma(-x, r, 1.0), r);
  return fma(r, fma(-x, r, 1.0), r);
}
__device__ __forceinline__ double fast_sqrt(double x) {
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  r = r * fma(-0.5 * x * r, r, 1.5);            // Newton on 1/sqrt
  r = r * fma(-0.5 * x * r, r, 1.5);
  r = r * fma(-0.5 * x * r, r, 1.5);
  const double sq = x * r;
  return fma(0.5 * r, fma(-sq, sq, x), sq);     // one correction step on sqrt itself
}

// Columns whose squared norm is outside [1e-280, 1e280] (zero, subnormal, tiny next to nothing, or huge) take this
// path.  A Householder reflection depends only on the DIRECTION of v = (R_jj - alpha; a_:j), so the column is first
// multiplied by an exact power of two that brings its largest entry near 1; norm, alpha, vp and beta are then formed
// from the scaled column with IEEE sqrt and division, out of line, and alpha goes back to the column's own scale.
// Without the rescaling 2 / v.v overflows for a column of size ~1e-160 and below — e.g. the rounding residue a far
// Gaussian tail leaves in a column after the previous reflection (found by
// tests/test_gpu_scalar.py::test_two_exponential_family_at_random_parameters: NaN triangles).
__device__ __noinline__ double householder_scale(double amax, int& e) {
  e = 0;
  if (!(amax > 0.0) || !(amax < 1.0e308)) return 1.0;      // zero column; inf / NaN: left to propagate through the products
  (void)frexp(amax, &e);                                   // amax = m 2^e, m in [0.5, 1)
  e = e < -1000 ? -1000 : (e > 1000 ? 1000 : e);
  return ldexp(1.0, -e);
}
__device__ __noinline__ void householder_slow(double t2, double rjj, int e, double& alpha, double& vp, double& beta) {
  const double nrm = t2 > 0.0 ? sqrt(t2) : 0.0;
  const double den = nrm * (nrm + fabs(rjj));
  beta = den > 0.0 ? 1.0 / den : 0.0;
  const double as = rjj > 0.0 ? -nrm : nrm;
  vp = rjj - as;
  alpha = ldexp(as, e);
}

// compile-time loops: every index of R[] and a[][] below is a constant expression whatever the unroller decides
// (with plain `#pragma unroll` loops the 8-row fold of C >= 5 columns was left partly rolled and a[][] went to
// local memory: 10x slower).
template <int... I, class F> __device__ __forceinline__ void static_for_impl(std::integer_sequence<int, I...>, F&& f) { (f(std::integral_constant<int, I>()), ...); }
template <int N, class F> __device__ __forceinline__ void static_for(F&& f) { static_for_impl(std::make_integer_sequence<int, N>(), f); }

template <int C, int B>
__device__ __forceinline__ void qr_fold(double (&R)[C * (C + 1) / 2], double (&a)[B][C]) {
  static_for<C>([&](auto J) {
    constexpr int j = decltype(J)::value;
    double sigma = 0.0;
    static_for<B>([&](auto I) { constexpr int i = decltype(I)::value; sigma = fma(a[i][j], a[i][j], sigma); });
    const double rjj = R[tri_idx<C>(j, j)];
    const double t2 = fma(rjj, rjj, sigma);
    double alpha, vp, beta;
    if (t2 > 1.0e-280 && t2 < 1.0e280) {                   // the MUFU seeds flush subnormals: normal range only
      const double nrm = fast_sqrt(t2);
      beta = fast_rcp(nrm * (nrm + fabs(rjj)));            // 2 / v.v,  v.v = -2 alpha vp = 2 nrm (nrm + |R_jj|)
      alpha = rjj > 0.0 ? -nrm : nrm;                      // -sign(R_jj) ||(R_jj; a_:j)||: no cancellation in vp
      vp = rjj - alpha;
    } else {                                               // rare: rescaled reflector (see householder_scale)
      double amax = fabs(rjj);
      static_for<B>([&](auto I) { constexpr int i = decltype(I)::value; amax = fmax(amax, fabs(a[i][j])); });
      int e;
      const double sc = householder_scale(amax, e);
      const double rs = rjj * sc;
      double sig = 0.0;
      static_for<B>([&](auto I) { constexpr int i = decltype(I)::value; a[i][j] *= sc; sig = fma(a[i][j], a[i][j], sig); });
      householder_slow(fma(rs, rs, sig), rs, e, alpha, vp, beta);
    }
    R[tri_idx<C>(j, j)] = alpha;
    static_for<C - 1 - j>([&](auto K) {
      constexpr int k = j + 1 + decltype(K)::value;
      double s = vp * R[tri_idx<C>(j, k)];
      static_for<B>([&](auto I) { constexpr int i = decltype(I)::value; s = fma(a[i][j], a[i][k], s); });
      s *= beta;
      R[tri_idx<C>(j, k)] = fma(-s, vp, R[tri_idx<C>(j, k)]);
      static_for<B>([&](auto I) { constexpr int i = decltype(I)::value; a[i][k] = fma(-s, a[i][j], a[i][k]); });
    });
  });
}

// R <- R factor of [R; S], S a packed upper triangle
template <int C>
__device__ __forceinline__ void qr_merge(double (&R)[C * (C + 1) / 2], const double (&S)[C * (C + 1) / 2]) {
  double a[C][C];
#pragma unroll
  for (int i = 0; i < C; ++i)
#pragma unroll
    for (int k = 0; k < C; ++k) a[i][k] = k >= i ? S[tri_idx<C>(i, k)] : 0.0;
  qr_fold<C, C>(R, a);
}

// lanes 0 of every warp -> warp 0 -> (thread 0 of the block): fixed binary tree
template <int C>
__device__ __forceinline__ void qr_block_tree(double (&R)[C * (C + 1) / 2], double* smem) {
  constexpr int Q = C * (C + 1) / 2;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  auto warp_tree = [&]() {
    for (int o = 16; o >= 1; o >>= 1) {
      double S[Q];
#pragma unroll
      for (int q = 0; q < Q; ++q) S[q] = __shfl_down_sync(0xffffffffu, R[q], o);
      qr_merge<C>(R, S);
    }
  };
  warp_tree();
  if (lane == 0)
#pragma unroll
    for (int q = 0; q < Q; ++q) smem[warp * Q + q] = R[q];
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int q = 0; q < Q; ++q) R[q] = lane < nwarps ? smem[lane * Q + q] : 0.0;
    warp_tree();
  }
  __syncthreads();
}

template <int C>
__device__ __forceinline__ void qr_grid_reduce(double (&R)[C * (C + 1) / 2], double* smem, double* __restrict__ partials,
                                               double* __restrict__ out, const double* __restrict__ colscale,
                                               unsigned int* ticket) {
  constexpr int Q = C * (C + 1) / 2;
  qr_block_tree<C>(R, smem);
  if (threadIdx.x == 0)
#pragma unroll
    for (int q = 0; q < Q; ++q) partials[(size_t)blockIdx.x * Q + q] = R[q];
  __threadfence();
  __syncthreads();
  __shared__ unsigned int s_last_qr;
  if (threadIdx.x == 0) s_last_qr = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
  __syncthreads();
  if (s_last_qr) {
    __threadfence();
#pragma unroll
    for (int q = 0; q < Q; ++q) R[q] = 0.0;
    for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) {      // block order within a thread, then the tree
      double S[Q];
#pragma unroll
      for (int q = 0; q < Q; ++q) S[q] = __ldcg(&partials[(size_t)b * Q + q]);
      qr_merge<C>(R, S);
    }
    qr_block_tree<C>(R, smem);
    if (threadIdx.x == 0) {
#pragma unroll
      for (int j = 0; j < C; ++j)
#pragma unroll
        for (int k = 0; k < C; ++k) out[j * C + k] = k >= j ? R[tri_idx<C>(j, k)] * colscale[k] : 0.0;
      *ticket = 0u;
    }
    so_finish(ticket, out, C * C);
  }
}


}  // namespace
}  // namespace so
