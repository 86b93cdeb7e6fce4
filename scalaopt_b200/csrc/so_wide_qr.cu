// so_wide_qr.cu — on-device Householder TSQR of [J | r] for the NARROW linear-predictor families (linear with or
// without intercept, logistic; n + 1 <= 9 columns): SURVEY.md 8f.1 widened from the single-input families.
//
// What it replaces: QR.apply (linalg/QR.scala:131-250) on MSEFunction.jacobianAndResidualsMatrix
// (function/MSEFunction.scala:132-136) — 2n+1 passes over the data set — by ONE pass that never forms J^T J.
// One thread owns whole observations (as in k_gram_small), keeps a packed upper-triangular R of C = n+1 columns in
// registers and folds B rows at a time into it (so_qr.cuh); the per-thread triangles are merged up the fixed tree
// lanes -> warps -> blocks -> (GPUs, in so_api.cu).  Rows are the basis rows (x~ | res) of the K2 kernel:
//   linear   : J = -sign(rho) x~, r = |rho|, rho = y - z   =>  R_A = R_U diag(-1, .., -1, 1)
//   logistic : J = x~, r = o - y (the reference trainer's convention, Neuron.scala:67-73)
// Bound: FP64 (a fold of B rows costs sum_j [B + (C-1-j)(2B+3)] FMAs), not HBM, from n = 4 on.
#include <algorithm>

#include "so_wide.cuh"
#include "so_qr.cuh"

namespace so {
namespace wide {
namespace {

template <int FAM, int NP, int ICPT, int TPB, int MINB>
__global__ void __launch_bounds__(TPB, MINB)
k_qr_small(const double* __restrict__ X, const double* __restrict__ y, int64_t rows, const __grid_constant__ Params P,
           double* __restrict__ partials, double* __restrict__ out, unsigned int* ticket) {
  constexpr int C = NP + 1, Q = C * (C + 1) / 2, F = NP - ICPT, B = C <= 6 ? 8 : 4;
  __shared__ double s_red[(TPB / 32) * Q];
  __shared__ double s_scale[C];
  if (FAM == kLogistic) exp_table_init();
  if (threadIdx.x < C) s_scale[threadIdx.x] = (threadIdx.x < NP && FAM == kLinear) ? -1.0 : 1.0;
  __syncthreads();
  double R[Q];
#pragma unroll
  for (int q = 0; q < Q; ++q) R[q] = 0.0;
  double p[NP];
#pragma unroll
  for (int i = 0; i < NP; ++i) p[i] = P.p[i];
  const int64_t stride = (int64_t)gridDim.x * TPB;
  for (int64_t row0 = (int64_t)blockIdx.x * TPB + threadIdx.x; row0 < rows; row0 += B * stride) {
    double a[B][C];
    static_for<B>([&](auto I) {
      constexpr int i = decltype(I)::value;
      const int64_t row = row0 + i * stride;
      const bool valid = row < rows;
      double xt[NP];
      if (ICPT) xt[0] = 1.0;
      static_for<F>([&](auto J) { constexpr int j = decltype(J)::value; xt[ICPT + j] = valid ? ld_stream(X + row * F + j) : 0.0; });
      const double yv = valid ? ld_stream(y + row) : 0.0;
      double z = 0.0;
      static_for<NP>([&](auto K) { constexpr int k = decltype(K)::value; z = fma(p[k], xt[k], z); });
      double res;
      if (FAM == kLinear) {
        res = yv - z;
      } else {
        const double e = so_exp(-fabs(z));
        const double inv = so_rcp_fast(1.0 + e);
        res = ((z >= 0.0) ? inv : e * inv) - yv;
      }
      static_for<NP>([&](auto K) { constexpr int k = decltype(K)::value; a[i][k] = valid ? xt[k] : 0.0; });
      a[i][NP] = valid ? res : 0.0;
    });
    qr_fold<C, B>(R, a);
  }
  qr_grid_reduce<C>(R, s_red, partials, out, s_scale, ticket);
}

template <int FAM, int NP, int ICPT>
int launch_qr_cfg(const Shard& s, const Params& P) {
  // up to 6 columns: 3 blocks of 128 threads per SM (12 warps) instead of one block of 256 (as for the single-input
  // families, so_scalar.cu); wider triangles need the whole register file of one 256-thread block
  constexpr int C = NP + 1, Q = C * (C + 1) / 2, TPB = C <= 6 ? 128 : 256, MINB = C <= 6 ? 3 : 1;
  auto kern = k_qr_small<FAM, NP, ICPT, TPB, MINB>;
  const int64_t want = std::max<int64_t>(1, (s.rows + TPB - 1) / TPB);
  const int grid = (int)std::min<int64_t>(want, (int64_t)s.sm_count * MINB);
  if ((size_t)grid * Q > s.partials_doubles || (size_t)C * C > s.out_doubles) return SO_ENOMEM;
  kern<<<grid, TPB, 0, s.stream>>>(s.X, s.y, s.rows, P, s.partials, s.out, s.ticket);
  return cudaGetLastError() == cudaSuccess ? 1 : SO_ECUDA;
}

template <int FAM>
int dispatch_qr(const Shard& s, const Params& P) {
#define SO_CASE(NPV) case NPV: return P.icpt ? launch_qr_cfg<FAM, NPV, 1>(s, P) : launch_qr_cfg<FAM, NPV, 0>(s, P);
  switch (P.n) {
    case 1: return P.icpt ? SO_EUNSUPPORTED : launch_qr_cfg<FAM, 1, 0>(s, P);
    SO_CASE(2) SO_CASE(3) SO_CASE(4) SO_CASE(5) SO_CASE(6) SO_CASE(7) SO_CASE(8)
  }
#undef SO_CASE
  return SO_EUNSUPPORTED;
}

}  // namespace

int launch_qr(const Shard& s, int fam, const Params& P) {
  return fam == kLinear ? dispatch_qr<kLinear>(s, P) : dispatch_qr<kLogistic>(s, P);
}

}  // namespace wide
}  // namespace so
