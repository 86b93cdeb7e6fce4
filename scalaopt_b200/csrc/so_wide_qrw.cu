// so_wide_qrw.cu — on-device Householder TSQR of [J | r] for the WIDE linear-predictor rows: 9 <= n + 1 <= 64 columns
// (linear with or without intercept, logistic).  SURVEY.md 8f.1 beyond the per-thread triangles of so_wide_qr.cu, which
// stop at 9 columns because a thread cannot hold a larger triangle in registers.
//
// What it replaces: QR.apply (linalg/QR.scala:131-250) on MSEFunction.jacobianAndResidualsMatrix
// (function/MSEFunction.scala:132-136) — 2n+1 passes over the data set — by ONE pass that never forms J^T J, so the
// condition number of J is not squared (LevenbergMarquardtSpec.scala:36-58 is n = 11; BASELINE configs 2 and 5 are n = 33 / 32).
//
// A warp owns one upper-triangular R (C x C, shared memory) and folds blocks of B rows into it.  LANES ARE COLUMNS: lane
// k holds column k (and k + 32) of the row block in registers and reads/writes column k of R.  For column j the reflector
// v = (R_jj - alpha; a_:j) lives in lane j % 32; it is broadcast with B shuffles, after which every lane forms
// s_k = beta (v_0 R_jk + a_:j . a_:k) and updates its own column with NO cross-lane reduction.  Rows arrive coalesced (the
// lanes of a warp read consecutive columns of a row); the linear predictor is one warp sum per row.
// Merge: warps of a block pairwise (the rows of the other triangle are folded like data rows), one C x C triangle per
// block to global memory, the last block folds them (all its warps, fixed assignment) — a fixed tree, deterministic.
// Conventions as in so_wide_qr.cu: rows are the basis rows (x~ | res); linear: R_A = R_U diag(-1, .., -1, 1).
// Bound: instruction issue (~30 instructions per column step of B rows), far from HBM — this is the robust route, the
// fast one is K2 (so_gram_ph.cu): n = 33, 1e8 rows: tens of ms here against 5.8 ms there (profiles/r2_tsqr_wide.jsonl).
#include <algorithm>

#include "so_wide.cuh"
#include "so_qr.cuh"

namespace so {
namespace wide {
namespace {

constexpr int kQwThreads = 128, kQwWarps = kQwThreads / 32;

// fold the B rows held column-wise in a[][] into the warp's triangle Rw (C x C, row-major, entries below the diagonal zero)
template <int CPL, int B>
__device__ __forceinline__ void qrw_fold(double* __restrict__ Rw, int C, double (&a)[B][CPL], int lane) {
  for (int j = 0; j < C; ++j) {
    const int src = j & 31, which = j >> 5;
    double cj[B];
#pragma unroll
    for (int i = 0; i < B; ++i) cj[i] = __shfl_sync(0xffffffffu, (CPL > 1 && which) ? a[i][CPL - 1] : a[i][0], src);
    double sigma = 0.0, sigma1 = 0.0;
#pragma unroll
    for (int i = 0; i < B; i += 2) { sigma = fma(cj[i], cj[i], sigma); sigma1 = fma(cj[i + 1], cj[i + 1], sigma1); }
    sigma += sigma1;
    const double rjj = Rw[j * C + j];
    const double t2 = fma(rjj, rjj, sigma);
    double alpha, vp, beta;
    if (t2 > 1.0e-280 && t2 < 1.0e280) {
      const double nrm = fast_sqrt(t2);
      beta = fast_rcp(nrm * (nrm + fabs(rjj)));
      alpha = rjj > 0.0 ? -nrm : nrm;
      vp = rjj - alpha;
    } else {                                                  // zero, tiny or huge column: rescaled reflector (so_qr.cuh)
      double amax = fabs(rjj);
#pragma unroll
      for (int i = 0; i < B; ++i) amax = fmax(amax, fabs(cj[i]));
      int e;
      const double sc = householder_scale(amax, e);
      const double rs = rjj * sc;
      double sig = 0.0;
#pragma unroll
      for (int i = 0; i < B; ++i) { cj[i] *= sc; sig = fma(cj[i], cj[i], sig); }
      householder_slow(fma(rs, rs, sig), rs, e, alpha, vp, beta);
    }
    __syncwarp();                                             // everybody has read R_jj
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
      const int k = lane + 32 * c;
      if (k > j && k < C) {
        const double rjk = Rw[j * C + k];
        double s = vp * rjk, s1 = 0.0;                         // two chains: the step is latency-bound
#pragma unroll
        for (int i = 0; i < B; i += 2) { s = fma(cj[i], a[i][c], s); s1 = fma(cj[i + 1], a[i + 1][c], s1); }
        s = (s + s1) * beta;
        Rw[j * C + k] = fma(-s, vp, rjk);
#pragma unroll
        for (int i = 0; i < B; ++i) a[i][c] = fma(-s, cj[i], a[i][c]);
      } else if (k == j) {
        Rw[j * C + j] = alpha;
      }
    }
  }
  __syncwarp();
}

// Rdst <- R factor of [Rdst; Rsrc]: the rows of Rsrc are folded like data rows, B at a time
template <int CPL, int B>
__device__ __forceinline__ void qrw_merge(double* __restrict__ Rdst, const double* __restrict__ Rsrc, int C, int lane) {
  for (int r0 = 0; r0 < C; r0 += B) {
    double a[B][CPL];
#pragma unroll
    for (int i = 0; i < B; ++i)
#pragma unroll
      for (int c = 0; c < CPL; ++c) {
        const int k = lane + 32 * c, r = r0 + i;
        a[i][c] = (r < C && k < C && k >= r) ? *reinterpret_cast<const volatile double*>(Rsrc + r * C + k) : 0.0;
      }
    qrw_fold<CPL, B>(Rdst, C, a, lane);
  }
}

template <int FAM, int CPL, int B>
__global__ void __launch_bounds__(kQwThreads)
k_qr_cols(const double* __restrict__ X, const double* __restrict__ y, int64_t rows, const __grid_constant__ Params P,
          double* __restrict__ partials, double* __restrict__ out, unsigned int* ticket) {
  extern __shared__ double qw_smem[];                         // kQwWarps triangles of C x C
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = P.n, f = P.f, icpt = P.icpt, C = n + 1;
  double* Rw = qw_smem + (size_t)warp * C * C;
  if (FAM == kLogistic) exp_table_init();
  for (int i = threadIdx.x; i < kQwWarps * C * C; i += kQwThreads) qw_smem[i] = 0.0;
  __syncthreads();
  // this lane's columns: k < n is parameter k (x~_k: the intercept's 1 or an input), k == n is the residual
  double pk[CPL];
  int xcol[CPL];                                              // input column of X (-1: the intercept's 1, -2: residual / none)
#pragma unroll
  for (int c = 0; c < CPL; ++c) {
    const int k = lane + 32 * c;
    pk[c] = k < n ? P.p[k] : 0.0;
    xcol[c] = k < n ? (icpt ? k - 1 : k) : -2;
  }
  const int64_t nblk = (rows + B - 1) / B;
  const int64_t wstride = (int64_t)gridDim.x * kQwWarps;
  for (int64_t blk = (int64_t)blockIdx.x * kQwWarps + warp; blk < nblk; blk += wstride) {
    double a[B][CPL];
#pragma unroll
    for (int i = 0; i < B; ++i) {
      const int64_t row = blk * B + i;
      const bool valid = row < rows;
      double part = 0.0;
#pragma unroll
      for (int c = 0; c < CPL; ++c) {
        double v = 0.0;
        if (valid) {
          if (xcol[c] >= 0) v = ld_stream(X + row * f + xcol[c]);
          else if (xcol[c] == -1) v = 1.0;
        }
        a[i][c] = v;
        part = fma(pk[c], v, part);
      }
      const double z = warp_sum(part);
      const double yv = valid ? __ldg(y + row) : 0.0;
      double res;
      if (FAM == kLinear) {
        res = yv - z;
      } else {
        const double e = so_exp(-fabs(z));
        const double inv = so_rcp_fast(1.0 + e);
        res = ((z >= 0.0) ? inv : e * inv) - yv;
      }
      if (!valid) res = 0.0;
#pragma unroll
      for (int c = 0; c < CPL; ++c)
        if (lane + 32 * c == n) a[i][c] = res;
    }
    qrw_fold<CPL, B>(Rw, C, a, lane);
  }
  __syncthreads();
  // warps of the block, pairwise
  for (int s = kQwWarps / 2; s >= 1; s >>= 1) {
    if (warp < s) qrw_merge<CPL, B>(Rw, qw_smem + (size_t)(warp + s) * C * C, C, lane);
    __syncthreads();
  }
  const int CC = C * C;
  for (int i = threadIdx.x; i < CC; i += kQwThreads) partials[(size_t)blockIdx.x * CC + i] = qw_smem[i];
  __threadfence();
  __syncthreads();
  __shared__ unsigned int s_last_qw;
  if (threadIdx.x == 0) s_last_qw = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
  __syncthreads();
  if (s_last_qw) {
    __threadfence();
    // every warp folds the blocks' triangles b = warp, warp + 4, ... (staged through shared memory), then the tree again
    for (int i = lane; i < CC; i += 32) Rw[i] = 0.0;
    __syncwarp();
    for (unsigned b = warp; b < gridDim.x; b += kQwWarps) qrw_merge<CPL, B>(Rw, partials + (size_t)b * CC, C, lane);   // rows straight from L2
    __syncthreads();
    for (int s = kQwWarps / 2; s >= 1; s >>= 1) {
      if (warp < s) qrw_merge<CPL, B>(Rw, qw_smem + (size_t)(warp + s) * C * C, C, lane);
      __syncthreads();
    }
    for (int i = threadIdx.x; i < CC; i += kQwThreads) {
      const int j = i / C, k = i - j * C;
      const double colscale = (k < n && FAM == kLinear) ? -1.0 : 1.0;
      out[i] = k >= j ? qw_smem[i] * colscale : 0.0;
    }
    if (threadIdx.x == 0) *ticket = 0u;
    so_finish(ticket, out, CC);
  }
}

template <int FAM, int CPL, int B>
int launch_qrw_cfg(const Shard& s, const Params& P) {
  auto kern = k_qr_cols<FAM, CPL, B>;
  const int C = P.n + 1, CC = C * C;
  const size_t smem = sizeof(double) * (size_t)kQwWarps * CC;           // one triangle per warp
  static bool attr[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 64 && !attr[dev]) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) return SO_ECUDA;
    attr[dev] = true;
  }
  int per_sm = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kQwThreads, smem);
  if (per_sm < 1) return SO_EUNSUPPORTED;
  const int64_t nblk = (s.rows + B - 1) / B;
  const int64_t want = std::max<int64_t>(1, (nblk + kQwWarps - 1) / kQwWarps);
  const int grid = (int)std::min<int64_t>(want, (int64_t)per_sm * s.sm_count);
  if ((size_t)grid * CC > s.partials_doubles || (size_t)CC > s.out_doubles) return SO_ENOMEM;
  kern<<<grid, kQwThreads, smem, s.stream>>>(s.X, s.y, s.rows, P, s.partials, s.out, s.ticket);
  return cudaGetLastError() == cudaSuccess ? 1 : SO_ECUDA;
}

}  // namespace

size_t qrw_partials_needed(int n, int sm_count) { return (size_t)sm_count * 16 * (size_t)(n + 1) * (n + 1); }

int launch_qr_wide(const Shard& s, int fam, const Params& P) {
  const int C = P.n + 1;
  if (C > SO_QR_COLS_MAX_N + 1) return SO_EUNSUPPORTED;
  if (C <= 32) return fam == kLinear ? launch_qrw_cfg<kLinear, 1, 16>(s, P) : launch_qrw_cfg<kLogistic, 1, 16>(s, P);
  return fam == kLinear ? launch_qrw_cfg<kLinear, 2, 8>(s, P) : launch_qrw_cfg<kLogistic, 2, 8>(s, P);
}

}  // namespace wide
}  // namespace so
