// so_wide.cu — K1 (fused loss + gradient) and K4 (Hessian-vector) for the linear-predictor families.
// See so_wide.cuh for the thread mapping.  Replaces (paths relative to
// core/src/main/scala/com/github/bruneli/scalaopt/core/ and std-apps/.../learning/nnet/):
//   K1  function/MSEFunction.scala:33-36,138-140 ; FFNeuralNetworkTrainer.scala:59-69,119-125
//   K4  function/ObjectiveFunctionWithGradient.scala:41-47 ; FFNeuralNetworkTrainer.scala:76-80
// Bytes per observation: 8 (f + 1); flops per observation ~ 4 f: HBM-bound.
#include <algorithm>
#include <cstdlib>

#include "so_wide.cuh"

namespace so {
namespace wide {
namespace {

enum { kModeLossGrad = 0, kModeHessVec = 1 };

// ODD: f odd streamed as (f + 1)/2 aligned 16-byte units per row, see so_wide.cuh.  The two row parities hold a column in
// different registers; they meet in the block reduction.
// PF: 0 no prefetch, 1 register prefetch of the next row group, 2 lane-private cp.async ring (so_wide.cuh; V = 2 only)
template <int FAM, int MODE, int L, int C, int V, int PF, bool ODD = false>
__global__ void __launch_bounds__(kThreads, 2)
k_wide(const double* __restrict__ X, const double* __restrict__ y, int64_t rows, const __grid_constant__ Params P,
       int want_grad, double* __restrict__ partials, double* __restrict__ out, unsigned int* ticket) {
  extern __shared__ __align__(128) double smem[];
  constexpr int RW = 32 / L, CV = C * V;
  static_assert(!ODD || V == 2, "ODD streams 16-byte units");
  static_assert(PF != 2 || V == 2, "the ring holds 16-byte units");
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int q = lane & (L - 1), r = lane / L;
  const int f = P.f, icpt = P.icpt;
  const int Q = (MODE == kModeLossGrad) ? P.n + 1 : P.n;
  if (FAM == kLogistic) { exp_table_init(); log_table_init(); __syncthreads(); }
  // parity of every row this lane will ever hold (0: the row starts 16-byte aligned)
  const int odd = ODD ? lane_row_parity<L>(X, warp, r) : 0;
  const double* Xs = X - odd;

  double pj[CV], dj[CV], g[CV];
#pragma unroll
  for (int c = 0; c < C; ++c)
#pragma unroll
    for (int v = 0; v < V; ++v) {
      const int col = (c * L + q) * V + v - odd;
      const bool live = col >= 0 && col < f;
      pj[c * V + v] = live ? P.p[icpt + col] : 0.0;
      dj[c * V + v] = (MODE == kModeHessVec && live) ? P.d[icpt + col] : 0.0;
      g[c * V + v] = 0.0;
    }
  const double p0 = icpt ? P.p[0] : 0.0, d0 = (icpt && MODE == kModeHessVec) ? P.d[0] : 0.0;
  double g0 = 0.0, loss = 0.0;

  const int64_t ngroups = (rows + RW - 1) / RW;
  const int64_t gstr = (int64_t)gridDim.x * kWarps;
  int64_t grp = (int64_t)blockIdx.x * kWarps + warp;

  auto step = [&](const Frag<C, V>& fr, double yv, bool valid) {
    double w;
    if (MODE == kModeLossGrad) {
      double dot = 0.0;
#pragma unroll
      for (int i = 0; i < CV; ++i) dot = fma(pj[i], fr.x[i], dot);
      const double z = row_sum<L>(dot) + p0;
      double l;
      link<FAM>(z, yv, l, w);
      if (!valid) { w = 0.0; l = 0.0; }
      if (q == 0) { loss += l; g0 += w; }
      // want_grad == 0 (loss only): the rank-1 update is done anyway — the pass is HBM-bound and the early return measured
      // 6-10 % SLOWER than the full step (cfg2: 4.39 vs 3.97 ms); the caller ignores the gradient sums
    } else {
      double dv = 0.0;
#pragma unroll
      for (int i = 0; i < CV; ++i) dv = fma(dj[i], fr.x[i], dv);
      dv = row_sum<L>(dv) + d0;
      if (FAM == kLogistic) {
        double dot = 0.0;
#pragma unroll
        for (int i = 0; i < CV; ++i) dot = fma(pj[i], fr.x[i], dot);
        const double z = row_sum<L>(dot) + p0;
        w = curvature<FAM>(z) * dv;
      } else {
        w = dv;
      }
      if (!valid) w = 0.0;
      if (q == 0) g0 += w;
    }
#pragma unroll
    for (int i = 0; i < CV; ++i) g[i] = fma(w, fr.x[i], g[i]);
  };

  if constexpr (PF == 2) {
    if constexpr (V == 2) {
      // [ (kWarps + 1) Q doubles of reduction scratch, rounded up to 16 bytes | ring ]
      const LaneRing<C> rg = lane_ring<C>(smem + ring_offset_doubles((kWarps + 1) * Q));
      int64_t gi = grp;
#pragma unroll
      for (int st = 0; st < kRingDepth; ++st, gi += gstr) {
        if (gi < ngroups) { const int64_t row = gi * RW + r; ring_issue<L, C, (MODE == kModeLossGrad)>(rg, st, Xs, y, row, row < rows, f, q); }
        cpa_commit();
      }
      for (int it = 0; grp < ngroups; grp += gstr, gi += gstr, ++it) {
        const int st = it & (kRingDepth - 1);
        cpa_wait<kRingDepth - 1>();
        Frag<C, V> cur;
        double yv;
        const int64_t row = grp * RW + r;
        ring_read<L, C, ODD, (MODE == kModeLossGrad)>(cur, yv, rg, st, row, rows, f, q, odd);
        step(cur, yv, row < rows);
        if (gi < ngroups) { const int64_t nrow = gi * RW + r; ring_issue<L, C, (MODE == kModeLossGrad)>(rg, st, Xs, y, nrow, nrow < rows, f, q); }
        cpa_commit();
      }
      cpa_wait<0>();
    }
  } else if (PF) {
    // software pipeline: the next row group's loads are in flight while this one is reduced
    Frag<C, V> cur, nxt;
    double ycur = 0.0, ynxt = 0.0;
    bool vcur = false, vnxt = false;
    if (grp < ngroups) {
      const int64_t row = grp * RW + r;
      vcur = row < rows;
      load_row<L, C, V, ODD>(cur, Xs, row, rows, vcur, f, q, odd);
      ycur = vcur ? ld_stream(y + row) : 0.0;
    }
    while (grp < ngroups) {
      const int64_t ng = grp + gstr;
      if (ng < ngroups) {
        const int64_t row = ng * RW + r;
        vnxt = row < rows;
        load_row<L, C, V, ODD>(nxt, Xs, row, rows, vnxt, f, q, odd);
        ynxt = vnxt ? ld_stream(y + row) : 0.0;
      }
      step(cur, ycur, vcur);
      cur = nxt; ycur = ynxt; vcur = vnxt;
      grp = ng;
    }
  } else {
    for (; grp < ngroups; grp += gstr) {
      Frag<C, V> cur;
      const int64_t row = grp * RW + r;
      const bool valid = row < rows;
      load_row<L, C, V, ODD>(cur, Xs, row, rows, valid, f, q, odd);
      const double yv = valid ? ld_stream(y + row) : 0.0;
      step(cur, yv, valid);
    }
  }

  // ---- block reduction in warp order -> per-block partials -> last block adds in block order ----
  double* wsum = smem;                 // [kWarps][Q]
  double* vals = smem + kWarps * Q;    // [Q]
  const int base = (MODE == kModeLossGrad) ? 1 : 0;
  if constexpr (!ODD) {
#pragma unroll
    for (int c = 0; c < C; ++c)
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const double s = col_sum<L>(g[c * V + v]);
        const int col = (c * L + q) * V + v;
        if (r == 0 && col < f) wsum[warp * Q + base + icpt + col] = s;
      }
  } else {
    // rows of equal parity first (the butterfly skips the lowest row bit), then the r = 0 lanes store their columns and
    // the r = 1 lanes (other parity, columns one register further) add theirs; one row per warp: a single parity, stored
    double cs[CV];
#pragma unroll
    for (int i = 0; i < CV; ++i) {
      double s = g[i];
#pragma unroll
      for (int o = 2 * L; o < 32; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      cs[i] = s;
    }
#pragma unroll
    for (int pass = 0; pass < (RW == 1 ? 1 : 2); ++pass) {
#pragma unroll
      for (int c = 0; c < C; ++c)
#pragma unroll
        for (int v = 0; v < V; ++v) {
          const int col = (c * L + q) * V + v - odd;
          if (r == pass && col >= 0 && col < f) {
            double* dst = &wsum[warp * Q + base + icpt + col];
            *dst = pass ? *dst + cs[c * V + v] : cs[c * V + v];
          }
        }
      __syncwarp();
    }
  }
  {
    const double s0 = col_sum<L>(g0), sl = col_sum<L>(loss);
    if (lane == 0) {
      if (MODE == kModeLossGrad) wsum[warp * Q] = sl;
      if (icpt) wsum[warp * Q + base] = s0;
    }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < Q; j += kThreads) {
    double s = 0.0;
    for (int w2 = 0; w2 < kWarps; ++w2) s += wsum[w2 * Q + j];
    vals[j] = s;
  }
  __syncthreads();
  finish_grid(vals, Q, (MODE == kModeLossGrad) ? 1 : 0, P.scale0, partials, out, ticket);
}

#ifndef SO_SMALL_MAX_N
#define SO_SMALL_MAX_N 12
#endif
#ifndef SO_RING_MIN_L
#define SO_RING_MIN_L 4      // lanes per row from which the ring is used (so_wide.cuh); K4 on odd f with 4 lanes per row stays on registers
#endif
#ifndef SO_WIDE_RING
#define SO_WIDE_RING 1      // 0: register prefetch everywhere; 1: ring where it measured faster (so_wide.cuh); 2: every 16-byte mapping
#endif
template <int FAM, int MODE, int L, int C, int V, bool ODD = false>
int launch_lcv(const Shard& s, const Params& P, int want_grad) {
  constexpr int PF = (V == 2 && C <= 8 && (SO_WIDE_RING == 2 || (SO_WIDE_RING == 1 && L >= SO_RING_MIN_L && (MODE == kModeLossGrad || L >= 8 || !ODD)))) ? 2 : (C <= 5 ? 1 : 0);
  auto kern = k_wide<FAM, MODE, L, C, V, PF, ODD>;
  const int Q = (MODE == kModeLossGrad) ? P.n + 1 : P.n;
  size_t smem = sizeof(double) * (size_t)(kWarps + 1) * Q;
  if (PF == 2) {
    smem = sizeof(double) * (size_t)ring_offset_doubles((kWarps + 1) * Q) + ring_bytes(C);
    static bool attr[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 64 && !attr[dev]) {
      if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) return SO_ECUDA;
      attr[dev] = true;
    }
  }
  int per_sm = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, smem);
  if (per_sm < 1) per_sm = 1;
  constexpr int RW = 32 / L;
  const int64_t ngroups = (s.rows + RW - 1) / RW;
  const int64_t want = std::max<int64_t>(1, (ngroups + kWarps - 1) / kWarps);
  const int grid = (int)std::min<int64_t>(want, (int64_t)per_sm * s.sm_count);
  if ((size_t)grid * Q > s.partials_doubles || (size_t)Q > s.out_doubles) return SO_ENOMEM;
  kern<<<grid, kThreads, smem, s.stream>>>(s.X, s.y, s.rows, P, want_grad, s.partials, s.out, s.ticket);
  return cudaGetLastError() == cudaSuccess ? 1 : SO_ECUDA;
}

// ---------------------------------------------------------------------------------------------------
// n <= 12: one thread per observation (K1 and K4).  With 8-96 bytes per row the lanes-per-row mapping above
// spends its time in shuffles (n = 4: 0.69 of the HBM roofline); a thread that keeps its whole row and n + 1
// accumulators in registers streams at the rate of k_gram_small (>= 1.0).  Up to n = 8 since round 1; 9 <= n <= 12
// (two blocks per SM, <= 104 registers) since the end of round 2 (profiles/r2_small17_ab.jsonl): f = 9: K1 0.64 -> 1.01 of
// the HBM copy figure, K4 0.68 -> 1.10, K3 0.59 -> 1.0; f = 10: 0.72 -> 0.97 / 1.07 / 0.98.  Not beyond: at f = 16 the rows
// of a warp are 128 bytes apart and every load instruction touches 32 lines (0.96 -> 0.37), f = 15 gains nothing.
// ---------------------------------------------------------------------------------------------------
template <int FAM, int NP, int ICPT, int MODE>
__global__ void __launch_bounds__(kThreads, NP <= 8 ? 4 : 2)
k_small(const double* __restrict__ X, const double* __restrict__ y, int64_t rows, const __grid_constant__ Params P,
        int want_grad, double* __restrict__ partials, double* __restrict__ out, unsigned int* ticket) {
  constexpr int F = NP - ICPT, Q = (MODE == kModeLossGrad) ? NP + 1 : NP, base = (MODE == kModeLossGrad) ? 1 : 0;
  __shared__ double s_red[kWarps * Q];
  __shared__ double s_scale[Q];
  if (FAM == kLogistic) { exp_table_init(); log_table_init(); }
  for (int q = threadIdx.x; q < Q; q += kThreads) s_scale[q] = (MODE == kModeLossGrad && q == 0) ? P.scale0 : 1.0;
  __syncthreads();
  double acc[Q], p[NP], d[NP];
#pragma unroll
  for (int q = 0; q < Q; ++q) acc[q] = 0.0;
#pragma unroll
  for (int i = 0; i < NP; ++i) { p[i] = P.p[i]; d[i] = P.d[i]; }
  for (int64_t row = (int64_t)blockIdx.x * kThreads + threadIdx.x; row < rows; row += (int64_t)gridDim.x * kThreads) {
    double xt[NP];
    if (ICPT) xt[0] = 1.0;
#pragma unroll
    for (int j = 0; j < F; ++j) xt[ICPT + j] = ld_stream(X + row * F + j);
    double z = 0.0;
#pragma unroll
    for (int j = 0; j < F; ++j) z = fma(p[ICPT + j], xt[ICPT + j], z);
    if (ICPT) z += p[0];
    double w;
    if (MODE == kModeLossGrad) {
      const double yv = ld_stream(y + row);
      double loss;
      link<FAM>(z, yv, loss, w);
      acc[0] += loss;
      if (!want_grad) continue;
    } else {
      double dv = 0.0;
#pragma unroll
      for (int j = 0; j < F; ++j) dv = fma(d[ICPT + j], xt[ICPT + j], dv);
      if (ICPT) dv += d[0];
      w = curvature<FAM>(z) * dv;
    }
#pragma unroll
    for (int i = 0; i < NP; ++i) acc[base + i] = fma(w, xt[i], acc[base + i]);
  }
  grid_reduce<Q>(acc, s_red, partials, out, s_scale, ticket);
}

template <int FAM, int MODE, int NP, int ICPT>
int launch_small_np(const Shard& s, const Params& P, int want_grad) {
  auto kern = k_small<FAM, NP, ICPT, MODE>;
  constexpr int Q = (MODE == kModeLossGrad) ? NP + 1 : NP;
  int per_sm = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, 0);
  per_sm = std::max(1, std::min(per_sm, 8));
  const int64_t want = std::max<int64_t>(1, (s.rows + kThreads - 1) / kThreads);
  const int grid = (int)std::min<int64_t>(want, (int64_t)per_sm * s.sm_count);
  if ((size_t)grid * Q > s.partials_doubles || (size_t)Q > s.out_doubles) return SO_ENOMEM;
  kern<<<grid, kThreads, 0, s.stream>>>(s.X, s.y, s.rows, P, want_grad, s.partials, s.out, s.ticket);
  return cudaGetLastError() == cudaSuccess ? 1 : SO_ECUDA;
}

template <int FAM, int MODE>
int launch_small(const Shard& s, const Params& P, int want_grad) {
#define SO_CASE(NPV) case NPV: return P.icpt ? launch_small_np<FAM, MODE, NPV, 1>(s, P, want_grad) : launch_small_np<FAM, MODE, NPV, 0>(s, P, want_grad);
  switch (P.n) {
    case 1: return P.icpt ? SO_EUNSUPPORTED : launch_small_np<FAM, MODE, 1, 0>(s, P, want_grad);
    SO_CASE(2) SO_CASE(3) SO_CASE(4) SO_CASE(5) SO_CASE(6) SO_CASE(7) SO_CASE(8)
    SO_CASE(9) SO_CASE(10) SO_CASE(11) SO_CASE(12)
  }
#undef SO_CASE
  return SO_EUNSUPPORTED;
}

template <int FAM, int MODE>
int launch_mapped(const Shard& s, const Params& P, int want_grad) {
  if ((P.f & 1) && P.f >= 9) {      // odd f: rows as (f + 1)/2 aligned 16-byte units (k_wide ODD)
    const Mapping mo = choose_mapping(P.f + 1);
#define SO_ODD(LL, CC) if (mo.L == LL && mo.C == CC && mo.V == 2) return launch_lcv<FAM, MODE, LL, CC, 2, true>(s, P, want_grad);
    SO_ODD(2, 4) SO_ODD(4, 4) SO_ODD(4, 5) SO_ODD(8, 4) SO_ODD(8, 5) SO_ODD(16, 4) SO_ODD(16, 5) SO_ODD(32, 4) SO_ODD(32, 5) SO_ODD(32, 8)
#undef SO_ODD
  }
  const Mapping m = choose_mapping(P.f);
#define SO_CASE(LL, CC, VV) if (m.L == LL && m.C == CC && m.V == VV) return launch_lcv<FAM, MODE, LL, CC, VV>(s, P, want_grad);
  // V = 1 (64-bit loads) has no caller left: odd f >= 9 streams aligned 16-byte units (ODD above), n <= 8 is k_small
  SO_CASE(1, 4, 2) SO_CASE(2, 4, 2) SO_CASE(4, 4, 2) SO_CASE(4, 5, 2) SO_CASE(8, 4, 2) SO_CASE(8, 5, 2) SO_CASE(16, 4, 2) SO_CASE(16, 5, 2) SO_CASE(32, 4, 2)
  SO_CASE(32, 5, 2) SO_CASE(32, 8, 2)
#undef SO_CASE
  return SO_EUNSUPPORTED;
}

}  // namespace

int launch_line(const Shard& s, int fam, const Params& P);                      // so_wide_line.cu
// rows of up to this many parameters go to the thread-per-row kernels (k_small, k_small_line)
int small_max_n() {
  static const int v = [] { const char* e = getenv("SCALAOPT_SMALL_MAX_N"); const int x = e ? atoi(e) : SO_SMALL_MAX_N; return x < 8 ? 8 : (x > 12 ? 12 : x); }();
  return v;
}
int launch_gram(const Shard& s, int fam, const Params& P);                      // so_gram.cu
int launch_qr(const Shard& s, int fam, const Params& P);                        // so_wide_qr.cu (n <= 8)
int launch_qr_wide(const Shard& s, int fam, const Params& P);                   // so_wide_qrw.cu (8 < n <= 47)
size_t qrw_partials_needed(int n, int sm_count);
size_t gram_partials_needed(int n, int f, int sm_count);                        // so_gram.cu

}  // namespace wide

size_t wide_partials_needed(Kind kind, int family, int n, int f, int k, int sm_count) {
  (void)family;
  if (kind == kNormalEq) return wide::gram_partials_needed(n, f, sm_count);
  if (kind == kQr && n > SO_QR_WIDE_MAX_N) return wide::qrw_partials_needed(n, sm_count);
  const size_t Q = (size_t)std::max(n + 1, 2 * std::max(k, 1));
  return (size_t)sm_count * 8 * Q;    // at most 8 resident blocks per SM
}

int launch_wide(Kind kind, const Shard& s, int family, const double* p, const double* d, int n,
                const double* alpha, int k, int want_grad) {
  using namespace wide;
  if (family != SO_FAMILY_LINEAR && family != SO_FAMILY_LINEAR_NOINT && family != SO_FAMILY_LOGISTIC) return SO_EUNSUPPORTED;
  if (s.f < 1 || s.f > SO_MAX_FEATURES) return SO_EUNSUPPORTED;
  Params P;
  P.n = n; P.f = s.f; P.icpt = (family == SO_FAMILY_LINEAR_NOINT) ? 0 : 1; P.k = k;
  if (n != s.f + P.icpt) return SO_EINVAL;
  for (int i = 0; i < n; ++i) { P.p[i] = p[i]; P.d[i] = d ? d[i] : 0.0; }
  for (int i = 0; i < k && i < SO_MAX_ALPHAS; ++i) P.alpha[i] = alpha[i];
  const int fam = (family == SO_FAMILY_LOGISTIC) ? kLogistic : kLinear;
  P.scale0 = (fam == kLinear) ? 0.5 : 1.0;
  switch (kind) {
    case kLossGrad:
      if (n <= small_max_n()) return fam == kLinear ? launch_small<kLinear, kModeLossGrad>(s, P, want_grad)
                                        : launch_small<kLogistic, kModeLossGrad>(s, P, want_grad);
      return fam == kLinear ? launch_mapped<kLinear, kModeLossGrad>(s, P, want_grad)
                            : launch_mapped<kLogistic, kModeLossGrad>(s, P, want_grad);
    case kHessVec:
      P.scale0 = 1.0;
      if (n <= small_max_n()) return fam == kLinear ? launch_small<kLinear, kModeHessVec>(s, P, 1)
                                        : launch_small<kLogistic, kModeHessVec>(s, P, 1);
      return fam == kLinear ? launch_mapped<kLinear, kModeHessVec>(s, P, 1)
                            : launch_mapped<kLogistic, kModeHessVec>(s, P, 1);
    case kLine:
      return launch_line(s, fam, P);
    case kNormalEq:
      return launch_gram(s, fam, P);
    case kQr:
      return n <= SO_QR_WIDE_MAX_N ? launch_qr(s, fam, P) : (n <= SO_QR_COLS_MAX_N ? launch_qr_wide(s, fam, P) : SO_EUNSUPPORTED);
    case kNll: break;      // single-input pdfs only (launch_nll)
  }
  return SO_EINVAL;
}

}  // namespace so
