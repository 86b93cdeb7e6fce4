// so_wide_line.cu — K3: k step sizes scored in ONE pass over the observations, linear-predictor families.
// phi(alpha) = loss(p + alpha d), phi'(alpha) = grad(p + alpha d) . d  (variable/Point.scala:54-57 as asked for
// by linesearch/StrongWolfe.scala:96-118,214-217).  For these families the linear predictor is affine in alpha:
// z(alpha) = x~.p + alpha x~.d, so a row costs two dot products (4 f flops) whatever k is, and the k evaluations of
// the scalar link are spread over the L lanes that share the row (lane q takes alpha_q, alpha_{q+L}, ...).
#include <algorithm>

#include "so_wide.cuh"

namespace so {
namespace wide {
int small_max_n();          // so_wide.cu
namespace {

template <int FAM, int L, int C, int V, int KPL, int PF, bool ODD = false>      // PF, ODD: as in k_wide (so_wide.cu, so_wide.cuh)
__global__ void __launch_bounds__(kThreads, 2)
k_wide_line(const double* __restrict__ X, const double* __restrict__ y, int64_t rows, const __grid_constant__ Params P,
            double* __restrict__ partials, double* __restrict__ out, unsigned int* ticket) {
  extern __shared__ __align__(128) double smem[];
  constexpr int RW = 32 / L, CV = C * V;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int q = lane & (L - 1), r = lane / L;
  const int f = P.f, icpt = P.icpt, k = P.k;
  const int Q = 2 * k;
  if (FAM == kLogistic) { exp_table_init(); log_table_init(); __syncthreads(); }
  static_assert(!ODD || V == 2, "ODD streams 16-byte units");
  const int odd = ODD ? lane_row_parity<L>(X, warp, r) : 0;
  const double* Xs = X - odd;

  double pj[CV], dj[CV];
#pragma unroll
  for (int c = 0; c < C; ++c)
#pragma unroll
    for (int v = 0; v < V; ++v) {
      const int col = (c * L + q) * V + v - odd;
      const bool live = col >= 0 && col < f;
      pj[c * V + v] = live ? P.p[icpt + col] : 0.0;
      dj[c * V + v] = live ? P.d[icpt + col] : 0.0;
    }
  const double p0 = icpt ? P.p[0] : 0.0, d0 = icpt ? P.d[0] : 0.0;
  double al[KPL], phi[KPL], dphi[KPL];
#pragma unroll
  for (int i = 0; i < KPL; ++i) {
    const int a = q + i * L;
    al[i] = P.alpha[a < k ? a : k - 1];
    phi[i] = 0.0; dphi[i] = 0.0;
  }

  const int64_t ngroups = (rows + RW - 1) / RW;
  const int64_t gstr = (int64_t)gridDim.x * kWarps;
  int64_t grp = (int64_t)blockIdx.x * kWarps + warp;

  auto step = [&](const Frag<C, V>& fr, double yv, bool valid) {
    double u = 0.0, dv = 0.0;
#pragma unroll
    for (int i = 0; i < CV; ++i) { u = fma(pj[i], fr.x[i], u); dv = fma(dj[i], fr.x[i], dv); }
    u = row_sum<L>(u) + p0;
    dv = row_sum<L>(dv) + d0;
    if (!valid) return;              // lanes of one row agree on `valid`; no shuffles below
#pragma unroll
    for (int i = 0; i < KPL; ++i) {
      double l, w;
      link<FAM>(fma(al[i], dv, u), yv, l, w);
      phi[i] += l;
      dphi[i] = fma(w, dv, dphi[i]);
    }
  };

  if constexpr (PF == 2) {
    if constexpr (V == 2) {
      const LaneRing<C> rg = lane_ring<C>(smem + ring_offset_doubles((kWarps + 1) * Q));
      int64_t gi = grp;
#pragma unroll
      for (int st = 0; st < kRingDepth; ++st, gi += gstr) {
        if (gi < ngroups) { const int64_t row = gi * RW + r; ring_issue<L, C, true>(rg, st, Xs, y, row, row < rows, f, q); }
        cpa_commit();
      }
      for (int it = 0; grp < ngroups; grp += gstr, gi += gstr, ++it) {
        const int st = it & (kRingDepth - 1);
        cpa_wait<kRingDepth - 1>();
        Frag<C, V> cur;
        double yv;
        const int64_t row = grp * RW + r;
        ring_read<L, C, ODD, true>(cur, yv, rg, st, row, rows, f, q, odd);
        step(cur, yv, row < rows);
        if (gi < ngroups) { const int64_t nrow = gi * RW + r; ring_issue<L, C, true>(rg, st, Xs, y, nrow, nrow < rows, f, q); }
        cpa_commit();
      }
      cpa_wait<0>();
    }
  } else if (PF) {
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

  double* wsum = smem;                 // [kWarps][Q]
  double* vals = smem + kWarps * Q;    // [Q]
#pragma unroll
  for (int i = 0; i < KPL; ++i) {
    const double sp = col_sum<L>(phi[i]), sd = col_sum<L>(dphi[i]);
    const int a = q + i * L;
    if (r == 0 && a < k) { wsum[warp * Q + a] = sp; wsum[warp * Q + k + a] = sd; }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < Q; j += kThreads) {
    double s = 0.0;
    for (int w2 = 0; w2 < kWarps; ++w2) s += wsum[w2 * Q + j];
    vals[j] = s;
  }
  __syncthreads();
  finish_grid(vals, Q, k, P.scale0, partials, out, ticket);
}

#ifndef SO_RING_MIN_L
#define SO_RING_MIN_L 4      // logistic with 4 lanes per row stays on registers: k = 4 / 8 step sizes 4.46 / 6.10 -> 4.71 / 6.82 ms with the ring
#endif
#ifndef SO_RING_K3_EVEN
#define SO_RING_K3_EVEN 1
#endif
#ifndef SO_WIDE_RING
#define SO_WIDE_RING 1      // as in so_wide.cu
#endif
template <int FAM, int L, int C, int V, int KPL, bool ODD>
int launch_one(const Shard& s, const Params& P) {
  constexpr int PF = (V == 2 && C <= 8 && (SO_WIDE_RING == 2 || (SO_WIDE_RING == 1 && L >= SO_RING_MIN_L && (ODD || SO_RING_K3_EVEN) && (L >= 8 || FAM == kLinear)))) ? 2 : (C <= 5 ? 1 : 0);
  auto kern = k_wide_line<FAM, L, C, V, KPL, PF, ODD>;
  const int Q = 2 * P.k;
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
  kern<<<grid, kThreads, smem, s.stream>>>(s.X, s.y, s.rows, P, s.partials, s.out, s.ticket);
  return cudaGetLastError() == cudaSuccess ? 1 : SO_ECUDA;
}

template <int FAM, int L, int C, int V, bool ODD = false>
int launch_kpl(const Shard& s, const Params& P) {
  const int kpl = (P.k + L - 1) / L;      // step sizes per lane
  if (kpl <= 1) return launch_one<FAM, L, C, V, 1, ODD>(s, P);
  if (kpl <= 2) return launch_one<FAM, L, C, V, 2, ODD>(s, P);
  if (kpl <= 4) return launch_one<FAM, L, C, V, 4, ODD>(s, P);
  if (kpl <= 8) return launch_one<FAM, L, C, V, 8, ODD>(s, P);
  return SO_EINVAL;                        // the API layer sends at most 8 step sizes per pass
}

// n <= 8: one thread per observation, KB step sizes in registers (see k_small in so_wide.cu)
template <int FAM, int NP, int ICPT, int KB>
__global__ void __launch_bounds__(kThreads, NP <= 8 ? 4 : 2)
k_small_line(const double* __restrict__ X, const double* __restrict__ y, int64_t rows, const __grid_constant__ Params P,
             double* __restrict__ partials, double* __restrict__ out, unsigned int* ticket) {
  constexpr int F = NP - ICPT, Q = 2 * KB;
  __shared__ double s_red[kWarps * Q];
  __shared__ double s_scale[Q];
  __shared__ signed char s_src[Q];
  const int k = P.k;
  if (FAM == kLogistic) { exp_table_init(); log_table_init(); }
  for (int q = threadIdx.x; q < Q; q += kThreads) {          // out: phi[0..k), dphi[0..k) from accumulators [0..KB), [KB..2KB)
    s_scale[q] = (q < k) ? P.scale0 : 1.0;
    s_src[q] = (signed char)((q < k) ? q : KB + (q - k));
  }
  __syncthreads();
  double acc[Q], p[NP], d[NP], al[KB];
#pragma unroll
  for (int q = 0; q < Q; ++q) acc[q] = 0.0;
#pragma unroll
  for (int i = 0; i < NP; ++i) { p[i] = P.p[i]; d[i] = P.d[i]; }
#pragma unroll
  for (int a = 0; a < KB; ++a) al[a] = P.alpha[a < k ? a : k - 1];
  for (int64_t row = (int64_t)blockIdx.x * kThreads + threadIdx.x; row < rows; row += (int64_t)gridDim.x * kThreads) {
    double u = 0.0, dv = 0.0;
#pragma unroll
    for (int j = 0; j < F; ++j) {
      const double x = ld_stream(X + row * F + j);
      u = fma(p[ICPT + j], x, u);
      dv = fma(d[ICPT + j], x, dv);
    }
    if (ICPT) { u += p[0]; dv += d[0]; }
    const double yv = ld_stream(y + row);
#pragma unroll
    for (int a = 0; a < KB; ++a) {
      double l, w;
      link<FAM>(fma(al[a], dv, u), yv, l, w);
      acc[a] += l;
      acc[KB + a] = fma(w, dv, acc[KB + a]);
    }
  }
  grid_reduce<Q>(acc, s_red, partials, out, s_scale, ticket, s_src, 2 * k);
}

template <int FAM, int NP, int ICPT, int KB>
int launch_small_line_kb(const Shard& s, const Params& P) {
  auto kern = k_small_line<FAM, NP, ICPT, KB>;
  constexpr int Q = 2 * KB;
  int per_sm = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, 0);
  per_sm = std::max(1, std::min(per_sm, 8));
  const int64_t want = std::max<int64_t>(1, (s.rows + kThreads - 1) / kThreads);
  const int grid = (int)std::min<int64_t>(want, (int64_t)per_sm * s.sm_count);
  if ((size_t)grid * Q > s.partials_doubles || (size_t)Q > s.out_doubles) return SO_ENOMEM;
  kern<<<grid, kThreads, 0, s.stream>>>(s.X, s.y, s.rows, P, s.partials, s.out, s.ticket);
  return cudaGetLastError() == cudaSuccess ? 1 : SO_ECUDA;
}

template <int FAM, int NP, int ICPT>
int launch_small_line_np(const Shard& s, const Params& P) {
  if constexpr (NP > 8) {      // fewer instantiations for the wider rows
    if (P.k <= 2) return launch_small_line_kb<FAM, NP, ICPT, 2>(s, P);
    if (P.k <= 4) return launch_small_line_kb<FAM, NP, ICPT, 4>(s, P);
    return launch_small_line_kb<FAM, NP, ICPT, 8>(s, P);
  }
  if (P.k <= 1) return launch_small_line_kb<FAM, NP, ICPT, 1>(s, P);
  if (P.k <= 2) return launch_small_line_kb<FAM, NP, ICPT, 2>(s, P);
  if (P.k <= 4) return launch_small_line_kb<FAM, NP, ICPT, 4>(s, P);
  return launch_small_line_kb<FAM, NP, ICPT, 8>(s, P);
}

template <int FAM>
int launch_small_line(const Shard& s, const Params& P) {
#define SO_CASE(NPV) case NPV: return P.icpt ? launch_small_line_np<FAM, NPV, 1>(s, P) : launch_small_line_np<FAM, NPV, 0>(s, P);
  switch (P.n) {
    case 1: return P.icpt ? SO_EUNSUPPORTED : launch_small_line_np<FAM, 1, 0>(s, P);
    SO_CASE(2) SO_CASE(3) SO_CASE(4) SO_CASE(5) SO_CASE(6) SO_CASE(7) SO_CASE(8)
    SO_CASE(9) SO_CASE(10) SO_CASE(11) SO_CASE(12)
  }
#undef SO_CASE
  return SO_EUNSUPPORTED;
}

template <int FAM>
int launch_mapped(const Shard& s, const Params& P) {
  if ((P.f & 1) && P.f >= 9) {
    const Mapping mo = choose_mapping(P.f + 1);
#define SO_ODD(LL, CC) if (mo.L == LL && mo.C == CC && mo.V == 2) return launch_kpl<FAM, LL, CC, 2, true>(s, P);
    SO_ODD(2, 4) SO_ODD(4, 4) SO_ODD(4, 5) SO_ODD(8, 4) SO_ODD(8, 5) SO_ODD(16, 4) SO_ODD(16, 5) SO_ODD(32, 4) SO_ODD(32, 5) SO_ODD(32, 8)
#undef SO_ODD
  }
  const Mapping m = choose_mapping(P.f);
#define SO_CASE(LL, CC, VV) if (m.L == LL && m.C == CC && m.V == VV) return launch_kpl<FAM, LL, CC, VV>(s, P);
  // V = 1 (64-bit loads) has no caller left: odd f >= 9 streams aligned 16-byte units (ODD above), n <= 8 is k_small
  SO_CASE(1, 4, 2) SO_CASE(2, 4, 2) SO_CASE(4, 4, 2) SO_CASE(4, 5, 2) SO_CASE(8, 4, 2) SO_CASE(8, 5, 2) SO_CASE(16, 4, 2) SO_CASE(16, 5, 2) SO_CASE(32, 4, 2)
  SO_CASE(32, 5, 2) SO_CASE(32, 8, 2)
#undef SO_CASE
  return SO_EUNSUPPORTED;
}

}  // namespace

int launch_line(const Shard& s, int fam, const Params& P) {
  if (P.k < 1 || P.k > 8) return SO_EINVAL;
  if (P.n <= small_max_n()) return fam == kLinear ? launch_small_line<kLinear>(s, P) : launch_small_line<kLogistic>(s, P);
  return fam == kLinear ? launch_mapped<kLinear>(s, P) : launch_mapped<kLogistic>(s, P);
}

}  // namespace wide
}  // namespace so
