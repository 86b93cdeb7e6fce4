// so_scalar.cu — evaluation kernels for the single-input ("scalar t") model families:
//   exponential decay, Gaussian peak, Gaussian peak + exponential background, polynomial.
//
// Data layout: t[m], y[m] as two fp64 device arrays (16 B per observation).  One thread owns whole
// observations: each loop trip loads two 128-bit words of t and two of y (4 observations), coalesced,
// streaming (ld.global.cs), evaluates the model and its parameter derivatives in registers and adds the
// row's contribution to per-thread accumulators.  Nothing but the Q-double result ever leaves the SM.
//
// The per-observation algebra replaces (reference paths relative to
// core/src/main/scala/com/github/bruneli/scalaopt/core/):
//   K1 loss+gradient   function/MSEFunction.scala:33-36,45-48,106-108,138-140
//   K2 normal eq.      function/MSEFunction.scala:117-136 + what linalg/QR.scala:136-151 extracts
//   K3 line search     variable/Point.scala:54-57 driven by linesearch/StrongWolfe.scala:96-118
//   K4 Hessian.vector  function/ObjectiveFunctionWithGradient.scala:41-47
//
// Roofline: 16 B/row of compulsory HBM traffic.  At 6.5 TB/s that is 4.1e11 rows/s, i.e. ~0.7 SM-cycles
// per row per SM, while the FP64 pipe retires 64 DFMA/clk/SM — so every FP64 instruction per row counts.
// Derivatives are therefore expressed as df/dp_a = D_a * u_a with D_a constant over the data set: the
// kernels accumulate sums of u_a u_b, u_a rho, rho^2 and the last block applies D on the way out.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <type_traits>

#include "so_device.cuh"
#include "so_internal.h"

namespace so {
namespace {

constexpr int kThreads = 256;
constexpr double kSqrt2 = 1.4142135623730951;
constexpr double kInvSqrt2 = 0.70710678118654752;

// ---------------------------------------------------------------------------------------------------
// Families.  Pre = per-evaluation constants (kernel parameter space), u = derivative basis.
// Evaluation is split in two so that the kernel can range-check all exp arguments of a group of rows with
// one branch and then run the branch-free so_exp_fast on all of them (cross-row ILP on the FP64 pipe):
//   args(pre, t, a)              -> the NE exp arguments of this row
//   finish(pre, t, ex, f, u)     -> model value and derivative basis from the NE exponentials
// ---------------------------------------------------------------------------------------------------
struct FamExpDecay {   // f = p0 exp(p1 t)                       LevenbergMarquardtSpec.scala:62-63
  static constexpr int N = 2, NE = 1;
  struct Pre { double p0, p1; double D[N]; };
  static Pre prepare(const double* p) { Pre c; c.p0 = p[0]; c.p1 = p[1]; c.D[0] = 1.0; c.D[1] = p[0]; return c; }
  __device__ static __forceinline__ void args(const Pre& c, double t, double* a) { a[0] = c.p1 * t; }
  __device__ static __forceinline__ void finish(const Pre& c, double t, const double* ex, double& f, double (&u)[N]) {
    u[0] = ex[0];          // df/dp0 = e
    u[1] = t * ex[0];      // df/dp1 = p0 t e
    f = c.p0 * ex[0];
  }
  __device__ static __forceinline__ void hdir(const Pre& c, double t, const double (&u)[N], const double* d, double (&h)[N]) {
    h[0] = u[1] * d[1];
    h[1] = fma(c.p0 * t * u[1], d[1], u[1] * d[0]);
  }
};

struct FamGauss {      // f = A exp(-(t-mu)^2 / (2 sigma^2)),  p = (A, mu, sigma)
  static constexpr int N = 3, NE = 1;
  struct Pre { double A, mu, s, sr2; double D[N]; };   // s = 1/sigma, sr2 = s/sqrt(2)
  static Pre prepare(const double* p) {
    Pre c; c.A = p[0]; c.mu = p[1]; c.s = 1.0 / p[2]; c.sr2 = c.s * kInvSqrt2;
    c.D[0] = 1.0; c.D[1] = p[0] * c.s * kSqrt2; c.D[2] = 2.0 * p[0] * c.s;
    return c;
  }
  __device__ static __forceinline__ void args(const Pre& c, double t, double* a) {
    const double zz = (t - c.mu) * c.sr2;          // z / sqrt(2),  z = (t - mu)/sigma
    a[0] = -zz * zz;
  }
  __device__ static __forceinline__ void finish(const Pre& c, double t, const double* ex, double& f, double (&u)[N]) {
    const double zz = (t - c.mu) * c.sr2;
    u[0] = ex[0];          // df/dA     = g
    u[1] = ex[0] * zz;     // df/dmu    = A g z / sigma      = (A s sqrt2) g zz
    u[2] = u[1] * zz;      // df/dsigma = A g z^2 / sigma    = (2 A s) g zz^2
    f = c.A * ex[0];
  }
  __device__ static __forceinline__ void hdir(const Pre& c, double t, const double (&u)[N], const double* d, double (&h)[N]) {
    const double z = (t - c.mu) * c.s, g = u[0], z2 = z * z;
    const double gs = g * c.s, Ags2 = c.A * gs * c.s;
    const double HAm = gs * z, HAs = gs * z2;
    const double Hmm = Ags2 * (z2 - 1.0), Hms = Ags2 * z * (z2 - 2.0), Hss = Ags2 * z2 * (z2 - 3.0);
    h[0] = HAm * d[1] + HAs * d[2];
    h[1] = HAm * d[0] + Hmm * d[1] + Hms * d[2];
    h[2] = HAs * d[0] + Hms * d[1] + Hss * d[2];
  }
};

struct FamGaussExpBkg {  // f = A exp(-(t-mu)^2/(2 sigma^2)) + B exp(-lambda t),  p = (A, mu, sigma, B, lambda)
  static constexpr int N = 5, NE = 2;   // functional form of stats/example/ExSignalPlusBackgroundFit.scala:78-114
  struct Pre { double A, mu, s, sr2, B, nlam; double D[N]; };
  static Pre prepare(const double* p) {
    Pre c; c.A = p[0]; c.mu = p[1]; c.s = 1.0 / p[2]; c.sr2 = c.s * kInvSqrt2; c.B = p[3]; c.nlam = -p[4];
    c.D[0] = 1.0; c.D[1] = p[0] * c.s * kSqrt2; c.D[2] = 2.0 * p[0] * c.s; c.D[3] = 1.0; c.D[4] = -p[3];
    return c;
  }
  __device__ static __forceinline__ void args(const Pre& c, double t, double* a) {
    const double zz = (t - c.mu) * c.sr2;
    a[0] = -zz * zz;
    a[1] = c.nlam * t;
  }
  __device__ static __forceinline__ void finish(const Pre& c, double t, const double* ex, double& f, double (&u)[N]) {
    const double zz = (t - c.mu) * c.sr2;
    u[0] = ex[0];
    u[1] = ex[0] * zz;
    u[2] = u[1] * zz;
    u[3] = ex[1];          // df/dB      = e
    u[4] = t * ex[1];      // df/dlambda = -B t e
    f = fma(c.A, ex[0], c.B * ex[1]);
  }
  __device__ static __forceinline__ void hdir(const Pre& c, double t, const double (&u)[N], const double* d, double (&h)[N]) {
    const double z = (t - c.mu) * c.s, g = u[0], z2 = z * z;
    const double gs = g * c.s, Ags2 = c.A * gs * c.s;
    const double HAm = gs * z, HAs = gs * z2;
    const double Hmm = Ags2 * (z2 - 1.0), Hms = Ags2 * z * (z2 - 2.0), Hss = Ags2 * z2 * (z2 - 3.0);
    h[0] = HAm * d[1] + HAs * d[2];
    h[1] = HAm * d[0] + Hmm * d[1] + Hms * d[2];
    h[2] = HAs * d[0] + Hms * d[1] + Hss * d[2];
    h[3] = -u[4] * d[4];                                  // d2f/dB dlambda = -t e
    h[4] = fma(c.B * t * u[4], d[4], -u[4] * d[3]);       // d2f/dlambda^2 = B t^2 e
  }
};

template <int NP>
struct FamPoly {       // f = sum_k p_k t^k
  static constexpr int N = NP, NE = 0;
  struct Pre { double p[N]; double D[N]; };
  static Pre prepare(const double* p) { Pre c; for (int k = 0; k < N; ++k) { c.p[k] = p[k]; c.D[k] = 1.0; } return c; }
  __device__ static __forceinline__ void args(const Pre&, double, double*) {}
  __device__ static __forceinline__ void finish(const Pre& c, double t, const double*, double& f, double (&u)[N]) {
    u[0] = 1.0;
    double s = c.p[0];
#pragma unroll
    for (int k = 1; k < N; ++k) { u[k] = u[k - 1] * t; s = fma(c.p[k], u[k], s); }
    f = s;
  }
  __device__ static __forceinline__ void hdir(const Pre&, double, const double (&)[N], const double*, double (&h)[N]) {
#pragma unroll
    for (int k = 0; k < N; ++k) h[k] = 0.0;
  }
};

// ---------------------------------------------------------------------------------------------------
// Per-kind row operators.  NEXP exp arguments per row, G rows per range-checked group.
// ---------------------------------------------------------------------------------------------------
template <class Fam, bool GRAD>
struct OpLossGrad {
  static constexpr int N = Fam::N, Q = GRAD ? 1 + N : 1, NEXP = Fam::NE, G = 4;
  struct Args { typename Fam::Pre pre; double scale[Q]; };
  __device__ static __forceinline__ void args(const Args& a, double t, double* ea) { Fam::args(a.pre, t, ea); }
  __device__ static __forceinline__ void accum(const Args& a, double t, double y, const double* ex, double (&acc)[Q]) {
    double f, u[N];
    Fam::finish(a.pre, t, ex, f, u);
    const double rho = y - f;
    acc[0] = fma(rho, rho, acc[0]);
    if constexpr (GRAD) {
#pragma unroll
      for (int i = 0; i < N; ++i) acc[1 + i] = fma(u[i], rho, acc[1 + i]);
    }
  }
};

template <class Fam, int GROUP = 4>
struct OpNormalEq {
  static constexpr int N = Fam::N, T = N * (N + 1) / 2, Q = T + N + 1, NEXP = Fam::NE, G = GROUP;
  struct Args { typename Fam::Pre pre; double scale[Q]; };
  __device__ static __forceinline__ void args(const Args& a, double t, double* ea) { Fam::args(a.pre, t, ea); }
  __device__ static __forceinline__ void accum(const Args& a, double t, double y, const double* ex, double (&acc)[Q]) {
    double f, u[N];
    Fam::finish(a.pre, t, ex, f, u);
    const double rho = y - f;
    int q = 0;
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int j = i; j < N; ++j) { acc[q] = fma(u[i], u[j], acc[q]); ++q; }
#pragma unroll
    for (int i = 0; i < N; ++i) acc[T + i] = fma(u[i], rho, acc[T + i]);
    acc[T + N] = fma(rho, rho, acc[T + N]);
  }
};

template <class Fam, int KB>
struct OpLine {
  static constexpr int N = Fam::N, Q = 2 * KB, NEXP = KB * Fam::NE, G = (KB <= 1 ? 4 : (KB <= 2 ? 2 : 1));
  struct Args { typename Fam::Pre pre[KB]; double dd[KB][N]; double scale[Q]; };   // dd = D(p + alpha d) o d
  __device__ static __forceinline__ void args(const Args& a, double t, double* ea) {
#pragma unroll
    for (int k = 0; k < KB; ++k) Fam::args(a.pre[k], t, ea + k * Fam::NE);
  }
  __device__ static __forceinline__ void accum(const Args& a, double t, double y, const double* ex, double (&acc)[Q]) {
#pragma unroll
    for (int k = 0; k < KB; ++k) {
      double f, u[N];
      Fam::finish(a.pre[k], t, ex + k * Fam::NE, f, u);
      const double rho = y - f;
      double jd = 0.0;
#pragma unroll
      for (int i = 0; i < N; ++i) jd = fma(a.dd[k][i], u[i], jd);
      acc[k] = fma(rho, rho, acc[k]);
      acc[KB + k] = fma(rho, jd, acc[KB + k]);
    }
  }
};

template <class Fam>
struct OpHessVec {
  static constexpr int N = Fam::N, Q = N, NEXP = Fam::NE, G = 2;
  struct Args { typename Fam::Pre pre; double d[N]; double scale[Q]; };
  __device__ static __forceinline__ void args(const Args& a, double t, double* ea) { Fam::args(a.pre, t, ea); }
  __device__ static __forceinline__ void accum(const Args& a, double t, double y, const double* ex, double (&acc)[Q]) {
    double f, u[N], h[N], J[N];
    Fam::finish(a.pre, t, ex, f, u);
    const double rho = y - f;
    double jd = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) { J[i] = a.pre.D[i] * u[i]; jd = fma(J[i], a.d[i], jd); }
    Fam::hdir(a.pre, t, u, a.d, h);
#pragma unroll
    for (int i = 0; i < N; ++i) acc[i] += fma(jd, J[i], -rho * h[i]);
  }
};

// G rows: all exp arguments first, ONE range check, then branch-free exponentials, then the accumulation.
template <class Op, int G>
__device__ __forceinline__ void process_rows(const typename Op::Args& args, const double (&t)[G], const double (&y)[G],
                                             double (&acc)[Op::Q]) {
  constexpr int NE = Op::NEXP;
  if constexpr (NE == 0) {
#pragma unroll
    for (int r = 0; r < G; ++r) Op::accum(args, t[r], y[r], nullptr, acc);
  } else {
    double ea[G][NE];
    int mx = 0;
#pragma unroll
    for (int r = 0; r < G; ++r) {
      Op::args(args, t[r], ea[r]);
#pragma unroll
      for (int e = 0; e < NE; ++e) mx = max(mx, exp_abs_hi(ea[r][e]));
    }
    if (mx < kExpHiLimit) {
#pragma unroll
      for (int r = 0; r < G; ++r)
#pragma unroll
        for (int e = 0; e < NE; ++e) ea[r][e] = so_exp_fast(ea[r][e]);
    } else {      // some |arg| >= 708, inf or NaN in this group: full-range libdevice exp for the whole group
#pragma unroll
      for (int r = 0; r < G; ++r)
#pragma unroll
        for (int e = 0; e < NE; ++e) ea[r][e] = exp_slow(ea[r][e]);
    }
#pragma unroll
    for (int r = 0; r < G; ++r) Op::accum(args, t[r], y[r], ea[r], acc);
  }
}

// a loop trip carries 4 rows (two 128-bit loads of t and of y); they are processed in groups of Op::G
template <class Op>
__device__ __forceinline__ void process4(const typename Op::Args& args, double2 ta, double2 ya, double2 tb, double2 yb,
                                         double (&acc)[Op::Q]) {
  constexpr int G = Op::G;
  const double t4[4] = {ta.x, ta.y, tb.x, tb.y}, y4[4] = {ya.x, ya.y, yb.x, yb.y};
#pragma unroll
  for (int g0 = 0; g0 < 4; g0 += G) {
    double tg[G], yg[G];
#pragma unroll
    for (int r = 0; r < G; ++r) { tg[r] = t4[g0 + r]; yg[r] = y4[g0 + r]; }
    process_rows<Op, G>(args, tg, yg, acc);
  }
}
template <class Op>
__device__ __forceinline__ void process1(const typename Op::Args& args, double t, double y, double (&acc)[Op::Q]) {
  const double tg[1] = {t}, yg[1] = {y};
  process_rows<Op, 1>(args, tg, yg, acc);
}

// ---------------------------------------------------------------------------------------------------
// The kernel: grid-stride over pairs of observations, 2 pairs (4 observations) in flight per thread.
// ---------------------------------------------------------------------------------------------------
template <class Op, int MINBLOCKS>
__global__ void __launch_bounds__(kThreads, MINBLOCKS)
k_scalar(const double* __restrict__ t, const double* __restrict__ y, int64_t rows,
         const __grid_constant__ typename Op::Args args,
         double* __restrict__ partials, double* __restrict__ out, unsigned int* ticket) {
  constexpr int Q = Op::Q;
  __shared__ double s_red[(kThreads / 32) * Q];
  if constexpr (Op::NEXP > 0) {
    exp_table_init();
    __syncthreads();
  }

  double acc[Q];
#pragma unroll
  for (int q = 0; q < Q; ++q) acc[q] = 0.0;

  // a slice may start on an odd row: peel it so the 128-bit loads stay aligned
  const int head = (int)((reinterpret_cast<uintptr_t>(t) >> 3) & 1) && rows > 0 ? 1 : 0;
  const double2* __restrict__ t2 = reinterpret_cast<const double2*>(t + head);
  const double2* __restrict__ y2 = reinterpret_cast<const double2*>(y + head);
  const int64_t npairs = (rows - head) >> 1;
  const int tail = (int)((rows - head) & 1);
  const int64_t gtid = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  const int64_t gstride = (int64_t)gridDim.x * kThreads;

  int64_t i = gtid;
  for (; i + gstride < npairs; i += 2 * gstride) {
    const double2 ta = ld_stream2(t2 + i), ya = ld_stream2(y2 + i);
    const double2 tb = ld_stream2(t2 + i + gstride), yb = ld_stream2(y2 + i + gstride);
    process4<Op>(args, ta, ya, tb, yb, acc);
  }
  if (i < npairs) {
    const double2 ta = ld_stream2(t2 + i), ya = ld_stream2(y2 + i);
    process1<Op>(args, ta.x, ya.x, acc);
    process1<Op>(args, ta.y, ya.y, acc);
  }
  if (gtid == 0 && head) process1<Op>(args, t[0], y[0], acc);
  if (gtid == gstride - 1 && tail) process1<Op>(args, t[rows - 1], y[rows - 1], acc);

  grid_reduce<Q>(acc, s_red, partials, out, args.scale, ticket);
}

int grid_for(const void* kernel, const Shard& s) {
  int per_sm = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, 0);
  if (per_sm < 1) per_sm = 1;
  int64_t want = (s.rows / 2 + kThreads - 1) / kThreads;
  int64_t cap = (int64_t)per_sm * s.sm_count;
  if (want < 1) want = 1;
  return (int)std::min<int64_t>(want, cap);
}

template <class Op, int MINBLOCKS>
int launch_op(const Shard& s, const typename Op::Args& args) {
  auto kern = k_scalar<Op, MINBLOCKS>;
  const int grid = grid_for(reinterpret_cast<const void*>(kern), s);
  if ((size_t)grid * Op::Q > s.partials_doubles || (size_t)Op::Q > s.out_doubles) return SO_ENOMEM;
  kern<<<grid, kThreads, 0, s.stream>>>(s.X, s.y, s.rows, args, s.partials, s.out, s.ticket);
  return cudaGetLastError() == cudaSuccess ? 1 : SO_ECUDA;
}

template <class Fam>
int launch_family(Kind kind, const Shard& s, const double* p, const double* d, const double* alpha, int k, int want_grad) {
  constexpr int N = Fam::N;
  const typename Fam::Pre pre = Fam::prepare(p);
  switch (kind) {
    case kLossGrad: {
      if (want_grad) {
        using Op = OpLossGrad<Fam, true>;
        typename Op::Args a; a.pre = pre;
        a.scale[0] = 0.5;
        for (int i = 0; i < N; ++i) a.scale[1 + i] = -pre.D[i];
        return launch_op<Op, 3>(s, a);
      }
      using Op = OpLossGrad<Fam, false>;
      typename Op::Args a; a.pre = pre; a.scale[0] = 0.5;
      return launch_op<Op, 4>(s, a);
    }
    case kNormalEq: {
      using Op = OpNormalEq<Fam>;
      typename Op::Args a; a.pre = pre;
      int q = 0;
      for (int i = 0; i < N; ++i) for (int j = i; j < N; ++j) a.scale[q++] = pre.D[i] * pre.D[j];
      for (int i = 0; i < N; ++i) a.scale[Op::T + i] = -pre.D[i];     // J = -sign(rho) df/dp, r = |rho|  =>  J r = -df/dp rho
      a.scale[Op::T + N] = 1.0;
      if constexpr (std::is_same<Fam, FamGaussExpBkg>::value) {
        // tuning variants (rows per range-checked group x resident blocks per SM), SO_K2_VARIANT=0..3
        const char* v = getenv("SO_K2_VARIANT");
        const int var = v ? atoi(v) : 0;
        auto relaunch = [&](auto optag, auto mbtag) -> int {
          using Op2 = OpNormalEq<Fam, decltype(optag)::value>;
          typename Op2::Args a2; a2.pre = a.pre;
          for (int q2 = 0; q2 < Op2::Q; ++q2) a2.scale[q2] = a.scale[q2];
          return launch_op<Op2, decltype(mbtag)::value>(s, a2);
        };
        if (var == 1) return relaunch(std::integral_constant<int, 2>(), std::integral_constant<int, 3>());
        if (var == 2) return relaunch(std::integral_constant<int, 2>(), std::integral_constant<int, 2>());
        if (var == 3) return relaunch(std::integral_constant<int, 1>(), std::integral_constant<int, 3>());
        if (var == 4) return relaunch(std::integral_constant<int, 4>(), std::integral_constant<int, 3>());
        return launch_op<Op, 2>(s, a);
      }
      return launch_op<Op, (Op::Q < 21 ? 3 : 2)>(s, a);
    }
    case kLine: {
      // batches of up to 8 step sizes per pass; longer schedules take several passes into consecutive out slots
      // (the API layer asks for at most 8 per call for these families)
      if (k < 1 || k > 8) return SO_EINVAL;
      auto run = [&](auto KBtag) -> int {
        constexpr int KB = decltype(KBtag)::value;
        using Op = OpLine<Fam, KB>;
        typename Op::Args a;
        double pk[N];
        for (int kk = 0; kk < KB; ++kk) {
          const double al = alpha[kk < k ? kk : k - 1];          // pad by repeating the last step size
          for (int i = 0; i < N; ++i) pk[i] = p[i] + d[i] * al;  // ptInit.x + ptInit.d * a1, StrongWolfe.scala:97
          a.pre[kk] = Fam::prepare(pk);
          for (int i = 0; i < N; ++i) a.dd[kk][i] = a.pre[kk].D[i] * d[i];
          a.scale[kk] = 0.5;
          a.scale[KB + kk] = -1.0;
        }
        return launch_op<Op, (KB >= 8 ? 1 : 2)>(s, a);
      };
      if (k == 1) return run(std::integral_constant<int, 1>());
      if (k == 2) return run(std::integral_constant<int, 2>());
      if (k <= 4) return run(std::integral_constant<int, 4>());
      return run(std::integral_constant<int, 8>());
    }
    case kHessVec: {
      using Op = OpHessVec<Fam>;
      typename Op::Args a; a.pre = pre;
      for (int i = 0; i < N; ++i) { a.d[i] = d[i]; a.scale[i] = 1.0; }
      return launch_op<Op, 3>(s, a);
    }
  }
  return SO_EINVAL;
}

}  // namespace

size_t scalar_partials_needed(Kind kind, int n, int k, int sm_count) {
  // at most 8 resident 256-thread blocks per SM, Q <= 45 (polynomial n = 8 normal equations)
  const int Q = std::max(out_doubles(kind, n, 8), 16);
  (void)k;
  return (size_t)sm_count * 8 * (size_t)Q;
}

// padded width of the kLine result for the scalar families (so_api reads phi at [0,k) and dphi at [KB, KB+k))
int scalar_line_block(int k) { return k == 1 ? 1 : (k == 2 ? 2 : (k <= 4 ? 4 : 8)); }

int launch_scalar(Kind kind, const Shard& s, int family, const double* p, const double* d, int n,
                  const double* alpha, int k, int want_grad) {
  if (s.f != 1) return SO_EINVAL;
  switch (family) {
    case SO_FAMILY_EXPDECAY: if (n != 2) return SO_EINVAL; return launch_family<FamExpDecay>(kind, s, p, d, alpha, k, want_grad);
    case SO_FAMILY_GAUSS: if (n != 3) return SO_EINVAL; return launch_family<FamGauss>(kind, s, p, d, alpha, k, want_grad);
    case SO_FAMILY_GAUSS_EXPBKG: if (n != 5) return SO_EINVAL; return launch_family<FamGaussExpBkg>(kind, s, p, d, alpha, k, want_grad);
    case SO_FAMILY_POLY:
      switch (n) {
        case 1: return launch_family<FamPoly<1>>(kind, s, p, d, alpha, k, want_grad);
        case 2: return launch_family<FamPoly<2>>(kind, s, p, d, alpha, k, want_grad);
        case 3: return launch_family<FamPoly<3>>(kind, s, p, d, alpha, k, want_grad);
        case 4: return launch_family<FamPoly<4>>(kind, s, p, d, alpha, k, want_grad);
        case 5: return launch_family<FamPoly<5>>(kind, s, p, d, alpha, k, want_grad);
        case 6: return launch_family<FamPoly<6>>(kind, s, p, d, alpha, k, want_grad);
        case 7: return launch_family<FamPoly<7>>(kind, s, p, d, alpha, k, want_grad);
        case 8: return launch_family<FamPoly<8>>(kind, s, p, d, alpha, k, want_grad);
        default: return SO_EUNSUPPORTED;
      }
    default: return SO_EUNSUPPORTED;
  }
}

}  // namespace so
