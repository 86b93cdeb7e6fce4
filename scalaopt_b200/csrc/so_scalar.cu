// so_scalar.cu — evaluation kernels for the single-input ("scalar t") model families:
//   exponential decay, Gaussian peak, Gaussian peak + exponential background, polynomial.
//
// Data layout: t[m], y[m] as two fp64 device arrays (16 B per observation).  Blocks stream chunks of 1024
// observations through a shared-memory ring filled by TMA bulk copies; one thread owns whole observations
// (4 per chunk), evaluates the model and its parameter derivatives in registers and adds the row's
// contribution to per-thread accumulators.  Nothing but the Q-double result ever leaves the SM.
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
#include <utility>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <type_traits>

#include "so_device.cuh"
#include "so_internal.h"
#include "so_qr.cuh"

namespace so {
namespace {

constexpr int kThreads = 256;
constexpr double kSqrt2 = 1.4142135623730951;
constexpr double kInvSqrt2 = 0.70710678118654752;

// Range check of the checked kernels, on the FP32 pipe and on values the row has anyway: every exp argument is either
// -(v^2) or v * const, so "|argument| < 700" is |v| < limit with a per-evaluation limit.  Limits travel as the HIGH WORD
// of the double (read as a float it orders like the double's magnitude; an inf/NaN v has an inf/NaN high word, which
// compares false): one FSETP with a free |.| per exponential, no FP64 instruction, no dependence on the exp's own chain.
inline int hi_word(double v) { long long b; std::memcpy(&b, &v, sizeof b); return (int)(b >> 32); }
inline int lim_linear(double scaled_rate) { return hi_word(kExpScaledLimit / std::fabs(scaled_rate)); }     // rate 0 -> inf: any finite v passes
inline int lim_square() { return hi_word(std::sqrt(kExpScaledLimit)); }
__device__ __forceinline__ bool hi_below(double v, int limit_hi) { return fabsf(__int_as_float(__double2hiint(v))) < __int_as_float(limit_hi); }

// ---------------------------------------------------------------------------------------------------
// Families.  Pre = per-evaluation constants (kernel parameter space), u = derivative basis.
// Evaluation is split in two so that the kernel can range-check all exp arguments of a group of rows with
// one branch and then run the branch-free exponential on all of them (cross-row ILP on the FP64 pipe):
//   sargs(pre, t, p, q)          -> the NE exp arguments of this row as PRODUCTS p q in units of ln2/256 (so_exps in
//                                   so_device.cuh: the 256/ln2 is folded into the per-evaluation constants, the
//                                   product is never formed: one FP64 instruction less per exponential)
//   finish(pre, t, ex, f, u)     -> model value and derivative basis from the NE exponentials
// The Gaussian's argument is -(zz')^2 with zz' = sqrt(256/ln2) z/sqrt2, and zz' is also what the derivative basis is
// built from (u1 = g zz', u2 = g zz'^2): the powers of sqrt(256/ln2) go into D.
// ---------------------------------------------------------------------------------------------------
struct FamExpDecay {   // f = p0 exp(p1 t)                       LevenbergMarquardtSpec.scala:62-63
  static constexpr int N = 2, NE = 1;
  __host__ __device__ static constexpr bool gram_alias(int, int, int&, int&) { return false; }
  struct Pre { double p0, p1s; double D[N]; int tlim; };          // p1s = p1 * 256/ln2
  static Pre prepare(const double* p) { Pre c; c.p0 = p[0]; c.p1s = p[1] * kExpScale; c.D[0] = 1.0; c.D[1] = p[0]; c.tlim = lim_linear(c.p1s); return c; }
  __device__ static __forceinline__ void sargs(const Pre& c, double t, double* p, double* q) { p[0] = t; q[0] = c.p1s; }
  __device__ static __forceinline__ bool in_range(const Pre& c, double t, const double*, const double*) { return hi_below(t, c.tlim); }
  __device__ static __forceinline__ void finish(const Pre& c, double t, const double* ex, double& f, double (&u)[N]) {
    u[0] = ex[0];          // df/dp0 = e
    u[1] = t * ex[0];      // df/dp1 = p0 t e
    f = c.p0 * ex[0];
  }
  __device__ static __forceinline__ double residual(const Pre& c, double y, const double* ex, double) { return fma(-c.p0, ex[0], y); }
  // sum_i dd_i u_i without forming u (the line-search kernel needs only this combination): Horner in t
  __device__ static __forceinline__ double dirder(const Pre&, const double* dd, double t, const double* ex) { return ex[0] * fma(t, dd[1], dd[0]); }
  // B_i += rho * (d2f/dp2 . d)_i :  h0 = t e d1,  h1 = t e d0 + p0 t^2 e d1
  struct HPre { double d0, d1, p0d1; };
  static HPre hprepare(const double* p, const double* d) { HPre h; h.d0 = d[0]; h.d1 = d[1]; h.p0d1 = p[0] * d[1]; return h; }
  __device__ static __forceinline__ void hess_b(const Pre&, const HPre& h, double t, const double (&u)[N], double rho, double* B) {
    const double w = rho * u[1];                       // rho t e
    B[0] = fma(w, h.d1, B[0]);
    B[1] = fma(w, fma(h.p0d1, t, h.d0), B[1]);
  }
};

struct FamGauss {      // f = A exp(-(t-mu)^2 / (2 sigma^2)),  p = (A, mu, sigma)
  static constexpr int N = 3, NE = 1;
  // u1*u1 = (g zz)^2 = g * (g zz^2) = u0*u2: the (1,1) sum is the (0,2) sum, accumulate it once
  __host__ __device__ static constexpr bool gram_alias(int i, int j, int& si, int& sj) { if (i == 1 && j == 1) { si = 0; sj = 2; return true; } return false; }
  struct Pre { double A, mu, s, sr2, nmsr2; double D[N]; int zlim; };   // s = 1/sigma, sr2 = sqrt(256/ln2) s/sqrt(2), nmsr2 = -mu sr2
  static Pre prepare(const double* p) {
    Pre c; c.A = p[0]; c.mu = p[1]; c.s = 1.0 / p[2]; c.sr2 = c.s * kInvSqrt2 * kExpScaleSqrt; c.nmsr2 = -c.mu * c.sr2;
    c.D[0] = 1.0; c.D[1] = p[0] * c.s * kSqrt2 / kExpScaleSqrt; c.D[2] = 2.0 * p[0] * c.s / kExpScale;
    c.zlim = lim_square();
    return c;
  }
  __device__ static __forceinline__ void sargs(const Pre& c, double t, double* p, double* q) {
    const double zz = fma(t, c.sr2, c.nmsr2);      // zz' = sqrt(256/ln2) z / sqrt(2),  z = (t - mu)/sigma, one FMA
    p[0] = -zz; q[0] = zz;
  }
  __device__ static __forceinline__ bool in_range(const Pre& c, double, const double*, const double* q) { return hi_below(q[0], c.zlim); }
  __device__ static __forceinline__ void finish(const Pre& c, double t, const double* ex, double& f, double (&u)[N]) {
    const double zz = fma(t, c.sr2, c.nmsr2);
    u[0] = ex[0];          // df/dA     = g
    u[1] = ex[0] * zz;     // df/dmu    = A g z / sigma      = (A s sqrt2 / S) g zz',   S = sqrt(256/ln2)
    u[2] = u[1] * zz;      // df/dsigma = A g z^2 / sigma    = (2 A s / S^2) g zz'^2
    f = c.A * ex[0];
  }
  __device__ static __forceinline__ double residual(const Pre& c, double y, const double* ex, double) { return fma(-c.A, ex[0], y); }
  __device__ static __forceinline__ double dirder(const Pre& c, const double* dd, double t, const double* ex) {
    const double zz = fma(t, c.sr2, c.nmsr2);
    return ex[0] * fma(zz, fma(zz, dd[2], dd[1]), dd[0]);          // g (dd0 + zz dd1 + zz^2 dd2)
  }
  // B_i += rho * (d2f/dp2 . d)_i in Horner form in v = zz' = S z/sqrt2 with host-side coefficients (see hgauss):
  //   h0 = g v (a0 + a1 v),  h1 = g (b0 + v (b1 + v (b2 + v b3))),  h2 = g v (c0 + v (c1 + v (c2 + v c3)))
  struct HGauss { double a0, a1, b0, b1, b2, b3, c0, c1, c2, c3; };
  static HGauss hgauss(const double* p, const double* d) {
    const double A = p[0], s = 1.0 / p[2], r2 = kSqrt2, As2 = A * s * s;
    HGauss h;
    h.a0 = s * d[1] * r2;            h.a1 = s * d[2] * 2.0;                    // z = r2 v, z^2 = 2 v^2
    h.b0 = -As2 * d[1];              h.b1 = (s * d[0] - 2.0 * As2 * d[2]) * r2;
    h.b2 = As2 * d[1] * 2.0;         h.b3 = As2 * d[2] * 2.0 * r2;
    h.c0 = -2.0 * As2 * d[1] * r2;   h.c1 = (s * d[0] - 3.0 * As2 * d[2]) * 2.0;
    h.c2 = As2 * d[1] * 2.0 * r2;    h.c3 = As2 * d[2] * 4.0;
    // the kernels' v is S z/sqrt2: every power of v (incl. the one in "g v") takes a factor 1/S
    const double iS = 1.0 / kExpScaleSqrt, iL = 1.0 / kExpScale;
    h.a0 *= iS; h.a1 *= iL;
    h.b1 *= iS; h.b2 *= iL; h.b3 *= iL * iS;
    h.c0 *= iS; h.c1 *= iL; h.c2 *= iL * iS; h.c3 *= iL * iL;
    return h;
  }
  __device__ static __forceinline__ void hess_gauss(const HGauss& h, double v, double g, double gv, double rho, double* B) {
    const double rg = rho * g, rgv = rho * gv;
    B[0] = fma(rgv, fma(h.a1, v, h.a0), B[0]);
    B[1] = fma(rg, fma(fma(fma(h.b3, v, h.b2), v, h.b1), v, h.b0), B[1]);
    B[2] = fma(rgv, fma(fma(fma(h.c3, v, h.c2), v, h.c1), v, h.c0), B[2]);
  }
  struct HPre { HGauss g; };
  static HPre hprepare(const double* p, const double* d) { HPre h; h.g = hgauss(p, d); return h; }
  __device__ static __forceinline__ void hess_b(const Pre& c, const HPre& h, double t, const double (&u)[N], double rho, double* B) {
    hess_gauss(h.g, fma(t, c.sr2, c.nmsr2), u[0], u[1], rho, B);
  }
};

struct FamGaussExpBkg {  // f = A exp(-(t-mu)^2/(2 sigma^2)) + B exp(-lambda t),  p = (A, mu, sigma, B, lambda)
  static constexpr int N = 5, NE = 2;   // functional form of stats/example/ExSignalPlusBackgroundFit.scala:78-114
  __host__ __device__ static constexpr bool gram_alias(int i, int j, int& si, int& sj) { if (i == 1 && j == 1) { si = 0; sj = 2; return true; } return false; }
  struct Pre { double A, mu, s, sr2, nmsr2, B, nlam; double D[N]; int zlim, tlim; };    // sr2, nmsr2 as in FamGauss; nlam = -lambda 256/ln2
  static Pre prepare(const double* p) {
    Pre c; c.A = p[0]; c.mu = p[1]; c.s = 1.0 / p[2]; c.sr2 = c.s * kInvSqrt2 * kExpScaleSqrt; c.nmsr2 = -c.mu * c.sr2; c.B = p[3]; c.nlam = -p[4] * kExpScale;
    c.D[0] = 1.0; c.D[1] = p[0] * c.s * kSqrt2 / kExpScaleSqrt; c.D[2] = 2.0 * p[0] * c.s / kExpScale; c.D[3] = 1.0; c.D[4] = -p[3];
    c.zlim = lim_square(); c.tlim = lim_linear(c.nlam);
    return c;
  }
  __device__ static __forceinline__ void sargs(const Pre& c, double t, double* p, double* q) {
    const double zz = fma(t, c.sr2, c.nmsr2);
    p[0] = -zz; q[0] = zz;
    p[1] = t; q[1] = c.nlam;
  }
  __device__ static __forceinline__ bool in_range(const Pre& c, double t, const double*, const double* q) { return hi_below(q[0], c.zlim) && hi_below(t, c.tlim); }
  __device__ static __forceinline__ void finish(const Pre& c, double t, const double* ex, double& f, double (&u)[N]) {
    const double zz = fma(t, c.sr2, c.nmsr2);
    u[0] = ex[0];
    u[1] = ex[0] * zz;
    u[2] = u[1] * zz;
    u[3] = ex[1];          // df/dB      = e
    u[4] = t * ex[1];      // df/dlambda = -B t e
    f = fma(c.A, ex[0], c.B * ex[1]);
  }
  // y - f in two FMAs (the accumulating kernels only need the residual)
  __device__ static __forceinline__ double residual(const Pre& c, double y, const double* ex, double) {
    return fma(-c.B, ex[1], fma(-c.A, ex[0], y));
  }
  __device__ static __forceinline__ double dirder(const Pre& c, const double* dd, double t, const double* ex) {
    const double zz = fma(t, c.sr2, c.nmsr2);
    const double h1 = fma(zz, fma(zz, dd[2], dd[1]), dd[0]), h2 = fma(t, dd[4], dd[3]);
    return fma(ex[0], h1, ex[1] * h2);                             // g (dd0 + zz dd1 + zz^2 dd2) + e (dd3 + t dd4): 5 instead of 3 + 5
  }
  // true when both exp arguments stay inside (-700, 700) for every t in [tmin, tmax] (they are monotone or convex in t,
  // so the end points decide); NaN anywhere compares false.  708 is the limit of so_exps: 1 % of margin for rounding.
  static bool args_in_range(const Pre& c, double tmin, double tmax) {
    const double z0 = std::fabs(tmin * c.sr2 + c.nmsr2), z1 = std::fabs(tmax * c.sr2 + c.nmsr2);
    const double zm = z0 > z1 ? z0 : z1, tm = std::fabs(tmin) > std::fabs(tmax) ? std::fabs(tmin) : std::fabs(tmax);
    return zm * zm < kExpScaledLimit && std::fabs(c.nlam) * tm < kExpScaledLimit;
  }
  // Gaussian part as in FamGauss; exponential part: h3 = -t e d4, h4 = -t e d3 + B t^2 e d4
  struct HPre { FamGauss::HGauss g; double nd4, nd3, Bd4; };
  static HPre hprepare(const double* p, const double* d) {
    HPre h; h.g = FamGauss::hgauss(p, d); h.nd4 = -d[4]; h.nd3 = -d[3]; h.Bd4 = p[3] * d[4]; return h;
  }
  __device__ static __forceinline__ void hess_b(const Pre& c, const HPre& h, double t, const double (&u)[N], double rho, double* B) {
    FamGauss::hess_gauss(h.g, fma(t, c.sr2, c.nmsr2), u[0], u[1], rho, B);
    const double w = rho * u[4];                       // rho t e
    B[3] = fma(w, h.nd4, B[3]);
    B[4] = fma(w, fma(h.Bd4, t, h.nd3), B[4]);
  }
};

template <int NP>
struct FamPoly {       // f = sum_k p_k t^k
  static constexpr int N = NP, NE = 0;
  // u_i u_j = t^(i+j): every anti-diagonal of the Gram is one sum; keep the entry on the first row / last column
  __host__ __device__ static constexpr bool gram_alias(int i, int j, int& si, int& sj) {
    if (i == 0 || j == NP - 1) return false;
    const int s = i + j; si = (s < NP) ? 0 : s - (NP - 1); sj = s - si; return true;
  }
  struct Pre { double p[N]; double D[N]; };
  static Pre prepare(const double* p) { Pre c; for (int k = 0; k < N; ++k) { c.p[k] = p[k]; c.D[k] = 1.0; } return c; }
  __device__ static __forceinline__ void sargs(const Pre&, double, double*, double*) {}
  __device__ static __forceinline__ bool in_range(const Pre&, double, const double*, const double*) { return true; }
  __device__ static __forceinline__ void finish(const Pre& c, double t, const double*, double& f, double (&u)[N]) {
    u[0] = 1.0;
    double s = c.p[0];
#pragma unroll
    for (int k = 1; k < N; ++k) { u[k] = u[k - 1] * t; s = fma(c.p[k], u[k], s); }
    f = s;
  }
  __device__ static __forceinline__ double residual(const Pre&, double y, const double*, double f) { return y - f; }
  struct HPre { double unused; };                    // linear in the parameters: no second-derivative term
  static HPre hprepare(const double*, const double*) { return HPre{0.0}; }
  __device__ static __forceinline__ void hess_b(const Pre&, const HPre&, double, const double (&)[N], double, double*) {}
};

// ---------------------------------------------------------------------------------------------------
// Per-kind row operators.  NEXP exp arguments per row, G rows per range-checked group.
// ---------------------------------------------------------------------------------------------------
template <class Fam, bool GRAD, int GROUP = 4>
struct OpLossGrad {
  static constexpr int N = Fam::N, Q = GRAD ? 1 + N : 1, NEXP = Fam::NE, G = GROUP;
  static constexpr bool kHasSrc = false;
  static constexpr bool kNeedsY = true;
  struct Args { typename Fam::Pre pre; double scale[Q]; };
  __device__ static __forceinline__ void sargs(const Args& a, double t, double* p, double* q) { Fam::sargs(a.pre, t, p, q); }
  __device__ static __forceinline__ bool in_range(const Args& a, double t, const double* p, const double* q) { return Fam::in_range(a.pre, t, p, q); }
  __device__ static __forceinline__ void accum(const Args& a, double t, double y, const double* ex, double (&acc)[Q]) {
    double f, u[N];
    Fam::finish(a.pre, t, ex, f, u);
    const double rho = Fam::residual(a.pre, y, ex, f);
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
  struct Args { typename Fam::Pre pre; double scale[Q]; signed char src[Q]; };
  static constexpr bool kHasSrc = true;
  static constexpr bool kNeedsY = true;
  __host__ __device__ static constexpr int packed(int i, int j) { return i * N - i * (i - 1) / 2 + (j - i); }
  __device__ static __forceinline__ void sargs(const Args& a, double t, double* p, double* q) { Fam::sargs(a.pre, t, p, q); }
  __device__ static __forceinline__ bool in_range(const Args& a, double t, const double* p, const double* q) { return Fam::in_range(a.pre, t, p, q); }
  __device__ static __forceinline__ void accum(const Args& a, double t, double y, const double* ex, double (&acc)[Q]) {
    double f, u[N];
    Fam::finish(a.pre, t, ex, f, u);
    const double rho = Fam::residual(a.pre, y, ex, f);
    int q = 0;
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int j = i; j < N; ++j) {
        int si = 0, sj = 0;
        if (!Fam::gram_alias(i, j, si, sj)) acc[q] = fma(u[i], u[j], acc[q]);   // aliased sums are accumulated once
        ++q;
      }
#pragma unroll
    for (int i = 0; i < N; ++i) acc[T + i] = fma(u[i], rho, acc[T + i]);
    acc[T + N] = fma(rho, rho, acc[T + N]);
  }
};

template <class Fam, int KB, int GROUP = (KB <= 1 ? 4 : (KB <= 2 ? 2 : 1))>
struct OpLine {
  static constexpr int N = Fam::N, Q = 2 * KB, NEXP = KB * Fam::NE, G = GROUP;
  static constexpr bool kHasSrc = false;
  static constexpr bool kNeedsY = true;
  struct Args { typename Fam::Pre pre[KB]; double dd[KB][N]; double scale[Q]; };   // dd = D(p + alpha d) o d
  __device__ static __forceinline__ void sargs(const Args& a, double t, double* p, double* q) {
#pragma unroll
    for (int k = 0; k < KB; ++k) Fam::sargs(a.pre[k], t, p + k * Fam::NE, q + k * Fam::NE);
  }
  __device__ static __forceinline__ bool in_range(const Args& a, double t, const double* p, const double* q) {
    bool ok = true;
#pragma unroll
    for (int k = 0; k < KB; ++k) ok = ok && Fam::in_range(a.pre[k], t, p + k * Fam::NE, q + k * Fam::NE);
    return ok;
  }
  __device__ static __forceinline__ void accum(const Args& a, double t, double y, const double* ex, double (&acc)[Q]) {
#pragma unroll
    for (int k = 0; k < KB; ++k) {
      double rho, jd;
      if constexpr (requires { Fam::dirder(a.pre[k], a.dd[k], t, ex); }) {     // residual from the exponentials, J.d in Horner form
        rho = Fam::residual(a.pre[k], y, ex + k * Fam::NE, 0.0);
        jd = Fam::dirder(a.pre[k], a.dd[k], t, ex + k * Fam::NE);
      } else {
        double f, u[N];
        Fam::finish(a.pre[k], t, ex + k * Fam::NE, f, u);
        rho = Fam::residual(a.pre[k], y, ex + k * Fam::NE, f);
        jd = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) jd = fma(a.dd[k][i], u[i], jd);
      }
      acc[k] = fma(rho, rho, acc[k]);
      acc[KB + k] = fma(rho, jd, acc[KB + k]);
    }
  }
};

template <class Fam, int GROUP = 2>
struct OpHessVec {
  // H d = (1/m) sum_i [ (J_f . d) J_f - rho (d2f/dp2 . d) ]  with J_f = D o u:
  //   acc[0..N)  = sum (dd . u) u_i      (dd = D o d; scaled by D_i on the way out)
  //   acc[N..2N) = sum rho h_i           (subtracted on the way out)
  static constexpr int N = Fam::N, Q = 2 * N, NEXP = Fam::NE, G = GROUP;
  static constexpr bool kHasSrc = false;
  static constexpr bool kNeedsY = true;
  struct Args { typename Fam::Pre pre; typename Fam::HPre h; double dd[N]; double scale[N]; double scale2[N]; };
  __device__ static __forceinline__ void sargs(const Args& a, double t, double* p, double* q) { Fam::sargs(a.pre, t, p, q); }
  __device__ static __forceinline__ bool in_range(const Args& a, double t, const double* p, const double* q) { return Fam::in_range(a.pre, t, p, q); }
  __device__ static __forceinline__ void accum(const Args& a, double t, double y, const double* ex, double (&acc)[Q]) {
    double f, u[N];
    Fam::finish(a.pre, t, ex, f, u);
    const double rho = Fam::residual(a.pre, y, ex, f);
    double jd = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) jd = fma(a.dd[i], u[i], jd);
#pragma unroll
    for (int i = 0; i < N; ++i) acc[i] = fma(jd, u[i], acc[i]);
    Fam::hess_b(a.pre, a.h, t, u, rho, &acc[N]);
  }
};

// ---------------------------------------------------------------------------------------------------
// Negative log-likelihood: sum_i -2 log pdf(p, x_i), std-apps/.../stats/package.scala:97-103,
// pdf = N(x; mu, sigma) fSig + lambda exp(-lambda (x - xMin)) (1 - fSig), .../example/ExSignalPlusBackgroundFit.scala:78-114.
// Only x is read (8 B per observation).  The log is taken ONCE per thread: the kernel keeps the running product of the
// pdf mantissas in a double and the sum of their exponents in an integer (log prod = sum log), which costs 1 DMUL and
// ~6 integer instructions per observation instead of a ~40-instruction FP64 log.  A pdf value that is zero, subnormal,
// inf or NaN takes the libdevice log on the spot (rare), so -2 log 0 = +inf propagates like in the reference.
// ---------------------------------------------------------------------------------------------------
struct PdfGaussExpBkg {
  static constexpr int NE = 2;
  struct Pre { double sr2, nmsr2, nlam, xmin, cS, cB; int zlim, xlim; };   // sr2, nmsr2, nlam scaled as in FamGaussExpBkg
  static Pre prepare(const double* p, const double* consts) {
    Pre c;
    const double fSig = p[0], mu = p[1], sigma = p[2], lambda = p[3], xMin = consts[0];
    const double s = 1.0 / sigma;
    c.sr2 = s * kInvSqrt2 * kExpScaleSqrt; c.nmsr2 = -mu * c.sr2;
    c.nlam = -lambda * kExpScale; c.xmin = xMin;
    c.zlim = lim_square(); c.xlim = lim_linear(c.nlam);
    c.cS = fSig * s * 0.3989422804014327;        // fSig / (sqrt(2 pi) sigma)
    c.cB = (1.0 - fSig) * lambda;
    return c;
  }
  __device__ static __forceinline__ void sargs(const Pre& c, double x, double* p, double* q) {
    const double zz = fma(x, c.sr2, c.nmsr2);
    p[0] = -zz; q[0] = zz;                       // -z^2/2
    p[1] = x - c.xmin; q[1] = c.nlam;            // -lambda (x - xMin)
  }
  __device__ static __forceinline__ bool in_range(const Pre& c, double, const double* p, const double* q) { return hi_below(q[0], c.zlim) && hi_below(p[1], c.xlim); }
  __device__ static __forceinline__ double value(const Pre& c, const double* ex) { return fma(c.cS, ex[0], c.cB * ex[1]); }
  // both exp arguments inside (-700, 700) for every x in [xmin, xmax] (end points decide); NaN compares false
  static bool args_in_range(const Pre& c, double xmin, double xmax) {
    const double z0 = std::fabs(xmin * c.sr2 + c.nmsr2), z1 = std::fabs(xmax * c.sr2 + c.nmsr2), zm = z0 > z1 ? z0 : z1;
    const double b0 = std::fabs((xmin - c.xmin) * c.nlam), b1 = std::fabs((xmax - c.xmin) * c.nlam);
    return zm * zm < kExpScaledLimit && b0 < kExpScaledLimit && b1 < kExpScaledLimit;
  }
};

template <class Pdf, int GROUP = 4>
struct OpNll {
  // acc[0] = product of pdf mantissas (renormalised), acc[1] = exponent sum (as an integer in the low word),
  // acc[2] = sum of logs taken on the slow path; finalize() folds them into acc[0] = sum log pdf
  static constexpr int Q = 3, NEXP = Pdf::NE, G = GROUP;
  static constexpr bool kHasSrc = false, kNeedsY = false;
  struct Args { typename Pdf::Pre pre; double scale[Q]; };
  __device__ static __forceinline__ void init(double (&acc)[Q]) { acc[0] = 1.0; acc[1] = 0.0; acc[2] = 0.0; }
  __device__ static __forceinline__ void sargs(const Args& a, double t, double* p, double* q) { Pdf::sargs(a.pre, t, p, q); }
  __device__ static __forceinline__ bool in_range(const Args& a, double t, const double* p, const double* q) { return Pdf::in_range(a.pre, t, p, q); }
  __device__ static __forceinline__ void accum(const Args& a, double, double, const double* ex, double (&acc)[Q]) {
    const double v = Pdf::value(a.pre, ex);
    const int hi = __double2hiint(v);
    const unsigned e = (unsigned)hi >> 20;                         // sign + exponent field
    if (e - 1u < 0x7feu) {                                         // positive normal number
      const double mant = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(v));
      acc[0] *= mant;
      acc[1] = __hiloint2double(0, __double2loint(acc[1]) + (int)e - 1023);
    } else {
      acc[2] += log(v);
    }
  }
  // keep the running product in [1, 2): called once per chunk (<= 8 multiplications by mantissas in [1, 2))
  __device__ static __forceinline__ void renorm(double (&acc)[Q]) {
    const int hi = __double2hiint(acc[0]);
    const int e = (hi >> 20) - 1023;
    acc[0] = __hiloint2double(hi - (e << 20), __double2loint(acc[0]));
    acc[1] = __hiloint2double(0, __double2loint(acc[1]) + e);
  }
  __device__ static __forceinline__ void finalize(double (&acc)[Q]) {
    acc[0] = fma((double)__double2loint(acc[1]), 0.6931471805599453, log(acc[0])) + acc[2];
    acc[1] = 0.0; acc[2] = 0.0;
  }
};

// G rows: all exp arguments first, ONE range check, then branch-free exponentials, then the accumulation.
// ---------------------------------------------------------------------------------------------------
// TSQR of the m x (n+1) matrix [J | r]: what QR.apply (linalg/QR.scala:131-250) computes with 2n+1
// passes over the data set, in ONE pass and without forming J^T J.  Every thread keeps a private upper-triangular
// R (C = n+1 columns, C(C+1)/2 registers) and folds B rows at a time into it with C Householder reflections of the
// stacked matrix [R; rows] — the reflector of column j is (R_jj - alpha; a_1j .. a_Bj) because everything else in
// that column is already zero.  No communication until the end; then triangles are merged pairwise (the QR of two
// stacked triangles) up a fixed tree: lanes, warps, blocks.  Householder on the rows themselves is backward stable
// whatever the conditioning of J; the normal-equations route (K2 + Cholesky) squares the condition number.
// Rows are the derivative BASIS rows (u_0 .. u_{n-1}, rho): [J | r] = +-[u | rho] diag(-D, 1), so R_A = R_U diag(-D, 1).
// ---------------------------------------------------------------------------------------------------
template <class Fam>
struct OpQr {
  static constexpr int N = Fam::N, C = N + 1, Q = C * (C + 1) / 2, NEXP = Fam::NE, G = C <= 6 ? 8 : 4;   // rows per fold
  struct Args { typename Fam::Pre pre; double scale[C]; };
  static constexpr bool kHasSrc = false, kNeedsY = true, kGroupFold = true;
  __device__ static __forceinline__ void sargs(const Args& a, double t, double* p, double* q) { Fam::sargs(a.pre, t, p, q); }
  __device__ static __forceinline__ bool in_range(const Args& a, double t, const double* p, const double* q) { return Fam::in_range(a.pre, t, p, q); }
  __device__ static __forceinline__ void row(const Args& a, double t, double y, const double* ex, double (&r)[C]) {
    double f, u[N];
    Fam::finish(a.pre, t, ex, f, u);
#pragma unroll
    for (int i = 0; i < N; ++i) r[i] = u[i];
    r[N] = Fam::residual(a.pre, y, ex, f);
  }
  template <int B>
  __device__ static __forceinline__ void accum_group(const Args& a, const double (&t)[B], const double (&y)[B],
                                                     const double* ex, double (&acc)[Q]) {   // ex: [B][NEXP]
    double rows[B][C];
#pragma unroll
    for (int i = 0; i < B; ++i) row(a, t[i], y[i], ex + i * NEXP, rows[i]);
    qr_fold<C, B>(acc, rows);
  }
  __device__ static __forceinline__ void accum(const Args& a, double t, double y, const double* ex, double (&acc)[Q]) {
    const double tt[1] = {t}, yy[1] = {y};
    accum_group<1>(a, tt, yy, ex, acc);
  }
};

template <class Op, int G, bool CHECK = true>
__device__ __forceinline__ void process_rows(const typename Op::Args& args, const double (&t)[G], const double (&y)[G],
                                             double (&acc)[Op::Q]) {
  constexpr int NE = Op::NEXP;
  if constexpr (NE == 0) {
    if constexpr (requires { Op::kGroupFold; }) Op::template accum_group<G>(args, t, y, nullptr, acc);
    else {
#pragma unroll
      for (int r = 0; r < G; ++r) Op::accum(args, t[r], y[r], nullptr, acc);
    }
  } else {
    double ea[G][NE], pp[G][NE], qq[G][NE];
    bool ok = true;
#pragma unroll
    for (int r = 0; r < G; ++r) {
      Op::sargs(args, t[r], pp[r], qq[r]);
      if (CHECK) ok = ok && Op::in_range(args, t[r], pp[r], qq[r]);
    }
    if (!CHECK || ok) {
#pragma unroll
      for (int r = 0; r < G; ++r)
#pragma unroll
        for (int e = 0; e < NE; ++e) ea[r][e] = so_exps(pp[r][e], qq[r][e]);
    } else {      // some |arg| >= 700, inf or NaN in this group: full-range libdevice exp for the whole group
#pragma unroll
      for (int r = 0; r < G; ++r)
#pragma unroll
        for (int e = 0; e < NE; ++e) ea[r][e] = exps_slow(pp[r][e], qq[r][e]);
    }
    if constexpr (requires { Op::kGroupFold; }) Op::template accum_group<G>(args, t, y, &ea[0][0], acc);
    else {
#pragma unroll
      for (int r = 0; r < G; ++r) Op::accum(args, t[r], y[r], ea[r], acc);
    }
  }
}

// a loop trip carries 4 rows (two 128-bit loads of t and of y); they are processed in groups of Op::G
template <class Op, bool CHECK = true>
__device__ __forceinline__ void process4(const typename Op::Args& args, double2 ta, double2 ya, double2 tb, double2 yb,
                                         double (&acc)[Op::Q]) {
  constexpr int G = Op::G;
  const double t4[4] = {ta.x, ta.y, tb.x, tb.y}, y4[4] = {ya.x, ya.y, yb.x, yb.y};
#pragma unroll
  for (int g0 = 0; g0 < 4; g0 += G) {
    double tg[G], yg[G];
#pragma unroll
    for (int r = 0; r < G; ++r) { tg[r] = t4[g0 + r]; yg[r] = y4[g0 + r]; }
    process_rows<Op, G, CHECK>(args, tg, yg, acc);
  }
}
template <class Op>
__device__ __forceinline__ void process1(const typename Op::Args& args, double t, double y, double (&acc)[Op::Q]) {
  const double tg[1] = {t}, yg[1] = {y};
  process_rows<Op, 1>(args, tg, yg, acc);
}

template <class Op>
__device__ __noinline__ void process1_cold(const typename Op::Args& args, double t, double y, double (&acc)[Op::Q]) {
  process1<Op>(args, t, y, acc);
}

// ---------------------------------------------------------------------------------------------------
// The kernel.  A block walks chunks of RPT * TPB observations (grid-stride, fixed assignment, so the summation
// order is fixed).  A chunk's t and y (2 x 8-16 KB at TPB = 256) are brought into a 3-4-deep shared-memory ring by
// TMA bulk copies (cp.async.bulk + mbarrier complete_tx, SASS UBLKCP) issued by thread 0: the loads cost no
// registers and no issue slots in the compute warps, and 64-96 KB per block are in flight whatever the occupancy.
// A stage is released (fence.proxy.async + block barrier) after its rows were processed and refilled by thread 0
// right after that barrier (see "Releasing a stage" in the loop).
// What ncu / A-B runs showed on the way here (profiles/r1_k2_*.json, profiles/r1_k2_ab*.log):
//   direct LDG.128                     FP64 pipe 63% busy, top stall long_scoreboard        r1_k2_v1_direct_ldg.json
//   ring + __syncthreads per chunk     FP64 pipe 70% busy, 9% of samples on the barrier     r1_k2_v2_block_ring.json
//   warp-private rings (1 KB copies)   slower than the block ring (54% vs 64% of roofline)  DESIGN.md
//   barrier right after the LDS        2.5% faster than releasing at the end, but races with the refill (MODE 4 note below)
// ---------------------------------------------------------------------------------------------------
__host__ __device__ constexpr int stages_for(int rpt) { return rpt >= 8 ? 3 : 4; }
constexpr size_t smem_bytes(int tpb, int rpt) { return (size_t)stages_for(rpt) * tpb * rpt * 16; }

template <class Op, int TPB, int MINBLOCKS, int MODE = 0, int RPT = 4, bool CHECK = true>
__global__ void __launch_bounds__(TPB, MINBLOCKS)
k_scalar(const double* __restrict__ t, const double* __restrict__ y, int64_t rows,
         const __grid_constant__ typename Op::Args args,
         double* __restrict__ partials, double* __restrict__ out, unsigned int* ticket) {
  constexpr int Q = Op::Q, WPB = TPB / 32, kChunkRows = TPB * RPT, kStages = stages_for(RPT);
  constexpr size_t kStageBytes = (size_t)kChunkRows * 16;
  extern __shared__ __align__(128) unsigned char s_ring[];      // [kStages][ t | y ]
  __shared__ double s_red[WPB * Q];
  __shared__ __align__(8) uint64_t s_full[kStages], s_empty[kStages];
  const int tid = threadIdx.x, lane = tid & 31;

  // a slice may start on an odd row: peel it so that every bulk copy source is 16-byte aligned
  const int head = (int)((reinterpret_cast<uintptr_t>(t) >> 3) & 1) && rows > 0 ? 1 : 0;
  const double* tb = t + head;
  const double* yb = y + head;
  const int64_t body = (rows - head) & ~(int64_t)1;             // even number of rows handled through the ring
  const int tail = (int)((rows - head) & 1);
  const int64_t nchunks = (body + kChunkRows - 1) / kChunkRows;

  auto issue = [&](int64_t chunk, int stage) {                  // one thread: arm the barrier, start both copies
    const int64_t r0 = chunk * kChunkRows;
    const uint32_t nr = (uint32_t)min((int64_t)kChunkRows, body - r0);
    double* st = reinterpret_cast<double*>(s_ring + (size_t)stage * kStageBytes);
    mbar_arrive_expect_tx(&s_full[stage], Op::kNeedsY ? nr * 16u : nr * 8u);
    bulk_g2s(st, tb + r0, nr * 8u, &s_full[stage]);
    if constexpr (Op::kNeedsY) bulk_g2s(st + kChunkRows, yb + r0, nr * 8u, &s_full[stage]);
  };

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kStages; ++s) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], WPB); }
    mbar_fence_init();
  }
  if constexpr (Op::NEXP > 0) exp_table_init();
  __syncthreads();
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kStages; ++s) {
      const int64_t c = (int64_t)blockIdx.x + (int64_t)s * gridDim.x;
      if (c < nchunks) issue(c, s);
    }
  }

  double acc[Q];
#pragma unroll
  for (int q = 0; q < Q; ++q) acc[q] = 0.0;
  if constexpr (requires { Op::init(acc); }) Op::init(acc);

  int it = 0;
  for (int64_t chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x, ++it) {
    const int stage = it % kStages;
    mbar_wait(&s_full[stage], (uint32_t)((it / kStages) & 1));
    const double* st = reinterpret_cast<const double*>(s_ring + (size_t)stage * kStageBytes);
    const int npairs = (int)(min((int64_t)kChunkRows, body - chunk * kChunkRows) >> 1);
    // thread tid owns pairs tid + j*TPB of the chunk: conflict-free LDS.128
    double2 tv[RPT / 2], yv[RPT / 2];
#pragma unroll
    for (int j = 0; j < RPT / 2; ++j) {
      tv[j] = reinterpret_cast<const double2*>(st)[tid + j * TPB];
      yv[j] = Op::kNeedsY ? reinterpret_cast<const double2*>(st + kChunkRows)[tid + j * TPB] : make_double2(0.0, 0.0);
    }
    // Releasing a stage.  The refill is an async-proxy (TMA) write to shared memory that this block has just read through
    // the generic proxy; the two must be ordered, and a block barrier alone does not do that: LDS instructions may still
    // be in flight when their warp passes BAR.SYNC.  (Seen for real: a variant of this kernel whose accumulators lived in
    // local memory returned run-to-run different sums at 1e9 rows with the barrier right after the loads; with the
    // proxy fence it was reproducible again.  tools/k2_ab.sh, profiles/r1_k2_ab_release_modes.log.)
    //   MODE 0 (default): process the rows, then fence.proxy.async + __syncthreads, then thread 0 refills the stage:
    //           the loads have long completed (their values were consumed), so the fence costs nothing.
    //   MODE 1: as MODE 0 with an `empty` mbarrier instead of the block barrier: only thread 0 waits for the other warps.
    //   MODE 4: fence + barrier right after the loads (maximum prefetch distance, but every warp waits for its LDS data).
    auto release_stage = [&]() {
#ifndef SO_NO_PROXY_FENCE
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#endif
      __syncthreads();
      if (tid == 0) {
        const int64_t nc = chunk + (int64_t)kStages * gridDim.x;
        if (nc < nchunks) issue(nc, stage);
      }
    };
    if (MODE == 4) release_stage();
    if (MODE == 2) {
      // every warp that passes this barrier has finished the previous iteration: refill the stage consumed there.  The
      // barrier sits where the warps wait for their LDS data anyway; prefetch distance kStages - 1 chunks.
      __syncthreads();
      if (tid == 0 && it >= 1) {
        const int64_t nc = chunk + (int64_t)(kStages - 1) * gridDim.x;
        if (nc < nchunks) issue(nc, (it - 1) % kStages);
      }
    }
    if constexpr (Op::G == 8 && RPT == 8) {
      if (tid + 3 * TPB < npairs) {                             // all 8 rows of this thread are there: one fold of 8
        const double t8[8] = {tv[0].x, tv[0].y, tv[1].x, tv[1].y, tv[2].x, tv[2].y, tv[3].x, tv[3].y};
        const double y8[8] = {yv[0].x, yv[0].y, yv[1].x, yv[1].y, yv[2].x, yv[2].y, yv[3].x, yv[3].y};
        process_rows<Op, 8, CHECK>(args, t8, y8, acc);
      } else {                                                  // ragged last chunk: one row at a time, out of line
#pragma unroll
        for (int j = 0; j < RPT / 2; ++j)
          if (tid + j * TPB < npairs) { process1_cold<Op>(args, tv[j].x, yv[j].x, acc); process1_cold<Op>(args, tv[j].y, yv[j].y, acc); }
      }
    } else
#pragma unroll
    for (int j = 0; j < RPT / 2; j += 2) {
      if (tid + (j + 1) * TPB < npairs) {
        process4<Op, CHECK>(args, tv[j], yv[j], tv[j + 1], yv[j + 1], acc);
      } else {
        if (tid + j * TPB < npairs) { process1<Op>(args, tv[j].x, yv[j].x, acc); process1<Op>(args, tv[j].y, yv[j].y, acc); }
      }
    }
    if constexpr (requires { Op::renorm(acc); }) Op::renorm(acc);
    if (MODE == 0) {
      release_stage();                                          // same sequence, but after the rows were consumed
    } else if (MODE == 2) {
#ifndef SO_NO_PROXY_FENCE
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // this iteration's reads, before the refill one barrier later
#endif
    } else if (MODE == 1) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[stage]);
      if (tid == 0) {
        const int64_t nc = chunk + (int64_t)kStages * gridDim.x;
        if (nc < nchunks) {
          mbar_wait(&s_empty[stage], (uint32_t)((it / kStages) & 1));
          issue(nc, stage);
        }
      }
    }
  }
  if constexpr (requires { Op::kGroupFold; }) {
    if (blockIdx.x == 0 && tid == 0 && head) process1_cold<Op>(args, t[0], y[0], acc);
    if (blockIdx.x == gridDim.x - 1 && tid == TPB - 1 && tail) process1_cold<Op>(args, t[rows - 1], y[rows - 1], acc);
  } else {
    if (blockIdx.x == 0 && tid == 0 && head) process1<Op>(args, t[0], Op::kNeedsY ? y[0] : 0.0, acc);
    if (blockIdx.x == gridDim.x - 1 && tid == TPB - 1 && tail) process1<Op>(args, t[rows - 1], Op::kNeedsY ? y[rows - 1] : 0.0, acc);
  }
  if constexpr (requires { Op::finalize(acc); }) { Op::renorm(acc); Op::finalize(acc); }

  if constexpr (requires { Op::kGroupFold; }) qr_grid_reduce<Op::C>(acc, s_red, partials, out, args.scale, ticket);
  else if constexpr (Op::kHasSrc) grid_reduce<Q>(acc, s_red, partials, out, args.scale, ticket, args.src);
  else if constexpr (requires { args.scale2; }) grid_reduce<Q>(acc, s_red, partials, out, args.scale, ticket, nullptr, Q / 2, args.scale2);
  else grid_reduce<Q>(acc, s_red, partials, out, args.scale, ticket);
}

template <class Op, int TPB, int MINBLOCKS, int MODE = 0, int RPT = 4, bool CHECK = true>
int launch_op_cfg(const Shard& s, const typename Op::Args& args) {
  auto kern = k_scalar<Op, TPB, MINBLOCKS, MODE, RPT, CHECK>;
  constexpr size_t smem = smem_bytes(TPB, RPT);
  static int per_sm_dev[64] = {0};   // per instantiation and per device: function attributes are per-device state
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  if (per_sm_dev[dev] == 0) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return SO_ECUDA;
    int v = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, kern, TPB, smem);
    per_sm_dev[dev] = v < 1 ? 1 : v;
  }
  const int per_sm = per_sm_dev[dev];
  constexpr int kChunkRows = TPB * RPT;
  const int64_t want = std::max<int64_t>(1, (s.rows + kChunkRows - 1) / kChunkRows);
  const int grid = (int)std::min<int64_t>(want, (int64_t)per_sm * s.sm_count);
  if ((size_t)grid * Op::Q > s.partials_doubles || (size_t)Op::Q > s.out_doubles) return SO_ENOMEM;
  kern<<<grid, TPB, smem, s.stream>>>(s.X, s.y, s.rows, args, s.partials, s.out, s.ticket);
  return cudaGetLastError() == cudaSuccess ? 1 : SO_ECUDA;
}

template <class Op, int MINBLOCKS>
int launch_op(const Shard& s, const typename Op::Args& args) { return launch_op_cfg<Op, kThreads, MINBLOCKS, 0, (MINBLOCKS <= 2 ? 8 : 4)>(s, args); }

// launch configurations of the kernels that are FP64-issue-bound for the two-exponential family; overridable for the
// A/B harness (tools/k2_ab.sh)
// (profiles/r1_k1_k4_ab_configs.log, 1e9 rows: K1 2 rows per range-checked group 2.76 ms vs 2.99 with 4, 2 blocks per SM
// and 8 rows per thread and chunk 2.71, check-free with 4 rows per group 2.56; K4 4 rows per group, 2 blocks per SM,
// 8 rows per thread and chunk 3.88 ms vs 4.26, check-free 3.63)
#ifndef SO_K1_G
#define SO_K1_G 2
#define SO_K1_MINB 2
#define SO_K1_RPT 8
#endif
#ifndef SO_K1_NC_G
#define SO_K1_NC_G 4
#endif
#ifndef SO_K3_G1
#define SO_K3_G1 4      // rows per range-checked group of the line kernels with 1 / 2 step sizes per pass
#define SO_K3_G2 2
#endif
#ifndef SO_K4_G
#define SO_K4_G 4
#define SO_K4_MINB 2
#define SO_K4_RPT 8
#endif
// families that can tell from the input range that no exp argument leaves the fast path's domain
template <class Fam> constexpr bool kFamHasRange = requires(const typename Fam::Pre& c) { Fam::args_in_range(c, 0.0, 0.0); };
template <class Fam> bool fam_in_range(const typename Fam::Pre& c, double lo, double hi) {
  if constexpr (kFamHasRange<Fam>) return Fam::args_in_range(c, lo, hi);
  else return false;
}
template <class Fam>
int launch_family(Kind kind, const Shard& s, const double* p, const double* d, const double* alpha, int k, int want_grad) {
  constexpr int N = Fam::N;
  const typename Fam::Pre pre = Fam::prepare(p);
  switch (kind) {
    case kLossGrad: {
      if (want_grad) {
        if constexpr (requires { Fam::args_in_range(pre, 0.0, 0.0); }) {
          if (s.has_range && Fam::args_in_range(pre, s.tmin, s.tmax)) {      // check-free variant, see kNormalEq
            using Opn = OpLossGrad<Fam, true, SO_K1_NC_G>;
            typename Opn::Args an; an.pre = pre;
            an.scale[0] = 0.5;
            for (int i = 0; i < N; ++i) an.scale[1 + i] = -pre.D[i];
            return launch_op_cfg<Opn, kThreads, SO_K1_MINB, 0, SO_K1_RPT, false>(s, an);
          }
        }
        using Op = OpLossGrad<Fam, true, SO_K1_G>;
        typename Op::Args a; a.pre = pre;
        a.scale[0] = 0.5;
        for (int i = 0; i < N; ++i) a.scale[1 + i] = -pre.D[i];
        return launch_op_cfg<Op, kThreads, SO_K1_MINB, 0, SO_K1_RPT>(s, a);
      }
      // loss only (trial points of Levenberg-Marquardt, line searches): the same two variants as with the gradient
      if constexpr (requires { Fam::args_in_range(pre, 0.0, 0.0); }) {
        if (s.has_range && Fam::args_in_range(pre, s.tmin, s.tmax)) {
          using Opn = OpLossGrad<Fam, false, SO_K1_NC_G>;
          typename Opn::Args an; an.pre = pre; an.scale[0] = 0.5;
          return launch_op_cfg<Opn, kThreads, SO_K1_MINB, 0, SO_K1_RPT, false>(s, an);
        }
      }
#ifndef SO_SCALAR_MINIMAL
      using Op = OpLossGrad<Fam, false>;
      typename Op::Args a; a.pre = pre; a.scale[0] = 0.5;
      return launch_op<Op, 4>(s, a);
#else
      return SO_EUNSUPPORTED;
#endif
    }
    case kNormalEq: {
      using Op = OpNormalEq<Fam>;
      typename Op::Args a; a.pre = pre;
      int q = 0;
      for (int i = 0; i < N; ++i) for (int j = i; j < N; ++j) a.scale[q++] = pre.D[i] * pre.D[j];
      for (int i = 0; i < N; ++i) a.scale[Op::T + i] = -pre.D[i];     // J = -sign(rho) df/dp, r = |rho|  =>  J r = -df/dp rho
      a.scale[Op::T + N] = 1.0;
      for (int q2 = 0; q2 < Op::Q; ++q2) a.src[q2] = (signed char)q2;
      for (int i = 0; i < N; ++i) for (int j = i; j < N; ++j) {
        int si = 0, sj = 0;
        if (Fam::gram_alias(i, j, si, sj)) a.src[Op::packed(i, j)] = (signed char)Op::packed(si, sj);
      }
      if constexpr (std::is_same<Fam, FamGaussExpBkg>::value) {
        // the headline kernel (BASELINE config 3): 2 rows per range-checked group, 8 rows per thread per chunk,
        // 2 blocks/SM.  A/B runs on one box (profiles/r1_k2_call26_groups.log, r1_k2_call27_groups.log): 4 rows per group -1 %, 4 rows per
        // chunk -3 %, 4096-entry exp table with a degree-3 polynomial +1.5 % (removed), no range check at all +1.5 %, 20-24 warps/SM -8..-25 %.
#ifndef SO_K2_G
#define SO_K2_G 2
#define SO_K2_TPB 256
#define SO_K2_MINB 2
#define SO_K2_MODE 0
#define SO_K2_RPT 8
#endif
        // every exp argument in range for this shard's inputs (known after launch_range): no per-row range check, and
        // then 4 rows per group are the better grouping (profiles/r1_k2_ab5_group4.log .. ab7: 3.22 vs 3.31 ms per 1e9 rows)
        if (s.has_range && Fam::args_in_range(pre, s.tmin, s.tmax)) {
#ifndef SO_K2NC_G
#define SO_K2NC_G 4
#define SO_K2NC_TPB 256
#define SO_K2NC_MINB 2
#define SO_K2NC_RPT 8
#endif
          using Op4 = OpNormalEq<Fam, SO_K2NC_G>;
          typename Op4::Args a4; a4.pre = a.pre;
          for (int q2 = 0; q2 < Op4::Q; ++q2) { a4.scale[q2] = a.scale[q2]; a4.src[q2] = a.src[q2]; }
          return launch_op_cfg<Op4, SO_K2NC_TPB, SO_K2NC_MINB, 0, SO_K2NC_RPT, false>(s, a4);
        }
        using Op2 = OpNormalEq<Fam, SO_K2_G>;
        typename Op2::Args a2; a2.pre = a.pre;
        for (int q2 = 0; q2 < Op2::Q; ++q2) { a2.scale[q2] = a.scale[q2]; a2.src[q2] = a.src[q2]; }
        return launch_op_cfg<Op2, SO_K2_TPB, SO_K2_MINB, SO_K2_MODE, SO_K2_RPT, true>(s, a2);
      }
      return launch_op<Op, (Op::Q < 21 ? 3 : 2)>(s, a);
    }
    case kLine: {
      // batches of up to 8 step sizes per pass; longer schedules take several passes into consecutive out slots
      // (the API layer asks for at most 8 per call for these families)
      if (k < 1 || k > 8) return SO_EINVAL;
      auto run = [&](auto KBtag) -> int {
        constexpr int KB = decltype(KBtag)::value;
        constexpr int GR = KB <= 1 ? SO_K3_G1 : (KB <= 2 ? SO_K3_G2 : 1);
        using Op = OpLine<Fam, KB, GR>;
        typename Op::Args a;
        double pk[N];
        bool in_range = s.has_range;
        for (int kk = 0; kk < KB; ++kk) {
          const double al = alpha[kk < k ? kk : k - 1];          // pad by repeating the last step size
          for (int i = 0; i < N; ++i) pk[i] = p[i] + d[i] * al;  // ptInit.x + ptInit.d * a1, StrongWolfe.scala:97
          a.pre[kk] = Fam::prepare(pk);
          for (int i = 0; i < N; ++i) a.dd[kk][i] = a.pre[kk].D[i] * d[i];
          a.scale[kk] = 0.5;
          a.scale[KB + kk] = -1.0;
          in_range = in_range && fam_in_range<Fam>(a.pre[kk], s.tmin, s.tmax);
        }
        constexpr int MB = KB >= 5 ? 1 : 2, RP = 8;
        if constexpr (kFamHasRange<Fam>) {
          if (in_range) return launch_op_cfg<Op, kThreads, MB, 0, RP, false>(s, a);     // check-free, see kNormalEq
        }
        return launch_op_cfg<Op, kThreads, MB, 0, RP>(s, a);
      };
      if (k == 1) return run(std::integral_constant<int, 1>());          // block sizes as in scalar_line_block()
      if (k == 2) return run(std::integral_constant<int, 2>());
      if (k == 3) return run(std::integral_constant<int, 3>());
      if (k == 4) return run(std::integral_constant<int, 4>());
      if (k == 5) return run(std::integral_constant<int, 5>());
      if (k == 6) return run(std::integral_constant<int, 6>());
      if (k == 7) return run(std::integral_constant<int, 7>());
      return run(std::integral_constant<int, 8>());
    }
    case kHessVec: {
      using Op = OpHessVec<Fam, SO_K4_G>;
      typename Op::Args a; a.pre = pre; a.h = Fam::hprepare(p, d);
      for (int i = 0; i < N; ++i) { a.dd[i] = pre.D[i] * d[i]; a.scale[i] = pre.D[i]; a.scale2[i] = -1.0; }
      if constexpr (requires { Fam::args_in_range(pre, 0.0, 0.0); }) {
        if (s.has_range && Fam::args_in_range(pre, s.tmin, s.tmax))
          return launch_op_cfg<Op, kThreads, SO_K4_MINB, 0, SO_K4_RPT, false>(s, a);
      }
      return launch_op_cfg<Op, kThreads, SO_K4_MINB, 0, SO_K4_RPT>(s, a);
    }
    case kQr: {
      using Op = OpQr<Fam>;
      typename Op::Args a; a.pre = pre;
      for (int i = 0; i < N; ++i) a.scale[i] = -pre.D[i];          // J = -sign(rho) df/dp, r = |rho|: rows +-(-D u | rho)
      a.scale[N] = 1.0;
      // C <= 6 columns (<= 170 registers): 3 blocks of 128 threads per SM = 12 warps instead of one block of 8
      // (n = 5 on 1e9 rows: 7.74 vs 8.53 ms; 2 x 192 threads 8.27, 4 x 96 8.19, 6 x 64 8.23); wider triangles need the
      // whole register file of one 256-thread block
#ifdef SO_QR_TPB
      return launch_op_cfg<Op, SO_QR_TPB, SO_QR_MINB, 0, 8>(s, a);
#else
      if constexpr (Op::C <= 6) return launch_op_cfg<Op, 128, 3, 0, 8>(s, a);
      else return launch_op_cfg<Op, 256, 1, 0, 8>(s, a);
#endif
    }
    default: break;        // kNll is served by launch_nll
  }
  return SO_EINVAL;
}

// ---------------------------------------------------------------------------------------------------
// Input range of an f == 1 shard: min, max and the number of non-finite values, as order-preserving 64-bit keys
// combined with integer atomics (exact and order-independent, unlike a floating-point reduction would need to be).
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long range_key(double x) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(x);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__global__ void __launch_bounds__(256) k_range(const double* __restrict__ t, int64_t rows, unsigned long long* __restrict__ buf3) {
  double lo = INFINITY, hi = -INFINITY;
  unsigned int bad = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < rows; i += (int64_t)gridDim.x * blockDim.x) {
    const double v = ld_stream(t + i);
    if (!(fabs(v) <= 1.7976931348623157e308)) ++bad;      // inf or NaN
    lo = fmin(lo, v); hi = fmax(hi, v);
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    bad += __shfl_xor_sync(0xffffffffu, bad, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(&buf3[0], range_key(lo));
    atomicMax(&buf3[1], range_key(hi));
    if (bad) atomicAdd(&buf3[2], (unsigned long long)bad);
  }
}

}  // namespace

int launch_range(const Shard& s, unsigned long long* buf3) {
  if (s.f != 1 || s.rows < 1) return SO_EINVAL;
  const int64_t want = (s.rows + 256 * 8 - 1) / (256 * 8);
  const int grid = (int)std::min<int64_t>(std::max<int64_t>(want, 1), (int64_t)s.sm_count * 8);
  k_range<<<grid, 256, 0, s.stream>>>(s.X, s.rows, buf3);
  return cudaGetLastError() == cudaSuccess ? 1 : SO_ECUDA;
}

bool range_decode(const unsigned long long* h, double* tmin, double* tmax) {
  if (h[2] != 0 || h[0] > h[1]) return false;
  auto dec = [](unsigned long long k) {
    const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    double d; std::memcpy(&d, &b, sizeof d); return d;
  };
  *tmin = dec(h[0]); *tmax = dec(h[1]);
  return std::isfinite(*tmin) && std::isfinite(*tmax);
}

size_t scalar_partials_needed(Kind kind, int n, int k, int sm_count) {
  // at most 8 resident 256-thread blocks per SM, Q <= 45 (polynomial n = 8 normal equations)
  const int Q = std::max(out_doubles(kind, n, 8), 16);
  (void)k;
  return (size_t)sm_count * 8 * (size_t)Q;
}

#ifndef SO_NLL_G
#define SO_NLL_G 2      // profiles/r1_nll_ab11_groups.log: 2 rows per checked group 2.36 ms vs 2.56 with 4; check-free with 4: 2.27 ms per 1e9 rows
#define SO_NLL_MINB 3
#define SO_NLL_RPT 4
#define SO_NLL_NC_G 4
#endif
int launch_nll(const Shard& s, int pdf, const double* p, int n, const double* consts, int nconsts) {
  if (s.f != 1) return SO_EINVAL;
  if (pdf != SO_PDF_GAUSS_EXPBKG) return SO_EUNSUPPORTED;
  if (n != 4 || nconsts < 1 || !consts) return SO_EINVAL;
  const PdfGaussExpBkg::Pre pre = PdfGaussExpBkg::prepare(p, consts);
  // FP64-issue-bound (two exponentials per 8-byte observation: 22 FP64 + ~20 other instructions/row)
  if (s.has_range && PdfGaussExpBkg::args_in_range(pre, s.tmin, s.tmax)) {       // check-free, see kNormalEq
    using Opn = OpNll<PdfGaussExpBkg, SO_NLL_NC_G>;
    Opn::Args an; an.pre = pre;
    an.scale[0] = -2.0; an.scale[1] = 0.0; an.scale[2] = 0.0;
    return launch_op_cfg<Opn, kThreads, SO_NLL_MINB, 0, SO_NLL_RPT, false>(s, an);
  }
  using Op = OpNll<PdfGaussExpBkg, SO_NLL_G>;
  Op::Args a;
  a.pre = pre;
  a.scale[0] = -2.0; a.scale[1] = 0.0; a.scale[2] = 0.0;       // sum - 2.0 * Math.log(pdf)
  return launch_op_cfg<Op, kThreads, SO_NLL_MINB, 0, SO_NLL_RPT>(s, a);
}

// padded width of the kLine result for the scalar families (so_api reads phi at [0,k) and dphi at [KB, KB+k))
int scalar_line_block(int k) { return k; }     // one instantiation per k = 1..8: no padded step sizes

int launch_scalar(Kind kind, const Shard& s, int family, const double* p, const double* d, int n,
                  const double* alpha, int k, int want_grad) {
  if (s.f != 1) return SO_EINVAL;
  switch (family) {
#ifndef SO_SCALAR_MINIMAL
    case SO_FAMILY_EXPDECAY: if (n != 2) return SO_EINVAL; return launch_family<FamExpDecay>(kind, s, p, d, alpha, k, want_grad);
    case SO_FAMILY_GAUSS: if (n != 3) return SO_EINVAL; return launch_family<FamGauss>(kind, s, p, d, alpha, k, want_grad);
#endif
    case SO_FAMILY_GAUSS_EXPBKG: if (n != 5) return SO_EINVAL; return launch_family<FamGaussExpBkg>(kind, s, p, d, alpha, k, want_grad);
#ifndef SO_SCALAR_MINIMAL
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
#endif
    default: return SO_EUNSUPPORTED;
  }
}

}  // namespace so
