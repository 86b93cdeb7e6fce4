// scalaopt_host.hpp — host-side mirror of the reference's interface for the data-set evaluation path.
//
// The reference is Scala and there is no JVM in this image, so the host side above the C ABI is C++.
// Same names, argument meaning and error behaviour as the reference (paths relative to
// core/src/main/scala/com/github/bruneli/scalaopt/core/):
//
//   DataSet<A>, SeqDataSet<A>                      DataSet.scala:27-145, SeqDataSetConverter.scala:28-61
//   AugmentedRow, QR                               linalg/AugmentedRow.scala:31-50, linalg/QR.scala:34-334
//   ObjectiveFunction .. MSEFunction               function/{ObjectiveFunction,DifferentiableObjectiveFunction,MSEFunction}.scala
//   ObjectiveFunctionWithGradient, ..FiniteDiff..  function/ObjectiveFunctionWithGradient.scala:29-49, ..FiniteDiffDerivatives.scala:30-52
//   LineSearchPoint                                variable/Point.scala:48-79
//   StrongWolfe                                    linesearch/StrongWolfe.scala:38-230
//   BFGS, ConjugateGradient, NewtonCG, SteihaugCG  gradient/*.scala
//   LevenbergMarquardt                             leastsquares/LevenbergMarquardt.scala:107-508
//
// The optimizers are the reference's control logic, restated statement by statement (including its quirks,
// SURVEY.md §3) with loops in place of tail recursion: they are the *callers* of the hot path and stay on
// the host.  What this repo adds is in gpu.hpp: GpuDataSet / GpuMSEFunction / GpuObjective, which answer the
// same calls from the sm_100a kernels behind include/scalaopt_cuda.h.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <functional>
#include <limits>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace scalaopt {

using Vec = std::vector<double>;

// ---- exceptions (core/MaxIterException.scala; require(...) -> IllegalArgumentException) ----
struct MaxIterException : std::runtime_error { using std::runtime_error::runtime_error; };
struct IllegalArgumentException : std::invalid_argument { using std::invalid_argument::invalid_argument; };
struct UnsupportedOperationException : std::logic_error { using std::logic_error::logic_error; };
inline void require(bool cond, const char* msg) { if (!cond) throw IllegalArgumentException(std::string("requirement failed: ") + msg); }

// ---- DenseVectorLike arithmetic (linalg/DenseVectorLike.scala:68-97, DenseVector.scala:64-219): left-to-right sums ----
inline double inner(const Vec& a, const Vec& b) { double acc = 0.0; for (size_t i = 0; i < a.size(); ++i) acc += a[i] * b[i]; return acc; }
inline double dot(const Vec& a, const Vec& b) { return inner(a, b); }
inline double norm2(const Vec& a) { double acc = 0.0; for (double v : a) acc += v * v; return acc; }
inline double norm(const Vec& a) { return std::sqrt(norm2(a)); }
inline Vec add(const Vec& a, const Vec& b) { Vec r(a.size()); for (size_t i = 0; i < a.size(); ++i) r[i] = a[i] + b[i]; return r; }
inline Vec sub(const Vec& a, const Vec& b) { Vec r(a.size()); for (size_t i = 0; i < a.size(); ++i) r[i] = a[i] - b[i]; return r; }
inline Vec mul(const Vec& a, double s) { Vec r(a.size()); for (size_t i = 0; i < a.size(); ++i) r[i] = a[i] * s; return r; }
inline Vec div(const Vec& a, double s) { Vec r(a.size()); for (size_t i = 0; i < a.size(); ++i) r[i] = a[i] / s; return r; }
inline Vec neg(const Vec& a) { Vec r(a.size()); for (size_t i = 0; i < a.size(); ++i) r[i] = -a[i]; return r; }
inline Vec zeros(size_t n) { return Vec(n, 0.0); }
// DenseVector.permute / unpermute (DenseVector.scala:315-339)
inline Vec permute(const std::vector<int>& ipvt, const Vec& x) { Vec r(x.size()); for (size_t j = 0; j < x.size(); ++j) r[ipvt[j]] = x[j]; return r; }
inline Vec unpermute(const std::vector<int>& ipvt, const Vec& x) { Vec r(x.size()); for (size_t j = 0; j < x.size(); ++j) r[j] = x[ipvt[j]]; return r; }

// ---- DataPoint (variable/Point.scala:29-37), scalar y ----
struct DataPoint { Vec x; double y = 0.0; };

// =====================================================================================================
// DataSet[A] (DataSet.scala:27-145) and the in-memory implementation (SeqDataSetConverter.scala:28-61)
// =====================================================================================================
template <class A>
class DataSet {
 public:
  virtual ~DataSet() = default;
  virtual std::shared_ptr<DataSet<A>> concat(const DataSet<A>& that) const = 0;                         // ++
  virtual void aggregate_raw(const std::function<void(const A&)>& seqop) const = 0;                     // fold order of the backend
  virtual std::vector<A> collect() const = 0;
  virtual std::shared_ptr<DataSet<A>> filter(const std::function<bool(const A&)>& p) const = 0;
  virtual void foreach (const std::function<void(const A&)>& f) const = 0;
  virtual A head() const = 0;
  virtual std::shared_ptr<DataSet<A>> map(const std::function<A(const A&)>& f) const = 0;
  virtual int64_t size() const = 0;
  // aggregate(z)(seqop, combop): a sequential backend folds left from z (SeqDataSetConverter.scala:32-35)
  template <class B>
  B aggregate(B z, const std::function<B(const B&, const A&)>& seqop, const std::function<B(const B&, const B&)>&) const {
    B acc = std::move(z);
    aggregate_raw([&](const A& a) { acc = seqop(acc, a); });
    return acc;
  }
};

template <class A>
class SeqDataSet : public DataSet<A> {
 public:
  std::vector<A> v;
  SeqDataSet() = default;
  explicit SeqDataSet(std::vector<A> values) : v(std::move(values)) {}
  std::shared_ptr<DataSet<A>> concat(const DataSet<A>& that) const override {
    std::vector<A> r = v; auto o = that.collect(); r.insert(r.end(), o.begin(), o.end());
    return std::make_shared<SeqDataSet<A>>(std::move(r));
  }
  void aggregate_raw(const std::function<void(const A&)>& seqop) const override { for (const A& a : v) seqop(a); }
  std::vector<A> collect() const override { return v; }
  std::shared_ptr<DataSet<A>> filter(const std::function<bool(const A&)>& p) const override {
    std::vector<A> r; for (const A& a : v) if (p(a)) r.push_back(a);
    return std::make_shared<SeqDataSet<A>>(std::move(r));
  }
  void foreach (const std::function<void(const A&)>& f) const override { for (const A& a : v) f(a); }
  A head() const override { if (v.empty()) throw std::out_of_range("head of empty list"); return v.front(); }
  std::shared_ptr<DataSet<A>> map(const std::function<A(const A&)>& f) const override {
    std::vector<A> r; r.reserve(v.size()); for (const A& a : v) r.push_back(f(a));
    return std::make_shared<SeqDataSet<A>>(std::move(r));
  }
  int64_t size() const override { return (int64_t)v.size(); }
};

// =====================================================================================================
// AugmentedRow + QR (linalg/AugmentedRow.scala:31-50, linalg/QR.scala)
// =====================================================================================================
struct AugmentedRow {
  Vec a; double b = 0.0; int64_t i = 0;
  AugmentedRow() = default;
  AugmentedRow(Vec a_, double b_, int64_t i_) : a(std::move(a_)), b(b_), i(i_) {}
  AugmentedRow operator+(const AugmentedRow& that) const { return AugmentedRow(add(that.a, a), that.b + b, i); }   // :33-35
  static AugmentedRow zeros(int n) { return AugmentedRow(Vec(n, 0.0), 0.0, 0); }
};

class QR {
 public:
  std::vector<Vec> r; Vec rDiag, qtb; std::vector<int> ipvt; Vec acNorms; double bNorm = 0.0; int n = 0;
  QR(std::vector<Vec> r_, Vec rDiag_, Vec qtb_, std::vector<int> ipvt_, Vec acNorms_, double bNorm_)
      : r(std::move(r_)), rDiag(std::move(rDiag_)), qtb(std::move(qtb_)), ipvt(std::move(ipvt_)),
        acNorms(std::move(acNorms_)), bNorm(bNorm_), n((int)rDiag.size()) {
    require(n > 0, "QR decomposition is empty");                 // QR.scala:45
    require(n == (int)r.size(), "R is not a square matrix");    // QR.scala:46
  }
  double R(int i, int j) const {                                 // QR.scala:55-67
    if (i < 0 || i >= n) throw IllegalArgumentException("Row out of bounds");
    if (j < 0 || j >= n) throw IllegalArgumentException("Column out of bounds");
    if (i == j) return rDiag[i];
    return i < j ? r[i][j] : 0.0;
  }
  Vec Rcol(int j) const {                                        // QR.scala:75-91
    if (j < 0 || j >= n) throw IllegalArgumentException("Column out of bounds");
    Vec c(n);
    for (int i = 0; i < n; ++i) c[i] = (i > j) ? 0.0 : (i == j ? rDiag[j] : r[i][j]);
    return c;
  }
  Vec solution() const {                                         // QR.scala:97-114
    int firstIdxZero = -1;
    for (int j = 0; j < n; ++j) if (rDiag[j] == 0.0) { firstIdxZero = j; break; }
    const int nsing = (firstIdxZero == -1) ? n - 1 : firstIdxZero;
    Vec x(n);
    for (int c = 0; c < n; ++c) x[c] = (c <= nsing) ? qtb[c] : 0.0;
    for (int j = nsing; j >= 0; --j) {
      double sum = 0.0;
      for (int k = j + 1; k < n; ++k) sum += x[k] * r[j][k];
      x[j] = (x[j] - sum) / rDiag[j];
    }
    return permute(ipvt, x);
  }

  // QR.apply (QR.scala:131-152): Householder QR expressed as 1 + n aggregates and n maps over the data set
  static QR apply(const DataSet<AugmentedRow>& ab_in, int n, bool pivoting = true) {
    using Row = AugmentedRow;
    std::shared_ptr<DataSet<Row>> ab = ab_in.map([](const Row& row) { return row; });
    // abNorms = doSqrt(ab.aggregate(zeros(n+1))(columnsNorm, _ + _))   :136-137,259-267
    Row sum = Row::zeros(n);      // AugmentedRow.zeros(n + 1) carries n+1 numbers: a[n] and b
    sum.a.assign(n, 0.0);
    ab->aggregate_raw([&](const Row& row) {
      Vec sq(row.a.size());
      for (size_t k = 0; k < row.a.size(); ++k) sq[k] = row.a[k] * row.a[k];
      sum = sum + Row(sq, row.b * row.b, row.i);
    });
    Vec acNorms(n);
    for (int k = 0; k < n; ++k) acNorms[k] = std::sqrt(k < (int)sum.a.size() ? sum.a[k] : 0.0);
    const double bNorm = std::sqrt(sum.b);
    std::vector<int> ipvt(n);
    for (int k = 0; k < n; ++k) ipvt[k] = k;
    Vec rDiag = acNorms;
    for (int j = 0; j < n; ++j) houseHolder(n, pivoting, rDiag, ab, ipvt, j);   // :143-145
    // reducedMat = ab.filter(_.i < n).collect().sortBy(_.i)   :147
    std::vector<Row> reduced = ab->filter([n](const Row& row) { return row.i < n; })->collect();
    std::stable_sort(reduced.begin(), reduced.end(), [](const Row& l, const Row& rr) { return l.i < rr.i; });
    std::vector<Vec> r; Vec qtb;
    for (const Row& row : reduced) { r.push_back(row.a); qtb.push_back(row.b); }
    return QR(std::move(r), std::move(rDiag), std::move(qtb), std::move(ipvt), std::move(acNorms), bNorm);
  }
  static Vec solve(const DataSet<AugmentedRow>& ab, int n) { return QR::apply(ab, n).solution(); }   // :161-164

 private:
  static void houseHolder(int n, bool pivoting, Vec& rDiag, std::shared_ptr<DataSet<AugmentedRow>>& ab,
                          std::vector<int>& ipvt, int j) {     // :169-250
    using Row = AugmentedRow;
    if (pivoting) {
      int kMax = j; double best = rDiag[j];                    // indexColumnLargestNorm :269-273
      for (int k = j; k < n; ++k) if (rDiag[k] > best) { best = rDiag[k]; kMax = k; }
      if (kMax != j) {                                         // swapColumns :278-285 (rDiag is NOT swapped)
        ab = ab->map([j, kMax](const Row& row) { Row o = row; std::swap(o.a[j], o.a[kMax]); return o; });
        std::swap(ipvt[j], ipvt[kMax]);
      }
    }
    // (dotProducts, rowj) = aggregate((zeros(n), zeros(n)))(columnsDotProduct(j, j), sumResults)   :190-192,297-312
    Row dots = Row::zeros(n), rowj = Row::zeros(n);
    ab->aggregate_raw([&](const Row& row) {
      if (row.i == j) {
        rowj = row;
      } else if (row.i > (int64_t)j) {
        const double xj = row.a[j];
        Vec t(row.a.size());
        for (size_t k = 0; k < row.a.size(); ++k) t[k] = ((int)k >= j) ? row.a[k] * xj : row.a[k];
        dots = Row(add(dots.a, t), dots.b + xj * row.b, row.i);
      }
    });
    const double rjj = rowj.a[j];
    const double sgn = (rjj > 0.0) ? 1.0 : ((rjj < 0.0) ? -1.0 : 0.0);
    const double ajNorm = std::sqrt(dots.a[j] + rjj * rjj) * sgn;            // :195-196
    const double ajj = (ajNorm != 0.0) ? rjj / ajNorm + 1.0 : 0.0;            // :198
    if (ajNorm == 0.0) { rDiag[j] = 0.0; return; }                           // :237-238
    auto updateRow = [&](double aij, double aik, int k) {                    // :201-211
      if (k == j) return aij;
      if (k > j) { const double alpha = dots.a[k] / ajNorm / ajj + rowj.a[k]; return aik - alpha * aij; }
      return aik;
    };
    Vec ajUpdatedRow(n);                                                     // :240-241
    for (int k = 0; k < n; ++k) ajUpdatedRow[k] = updateRow(ajj / ajNorm + 1.0, rowj.a[k], k);
    for (int k = j + 1; k < n; ++k) {                                        // rDiagUpdate :227-234
      const double akj = ajUpdatedRow[k], rr = rDiag[k];
      rDiag[k] = rr * std::sqrt(std::max(0.0, 1 - akj * akj / rr / rr));
    }
    rDiag[j] = -ajNorm;                                                      // :244
    ab = ab->map([&, j](const Row& row) {                                    // updateColumn :214-224
      if (row.i >= (int64_t)j) {
        const double aij = (row.i == (int64_t)j) ? ajj : row.a[j] / ajNorm;
        const double alpha = dots.b / ajNorm / ajj + rowj.b;
        const double qtb = row.b - alpha * aij;
        Vec a2(row.a.size());
        for (size_t k = 0; k < row.a.size(); ++k) a2[k] = updateRow(aij, row.a[k], (int)k);
        return Row(a2, qtb, row.i);
      }
      return row;
    });
  }
};

// =====================================================================================================
// Objective functions
// =====================================================================================================
struct ConfigPars {                                            // Optimizer.scala:56-63
  double tol = 1.0e-5; int maxIter = 200; double eps = 1.0e-8;
  ConfigPars() = default;
  ConfigPars(double tol_, int maxIter_, double eps_) : tol(tol_), maxIter(maxIter_), eps(eps_) {
    require(tol > 0.0, "tolerance error must be strictly positive.");
    require(maxIter > 0, "Maximum number of iterations must be > 0.");
    require(eps > 0.0, "epsilon must be strictly positive.");
  }
};

class DifferentiableObjectiveFunction {                        // DifferentiableObjectiveFunction.scala:28-60
 public:
  virtual ~DifferentiableObjectiveFunction() = default;
  virtual double apply(const Vec& x) const = 0;                // ObjectiveFunction.scala:38
  virtual Vec gradient(const Vec& x) const = 0;
  virtual double dirder(const Vec& x, const Vec& d) const = 0;
  virtual Vec dirHessian(const Vec& x, const Vec& d) const = 0;
};

// DenseVector.gradient (DenseVector.scala:269-275): forward differences
inline Vec fd_gradient(const std::function<double(const Vec&)>& f, const Vec& x, double eps = 1.0e-8) {
  const double fx = f(x);
  Vec g(x.size()), xp = x;
  for (size_t i = 0; i < x.size(); ++i) { xp[i] = x[i] + eps; g[i] = (f(xp) - fx) / eps; xp[i] = x[i]; }
  return g;
}

class ObjectiveFunctionWithGradient : public DifferentiableObjectiveFunction {   // ObjectiveFunctionWithGradient.scala:29-49
 public:
  std::function<double(const Vec&)> f; std::function<Vec(const Vec&)> df; ConfigPars config;
  ObjectiveFunctionWithGradient(std::function<double(const Vec&)> f_, std::function<Vec(const Vec&)> df_, ConfigPars c = ConfigPars())
      : f(std::move(f_)), df(std::move(df_)), config(c) {}
  double apply(const Vec& x) const override { return f(x); }
  Vec gradient(const Vec& x) const override { return df(x); }
  double dirder(const Vec& x, const Vec& d) const override { return dot(gradient(x), d); }
  Vec dirHessian(const Vec& x, const Vec& d) const override {
    const Vec gradx = gradient(x);
    const Vec gradxd = gradient(add(x, mul(d, config.eps)));
    return div(sub(gradxd, gradx), config.eps);
  }
};

class ObjectiveFunctionFiniteDiffDerivatives : public DifferentiableObjectiveFunction {   // ..FiniteDiffDerivatives.scala:30-52
 public:
  std::function<double(const Vec&)> f; ConfigPars config;
  explicit ObjectiveFunctionFiniteDiffDerivatives(std::function<double(const Vec&)> f_, ConfigPars c = ConfigPars())
      : f(std::move(f_)), config(c) {}
  double apply(const Vec& x) const override { return f(x); }
  Vec gradient(const Vec& x) const override { return fd_gradient(f, x, config.eps); }
  double dirder(const Vec& x, const Vec& d) const override { return (f(add(x, mul(d, config.eps))) - f(x)) / config.eps; }
  Vec dirHessian(const Vec& x, const Vec& d) const override {
    const Vec gradx = gradient(x);
    const Vec gradxd = gradient(add(x, mul(d, config.eps)));
    return div(sub(gradxd, gradx), config.eps);
  }
};

class MSEFunction : public DifferentiableObjectiveFunction {   // MSEFunction.scala:12-76
 public:
  // jacobianAndResidualsMatrix(p): the system LevenbergMarquardt hands to QR.apply (:74)
  virtual std::shared_ptr<DataSet<AugmentedRow>> jacobianAndResidualsMatrix(const Vec& p) const = 0;
};

// =====================================================================================================
// LineSearchPoint (variable/Point.scala:48-79): lazy vals; copy(...) makes a fresh instance
// =====================================================================================================
class LineSearchPoint {
 public:
  Vec x; const DifferentiableObjectiveFunction* f = nullptr; Vec d;
  LineSearchPoint() = default;
  LineSearchPoint(Vec x_, const DifferentiableObjectiveFunction* f_, Vec d_) : x(std::move(x_)), f(f_), d(std::move(d_)) {}
  LineSearchPoint copyX(Vec nx) const { return LineSearchPoint(std::move(nx), f, d); }
  LineSearchPoint copyD(Vec nd) const { return LineSearchPoint(x, f, std::move(nd)); }
  LineSearchPoint copyXD(Vec nx, Vec nd) const { return LineSearchPoint(std::move(nx), f, std::move(nd)); }
  double fx() const { if (!fx_) fx_ = f->apply(x); return *fx_; }                        // :54
  double dfx() const { if (!dfx_) dfx_ = f->dirder(x, d); return *dfx_; }                // :57
  const Vec& grad() const { if (!grad_) grad_ = f->gradient(x); return *grad_; }         // :60
  const Vec& d2fxd() const { if (!d2_) d2_ = f->dirHessian(x, d); return *d2_; }         // :63
  double m(const Vec& p, double eps = 1.0e-8) const {                                     // :71-78
    if (norm(p) < eps) return fx();
    const double pTBp = dot(p, f->dirHessian(x, p));
    return fx() + dot(grad(), p) + pTBp / 2.0;
  }
 private:
  mutable std::optional<double> fx_, dfx_;
  mutable std::optional<Vec> grad_, d2_;
};

// =====================================================================================================
// StrongWolfe (linesearch/StrongWolfe.scala)
// =====================================================================================================
struct StrongWolfeConfig {                                     // :48-56
  int maxIterLine = 10, maxIterZoom = 10; double c1 = 1.0e-4, c2 = 0.9, c3 = 2.0;
};

namespace StrongWolfe {

inline LineSearchPoint zoomStepLength(double a, const LineSearchPoint& pta, double b, const LineSearchPoint& ptb,
                                      const LineSearchPoint& ptInit, const StrongWolfeConfig& pars) {   // :132-230
  auto pol3Min = [](double a, double fa, double dfa, double b, double fb, double c, double fc) {        // :141-167
    require((a != b) && (a != c) && (b != c), "a, b and c must be 3 distinctive points");
    const double c0 = fa, c1 = dfa;
    const double m12 = (b - a) * (b - a), m11 = m12 * (b - a), m22 = (c - a) * (c - a), m21 = m22 * (c - a);
    const double mdet = m11 * m22 - m12 * m21;
    if (mdet == 0.0) return -1.0;
    const double dfb = fb - c1 * (b - a) - c0, dfc = fc - c1 * (c - a) - c0;
    const double c2 = (m11 * dfc - m21 * dfb) / mdet, c3 = (m22 * dfb - m12 * dfc) / mdet;
    const double delta = c2 * c2 - 3 * c1 * c3;
    if ((c3 == 0.0) || (delta < 0.0)) return -1.0;
    return a + (std::sqrt(delta) - c2) / (3 * c3);
  };
  auto pol2Min = [](double a, double fa, double dfa, double b, double fb) {                            // :170-180
    require(a != b, "a and b must be distinctive points");
    const double c0 = fa, c1 = dfa, d = b - a;
    const double c2 = (fb - c1 * d - c0) / (d * d);
    if (c2 <= 0.0) return -1.0;
    return a - c1 / (2.0 * c2);
  };
  const double f0 = ptInit.fx(), df0 = ptInit.dfx();           // :182-183
  int iter = 0;
  double a1 = a, f1 = pta.fx(), df1 = pta.dfx(), a2 = b, f2 = ptb.fx(), a3 = 0.0, f3 = 0.0;   // :229
  for (;;) {
    if (iter >= pars.maxIterZoom) throw MaxIterException("Maximum number of iterations reached.");      // :190-192
    auto outside = [&](double xNew) { return xNew < a1 + 0.1 * (a2 - a1) || xNew > a1 + 0.9 * (a2 - a1); };
    double aStar;                                             // findAlphaMin(3, 0.0)  :194-214
    {
      double xNew = (iter > 0) ? pol3Min(a1, f1, df1, a2, f2, a3, f3) : -1.0;
      if (outside(xNew)) {
        xNew = pol2Min(a1, f1, df1, a2, f2);
        if (outside(xNew)) xNew = 0.5 * (a2 - a1);            // :211 (reference: not a1 + ...)
      }
      aStar = xNew;
    }
    LineSearchPoint ptStar = ptInit.copyX(add(ptInit.x, mul(ptInit.d, aStar)));   // :215
    const double fStar = ptStar.fx(), dfStar = ptStar.dfx();
    if ((fStar <= f0 + pars.c1 * aStar * df0) && (std::fabs(dfStar) <= -pars.c2 * df0)) return ptStar;   // :219-220
    if ((fStar > f0 + pars.c1 * aStar * df0) || (fStar > f1) || (dfStar * (a2 - a1) >= 0.0)) {          // :221-224
      a3 = a2; f3 = f2; a2 = aStar; f2 = fStar;
    } else {                                                                                            // :226
      a3 = a1; f3 = f1; a1 = aStar; f1 = fStar; df1 = dfStar;
    }
    ++iter;
  }
}

inline LineSearchPoint stepLength(const LineSearchPoint& ptInit, const StrongWolfeConfig& pars) {        // :91-118
  const double fInit = ptInit.fx();
  int iter = 0;
  LineSearchPoint pt0 = ptInit;          // same instance semantics: shares what ptInit has already evaluated
  double a0 = 0.0, a1 = 1.0;
  for (;;) {
    LineSearchPoint pt1 = pt0.copyX(add(ptInit.x, mul(ptInit.d, a1)));     // :97
    const double f1 = pt1.fx(), df1 = pt1.dfx();                          // :98 (evaluated before the iteration cap)
    if (iter >= pars.maxIterLine) throw MaxIterException("Maximum number of iterations reached.");   // :99-101
    if ((f1 > fInit + pars.c1 * a1 * df1) || ((iter > 0) && (pt1.fx() > pt0.fx())))                  // :104-106
      return zoomStepLength(a0, pt0, a1, pt1, ptInit, pars);
    if (std::fabs(df1) <= -pars.c2 * pt0.dfx()) return pt1;                                          // :108
    if (df1 >= 0.0) return zoomStepLength(a0, pt0, a1, pt0, ptInit, pars);                           // :111-112
    pt0 = pt1; a0 = a1; a1 = pars.c3 * a1; ++iter;                                                   // :114
  }
}

}  // namespace StrongWolfe

// =====================================================================================================
// Gradient methods (gradient/*.scala).  minimize returns the minimiser; Failure(throw ...) = a thrown exception.
// =====================================================================================================
struct IterationTrace { int iterations = 0; };   // filled by every minimize: number of outer iterations performed

struct BFGSConfig : ConfigPars {                               // BFGS.scala:125-146
  int maxIterLine = 10, maxIterZoom = 10; std::optional<int> maxIterFirstTime; double c1 = 1.0e-4, c2 = 0.9, c3 = 2.0;
  StrongWolfeConfig getStrongWolfe(int it) const {
    StrongWolfeConfig s{maxIterLine, maxIterZoom, c1, c2, c3};
    if (it == 0 && maxIterFirstTime) s.maxIterZoom = *maxIterFirstTime;      // :139-144
    return s;
  }
};

namespace BFGS {
inline Vec minimize(const DifferentiableObjectiveFunction& f, const Vec& x0, const BFGSConfig& pars = BFGSConfig(),
                    IterationTrace* trace = nullptr) {          // BFGS.scala:57-110
  const size_t n = x0.size();
  using Mat = std::vector<Vec>;
  auto identity = [n]() { Mat m(n, Vec(n, 0.0)); for (size_t i = 0; i < n; ++i) m[i][i] = 1.0; return m; };
  auto matmul = [n](const Mat& A, const Mat& B) {               // RealMatrix.multiply: sum over k in order
    Mat C(n, Vec(n, 0.0));
    for (size_t i = 0; i < n; ++i) for (size_t j = 0; j < n; ++j) { double s = 0.0; for (size_t k = 0; k < n; ++k) s += A[i][k] * B[k][j]; C[i][j] = s; }
    return C;
  };
  auto updateHessian = [&](const Mat& iHk, const LineSearchPoint& ptk, const LineSearchPoint& ptkpp) {   // :69-80
    const Vec sk = sub(ptkpp.x, ptk.x);
    const Vec yk = sub(ptkpp.grad(), ptk.grad());
    const double yDots = dot(yk, sk);
    const double rhok = (yDots != 0.0) ? 1.0 / yDots : 1000.0;
    const Vec ykr = mul(yk, rhok), skr = mul(sk, rhok);          // sk outer yk * rhok == sk outer (yk * rhok)
    Mat m1 = identity(), m2 = identity(), out(n, Vec(n));
    for (size_t i = 0; i < n; ++i) for (size_t j = 0; j < n; ++j) { m1[i][j] -= sk[i] * ykr[j]; m2[i][j] -= yk[i] * skr[j]; }
    const Mat t = matmul(matmul(m1, iHk), m2);
    for (size_t i = 0; i < n; ++i) for (size_t j = 0; j < n; ++j) out[i][j] = t[i][j] + (sk[i] * sk[j]) * rhok;
    return out;
  };
  int k = 0;
  LineSearchPoint ptk(x0, &f, neg(f.gradient(x0)));              // :109
  Mat iHk = identity();
  for (;;) {
    if (trace) trace->iterations = k;
    if (k >= pars.maxIter) throw MaxIterException("Maximum number of iterations reached.");   // :87-89
    LineSearchPoint ptkpp = StrongWolfe::stepLength(ptk, pars.getStrongWolfe(k));             // :93
    if (norm(ptkpp.grad()) < pars.tol) { if (trace) trace->iterations = k + 1; return ptkpp.x; }   // :95-96
    Mat iHkpp = updateHessian(iHk, ptk, ptkpp);                  // :99
    Vec d(n);
    for (size_t i = 0; i < n; ++i) { double s = 0.0; for (size_t j = 0; j < n; ++j) s += iHkpp[i][j] * ptkpp.grad()[j]; d[i] = s * -1.0; }   // :100
    ptk = ptkpp.copyD(std::move(d));                             // :101
    iHk = std::move(iHkpp);
    ++k;
  }
}
}  // namespace BFGS

struct CGConfig : ConfigPars {                                 // ConjugateGradient.scala:119-131
  int maxIterLine = 10, maxIterZoom = 10; double c1 = 1.0e-4, c2 = 0.4, c3 = 2.0; std::string method = "PR+";
  StrongWolfeConfig strongWolfe() const { return StrongWolfeConfig{maxIterLine, maxIterZoom, c1, c2, c3}; }
};

namespace ConjugateGradient {
inline double beta(const Vec& dfk, const Vec& dfkpp, const std::string& method) {   // :92-102
  const double norm2dfk = dot(dfk, dfk);
  if (method == "FR") return dot(dfkpp, dfkpp) / norm2dfk;
  if (method == "PR") return dot(dfkpp, sub(dfkpp, dfk)) / norm2dfk;
  return std::max(0.0, dot(dfkpp, sub(dfkpp, dfk)) / norm2dfk);
}
inline Vec minimize(const DifferentiableObjectiveFunction& f, const Vec& x0, const CGConfig& pars = CGConfig(),
                    IterationTrace* trace = nullptr) {          // :53-90
  int k = 0;
  LineSearchPoint ptk(x0, &f, neg(f.gradient(x0)));
  for (;;) {
    if (trace) trace->iterations = k;
    if (k >= pars.maxIter) throw MaxIterException("Maximum number of iterations reached.");
    LineSearchPoint ptkpp = StrongWolfe::stepLength(ptk, pars.strongWolfe());
    const Vec dfk = ptk.grad();
    const Vec xkpp = ptkpp.x;
    const Vec dfkpp = ptkpp.grad();
    const Vec pkpp = add(neg(dfkpp), mul(ptk.d, beta(dfk, dfkpp, pars.method)));
    if (norm(dfkpp) < pars.tol) { if (trace) trace->iterations = k + 1; return xkpp; }
    ptk = ptkpp.copyD(pkpp);
    ++k;
  }
}
}  // namespace ConjugateGradient

namespace NewtonCG {
inline LineSearchPoint searchDirection(LineSearchPoint ptk, Vec zj, Vec rj, double epsk) {   // NewtonCG.scala:85-105
  for (int j = 0;; ++j) {
    if (dot(ptk.d, ptk.d2fxd()) <= 0.0) return (j == 0) ? ptk.copyD(neg(ptk.grad())) : ptk.copyD(zj);
    const double alphaj = dot(rj, rj) / dot(ptk.d, ptk.d2fxd());
    const Vec zjpp = add(zj, mul(ptk.d, alphaj));
    const Vec rjpp = add(rj, mul(ptk.d2fxd(), alphaj));
    if (norm(rjpp) < epsk) return ptk.copyD(zjpp);
    const double betajpp = dot(rjpp, rjpp) / dot(rj, rj);
    const Vec djpp = add(neg(rjpp), mul(ptk.d, betajpp));
    ptk = ptk.copyD(djpp); zj = zjpp; rj = rjpp;
  }
}
inline Vec minimize(const DifferentiableObjectiveFunction& f, const Vec& x0, const CGConfig& pars = CGConfig(),
                    IterationTrace* trace = nullptr) {          // :55-82
  int k = 0;
  LineSearchPoint ptk(x0, &f, zeros(x0.size()));
  for (;;) {
    if (trace) trace->iterations = k;
    if (k >= pars.maxIter) throw MaxIterException("Maximum number of iterations reached.");
    const double dfxNorm = norm(ptk.grad());
    const double epsk = std::min(0.5, std::sqrt(dfxNorm)) * dfxNorm;
    const Vec z0 = zeros(ptk.x.size());
    const Vec r0 = ptk.grad();
    const Vec d0 = neg(r0);
    LineSearchPoint ptkPrim = searchDirection(ptk.copyD(d0), z0, r0, epsk);
    LineSearchPoint ptkpp = StrongWolfe::stepLength(ptkPrim, pars.strongWolfe());
    if (norm(ptkpp.grad()) < pars.tol) { if (trace) trace->iterations = k + 1; return ptkpp.x; }
    ptk = ptkpp;
    ++k;
  }
}
}  // namespace NewtonCG

struct SteihaugCGConfig : ConfigPars {                         // SteihaugCG.scala:162-168
  double delta0 = 1.0, deltaMax = 1.0e5, eta = 0.2;
};

namespace SteihaugCG {
inline Vec pkFromTrustRegionEdge(const Vec& zj, const LineSearchPoint& ptk, double deltak) {   // :131-148
  const Vec& dj = ptk.d;
  const double a = dot(dj, dj), b = dot(zj, dj), c = dot(zj, zj) - deltak * deltak;
  const double disc = b * b - a * c;
  const double tauPos = (-b + std::sqrt(disc)) / a, tauNeg = (-b - std::sqrt(disc)) / a;
  const Vec pkPos = add(zj, mul(ptk.d, tauPos)), pkNeg = add(zj, mul(ptk.d, tauNeg));
  (void)ptk.m(pkPos); (void)ptk.m(pkNeg); (void)ptk.fx();      // val mkPos, mkNeg, m0 (:144-146): evaluated, unused
  return (ptk.m(pkPos) < ptk.m(pkNeg)) ? pkPos : pkNeg;        // :147
}
inline LineSearchPoint searchDirection(LineSearchPoint ptk, Vec zj, Vec rj, double deltak, double epsk) {   // :94-121
  for (int j = 0;; ++j) {
    if (dot(ptk.d, ptk.d2fxd()) <= 0.0) {
      const Vec pkpp = pkFromTrustRegionEdge(zj, ptk, deltak);
      return ptk.copyXD(add(ptk.x, pkpp), pkpp);
    }
    const double alphaj = dot(rj, rj) / dot(ptk.d, ptk.d2fxd());
    const Vec zjpp = add(zj, mul(ptk.d, alphaj));
    if (norm(zjpp) > deltak) {
      const Vec pkpp = pkFromTrustRegionEdge(zj, ptk, deltak);
      return ptk.copyXD(add(ptk.x, pkpp), pkpp);
    }
    const Vec rjpp = add(rj, mul(ptk.d2fxd(), alphaj));
    if (norm(rjpp) < epsk) return ptk.copyXD(add(ptk.x, zjpp), zjpp);
    const double betajpp = dot(rjpp, rjpp) / dot(rj, rj);
    const Vec djpp = add(neg(rjpp), mul(ptk.d, betajpp));
    ptk = ptk.copyD(djpp); zj = zjpp; rj = rjpp;
    std::swap(deltak, epsk);                                   // :117 — the reference recurses with (epsk, deltak)
  }
}
inline Vec minimize(const DifferentiableObjectiveFunction& f, const Vec& x0,
                    const SteihaugCGConfig& pars = SteihaugCGConfig(), IterationTrace* trace = nullptr) {   // :51-91
  int k = 0;
  double deltak = pars.delta0;
  LineSearchPoint ptk(x0, &f, zeros(x0.size()));
  for (;;) {
    if (trace) trace->iterations = k;
    if (k >= pars.maxIter) throw MaxIterException("Maximum number of iterations reached.");
    const double dfxNorm = norm(ptk.grad());
    const double epsk = std::min(0.5, std::sqrt(dfxNorm)) * dfxNorm;
    const Vec z0 = zeros(ptk.x.size());
    const Vec r0 = ptk.grad();
    const Vec d0 = neg(r0);
    LineSearchPoint ptkpp = searchDirection(ptk.copyD(d0), z0, r0, deltak, epsk);
    const double actualReduction = ptk.fx() - ptkpp.fx();
    const double predictedReduction = ptk.fx() - ptk.m(ptkpp.d);
    const double rhok = actualReduction / predictedReduction;
    double deltakpp;
    if (rhok < 0.25) deltakpp = deltak / 4.0;
    else if (rhok > 3.0 / 4.0 && std::fabs(norm(ptkpp.d) - deltak) < pars.eps) deltakpp = std::min(2.0 * deltak, pars.deltaMax);
    else deltakpp = deltak;
    if (norm(ptkpp.grad()) < pars.tol) { if (trace) trace->iterations = k + 1; return ptkpp.x; }
    if (rhok > pars.eta) ptk = ptkpp;
    deltak = deltakpp;
    ++k;
  }
}
}  // namespace SteihaugCG


// =====================================================================================================
// NelderMead (derivativefree/NelderMead.scala:40-314): the default method of maxLikelihoodFit
// (std-apps/.../stats/package.scala:74-87).  Consumes only f(x); every Vertex evaluates f once.
// =====================================================================================================
struct NelderMeadConfig : ConfigPars {                          // :307-314
  double cReflection = 2.0, cExpansion = 1.0, cContraction = 0.5, cShrink = 0.5, relDelta = 0.05, absDelta = 0.00025;
};

namespace NelderMead {
struct Vertex { Vec x; double fx; };
inline Vec minimize(const DifferentiableObjectiveFunction& f, const Vec& x0, const NelderMeadConfig& pars = NelderMeadConfig(),
                    IterationTrace* trace = nullptr) {
  const int n = (int)x0.size();
  auto vertex = [&](Vec x) { Vertex v; v.fx = f.apply(x); v.x = std::move(x); return v; };
  auto barycenterOf = [&](const std::vector<Vertex>& vs) {      // Simplex.barycenter :286-291
    Vec r = div(vs.front().x, (double)(n + 1));
    for (size_t i = 1; i < vs.size(); ++i) r = add(r, div(vs[i].x, n + 1.0));
    return r;
  };
  auto byValue = [](const Vertex& a, const Vertex& b) { return a.fx < b.fx; };
  // Simplex.apply :272-283
  std::vector<Vertex> vertices;
  vertices.push_back(vertex(x0));
  for (int i = 0; i < n; ++i) {                                 // Vertex.shift :73-84
    Vec x = x0;
    x[i] = (x0[i] == 0.0) ? x0[i] + pars.absDelta : x0[i] * pars.relDelta;
    vertices.push_back(vertex(std::move(x)));
  }
  std::stable_sort(vertices.begin(), vertices.end(), byValue);
  require(vertices.size() > 1, "There must be at least 2 vertices");
  Vec barycenter = barycenterOf(vertices);
  int iter = 0;
  for (;;) {
    if (trace) trace->iterations = iter;
    const Vertex& vtxMin = vertices.front();
    const Vertex vtxMax = vertices.back();
    // barycenter(vtxMax) - vtxMax.x  :125-131
    const Vec move = sub(sub(mul(barycenter, 1.0 + 1.0 / n), div(vtxMax.x, (double)n)), vtxMax.x);
    if (norm(move) / std::sqrt((double)vertices.size()) < pars.tol) return vtxMin.x;      // stoppingRule :54-57
    if (iter >= pars.maxIter * (int)vertices.size()) throw MaxIterException("Maximum number of iterations reached.");   // :233-236
    auto update = [&](const Vertex& vtxNew) {                   // :245-259
      barycenter = add(barycenter, div(sub(vtxNew.x, vtxMax.x), (double)(n + 1)));
      std::vector<Vertex> init(vertices.begin(), vertices.end() - 1), out;
      if (vtxNew.fx < vertices.front().fx) { out.push_back(vtxNew); out.insert(out.end(), init.begin(), init.end()); }
      else if (vtxNew.fx > init.back().fx) { out = init; out.push_back(vtxNew); }
      else {
        size_t index = 0;
        while (index < vertices.size() && !(vertices[index].fx >= vtxNew.fx)) ++index;
        out.assign(vertices.begin(), vertices.begin() + index);
        out.push_back(vtxNew);
        if (index < init.size()) out.insert(out.end(), init.begin() + index, init.end());
      }
      vertices = std::move(out);
    };
    bool moved = false;
    {                                                           // reflection :138-150
      const Vertex vtxNew = vertex(add(vtxMax.x, mul(move, pars.cReflection)));
      if (vtxNew.fx <= vtxMin.fx) {                             // expansion :160-167
        const Vertex vtxExp = vertex(add(vtxNew.x, mul(move, pars.cExpansion)));
        update(vtxExp.fx <= vtxMin.fx ? vtxExp : vtxNew);
        moved = true;
      } else if (vtxNew.fx <= vtxMax.fx) { update(vtxNew); moved = true; }
    }
    if (!moved) {                                               // contraction :178-185
      const Vertex vtxNew = vertex(add(vtxMax.x, mul(move, pars.cContraction)));
      if (vtxNew.fx <= vtxMax.fx) { update(vtxNew); moved = true; }
    }
    if (!moved) {                                               // shrinkage :196-212
      const Vec xmin = vtxMin.x;
      std::vector<Vertex> shrunk;
      for (const Vertex& v : vertices) shrunk.push_back(vertex(add(xmin, mul(sub(v.x, xmin), pars.cShrink))));
      std::stable_sort(shrunk.begin(), shrunk.end(), byValue);
      vertices = std::move(shrunk);
      barycenter = barycenterOf(vertices);
    }
    ++iter;
  }
}
}  // namespace NelderMead

// =====================================================================================================
// LevenbergMarquardt (leastsquares/LevenbergMarquardt.scala:107-508): MINPACK lmder/lmpar/qrsolv as the
// reference restates them, quirks included
// =====================================================================================================
struct LevenbergMarquardtConfig {                              // :497-508
  double tol = 1.49012e-8; int maxIter = 10; double eps = 1.0e-8; double xTol = 1.49012e-8; double gTol = 0.0;
  double ratioTol = 0.0001; double stepBound = 100.0; double epsmch = 2.22044604926e-16;
  int maxInnerIter = 10; int maxStepLengthIter = 10; bool usePivoting = false;
};

namespace LevenbergMarquardt {

struct Trace { int outerIterations = 0; int innerIterations = 0; int gaussNewtonSteps = 0; int dampedSteps = 0; };

inline std::pair<Vec, Vec> qrSolve(const QR& qr, const Vec& diag) {                 // :398-476
  const int n = (int)diag.size();
  std::vector<Vec> s(n);
  for (int j = 0; j < n; ++j) s[j] = qr.Rcol(j);                                    // :453 (columns of R)
  Vec sDiag = zeros(n), qtb = qr.qtb;
  for (int j = 0; j < n; ++j) {                                                     // diagonalElimination :403-419
    const int jpvt = qr.ipvt[j];
    if (diag[jpvt] == 0.0) {
      const double old = s[j][j];
      s[j][j] = qr.R(j, j);
      sDiag[j] = old;                                                               // s0(j)(j) before the update
    } else {
      Vec sDiagInit = sDiag;
      sDiagInit[j] = diag[jpvt];
      for (int i = j + 1; i < n; ++i) sDiagInit[i] = 0.0;                           // :412-414
      Vec sj = s[j], sD = sDiagInit, q = qtb; double qtbpj = 0.0;
      for (int k = j; k < n; ++k) {                                                 // rotation :421-448
        if (sD[k] == 0.0) continue;
        double cosv, sinv;
        if (std::fabs(sj[k]) < std::fabs(sD[k])) {
          const double cot = sj[k] / sD[k];
          sinv = 1.0 / std::sqrt(1.0 + cot * cot); cosv = sinv * cot;
        } else {
          const double tanv = sD[k] / sj[k];
          cosv = 1.0 / std::sqrt(1.0 + tanv * tanv); sinv = cosv * tanv;
        }
        Vec sj2 = sj, sD2 = sD;
        for (int i = k; i < n; ++i) sj2[i] = cosv * sj[i] + sinv * sD[i];                              // :438-440
        for (int i = k + 1; i < n; ++i) sD2[i] = -sinv * sD[i] + cosv * sD[i];                         // :441-443: the lambda parameter
                                                                                    // shadows the outer s0, so both operands are sDiag0(i)
        const double qk = q[k];
        q[k] = cosv * qk + sinv * qtbpj;                                            // :444
        qtbpj = -sinv * qk + cosv * qtbpj;                                          // :445
        sj = sj2; sD = sD2;
      }
      // (s0.updated(j, sj.updated(j, R(j,j))), sDiag.updated(j, sj(j)), qtb)        :417
      const double sjj = sj[j];
      sj[j] = qr.R(j, j);
      s[j] = sj;
      sDiag = sD;
      sDiag[j] = sjj;
      qtb = q;
    }
  }
  int firstIdxZero = -1;
  for (int j = 0; j < n; ++j) if (sDiag[j] == 0.0) { firstIdxZero = j; break; }
  const int nsing = (firstIdxZero == -1) ? n - 1 : firstIdxZero - 1;                // :460-461
  Vec x(n);
  for (int c = 0; c < n; ++c) x[c] = (c <= nsing) ? qtb[c] : 0.0;
  for (int j = nsing; j >= 0; --j) {                                                // solve :465-473
    double sum = 0.0;
    for (int k = j + 1; k < n; ++k) sum += x[k] * s[j][k];
    x[j] = (x[j] - sum) / sDiag[j];
  }
  return {permute(qr.ipvt, x), sDiag};
}


// Determine the Levenberg-Marquardt parameter (:250-369).  Returns (par, step, scaled step).
struct LmParResult { double par; Vec x; Vec xDiag; bool gaussNewton; };
inline LmParResult lmPar(const QR& qr, const Vec& diag, double delta, double par0, int maxStepLengthIter) {
  const int n = (int)diag.size();
  const double p1 = 0.1;
  // gaussNewtonDirection :259-283
  int firstIdxZero = -1;
  for (int j = 0; j < n; ++j) if (qr.rDiag[j] == 0.0) { firstIdxZero = j; break; }
  const int nsing = (firstIdxZero == -1) ? n - 1 : firstIdxZero - 1;
  Vec xu(n);
  for (int c = 0; c < n; ++c) xu[c] = (c <= nsing) ? qr.qtb[c] : 0.0;
  for (int j = nsing; j >= 0; --j) {                             // qtbMinusR :267-279
    if (xu[j] == 0.0) break;                                     // stops at the first exactly-zero component
    const double alpha = xu[j] / qr.rDiag[j];
    const Vec rj = qr.Rcol(j);
    xu[j] = alpha;
    for (int i = 0; i < j; ++i) xu[i] = xu[i] - rj[i] * alpha;
  }
  const Vec x = permute(qr.ipvt, xu);
  Vec xdi(n);
  for (int i = 0; i < n; ++i) xdi[i] = diag[i] * x[i];
  const double dxNorm = norm(xdi);
  const double fp = dxNorm - delta;
  if (fp <= p1 * delta) return {par0, x, xdi, true};             // :290-291

  const Vec unpermutedDiag = unpermute(qr.ipvt, diag);
  double parLow = 0.0;
  if (nsing >= n) {                                              // :294-311 (unreachable: nsing <= n-1; kept verbatim)
    const Vec unpermutedXdi = unpermute(qr.ipvt, xdi);
    Vec aux(n);
    for (int i = 0; i < n; ++i) aux[i] = unpermutedDiag[i] * unpermutedXdi[i] / dxNorm;
    for (int j = 0; j < n; ++j) {
      const Vec rj = qr.Rcol(j);
      double sum = 0.0;
      for (int i = 0; i < j - 1; ++i) sum += aux[i] * rj[i];
      aux[j] = (aux[j] - sum) / qr.rDiag[j];
    }
    parLow = fp / delta / inner(aux, aux);
  }
  Vec aux(n);
  for (int i = 0; i < n; ++i) {                                  // :312-314 (sums exclude the diagonal: take(i))
    const Vec ri = qr.Rcol(i);
    double sum = 0.0;
    for (int k = 0; k < i; ++k) sum += ri[k] * qr.qtb[k];
    aux[i] = sum / unpermutedDiag[i];
  }
  const double parUp = norm(aux) / delta;

  double lo = parLow, cur = std::min(std::max(par0, parLow), parUp), up = parUp;   // parI :366
  double fp0 = fp;
  for (int iter = 0;; ++iter) {                                  // stepLength :317-365
    auto sol = qrSolve(qr, diag);                                // called WITHOUT the LM parameter (:322)
    const Vec& xs = sol.first; const Vec& sDiag = sol.second;
    Vec xDiag(n);
    for (int i = 0; i < n; ++i) xDiag[i] = xs[i] * sDiag[i];
    const double dxN = norm(xDiag);
    const double fpi = dxN - delta;
    if ((std::fabs(fpi) <= p1 * delta) || (lo == 0.0 && fpi <= fp0 && fp0 < 0.0) || (iter == maxStepLengthIter))
      return {cur, xs, xDiag, false};
    Vec pivotedAux0(n);
    for (int i = 0; i < n; ++i) pivotedAux0[i] = diag[i] * xDiag[i] / dxN;
    Vec a2 = unpermute(qr.ipvt, pivotedAux0);
    for (int j = 0; j < n; ++j) {                                // multiply :333-352
      const double auxj = a2[j] / sDiag[j];
      const Vec rcol = qr.Rcol(j);
      double rUpperNorm2 = 0.0;
      for (int k = j + 1; k < n; ++k) rUpperNorm2 += rcol[k] * rcol[k];   // sub-diagonal part of a column: zero
      Vec upd = a2;
      for (int i = 0; i < n; ++i) {
        const double xv = a2[i];
        if (i == j) upd[i] = auxj;
        else if (xv > j && rUpperNorm2 != 0.0) {                 // compares the VALUE with j (reference quirk)
          const int idx = i - j - 1;
          if (idx < 0 || idx >= n - j - 1) throw std::out_of_range("rUpper index");
          upd[i] = xv - rcol[j + 1 + idx] / auxj;
        }
      }
      a2 = upd;
    }
    const double parc = fpi / delta / inner(a2, a2);
    if (fpi > 0.0) {
      const double pl = std::max(lo, cur);
      const double nc = std::max(pl, cur + parc);
      lo = pl; cur = nc;                                         // (parLow', max(parLow', cur + parc), up)
    } else {
      const double nc = std::max(parLow, cur + parc);
      const double nu = std::min(cur, up);
      cur = nc; up = nu;                                         // (lo, max(parLow, cur + parc), min(cur, up))
    }
    fp0 = fpi;
  }
}

inline Vec minimize(const MSEFunction& f, const Vec& x0in, const LevenbergMarquardtConfig& pars = LevenbergMarquardtConfig(),
                    Trace* trace = nullptr) {                    // :111-226; always Success (:225)
  const int n = (int)x0in.size();
  Vec x0 = x0in, diag0 = zeros(n);
  double xNorm0 = 0.0, delta0 = 0.0;
  for (int outer = 0;; ++outer) {                                // iterate :118-150
    if (trace) trace->outerIterations = outer + 1;
    auto ab = f.jacobianAndResidualsMatrix(x0);
    const QR qr = QR::apply(*ab, n, pars.usePivoting);           // :124
    const double fNorm0 = qr.bNorm;
    Vec diag(n); double xNorm, delta1;                           // updateDiagNormDelta :228-247
    if (outer == 0) {
      for (int i = 0; i < n; ++i) diag[i] = (qr.acNorms[i] > 0.0) ? qr.acNorms[i] : 1.0;
      xNorm = std::sqrt(dot(diag, x0));                          // sqrt(diag dot x) (reference: not ||diag o x||)
      delta1 = (xNorm == 0.0) ? pars.stepBound : pars.stepBound * xNorm;
    } else {
      for (int i = 0; i < n; ++i) diag[i] = std::max(diag0[i], qr.acNorms[i]);
      xNorm = xNorm0; delta1 = delta0;
    }
    (void)xNorm;
    double gNorm = 0.0;                                          // scaledGradientNorm :130-141
    for (int j = 0; j < n; ++j) {
      const double acNormj = qr.acNorms[qr.ipvt[j]];
      if (acNormj == 0.0) continue;
      double sum = 0.0;
      for (int i = 0; i < j; ++i) sum = sum + qr.R(i, j) * qr.qtb[i] / fNorm0;
      gNorm = std::max(gNorm, std::fabs(sum / acNormj));
    }
    // innerLoop :152-223
    Vec xin = x0, dg = diag; double dl = delta1, par0 = 0.0, fN0 = fNorm0;
    Vec x; double fNorm = 0.0, delta2 = 0.0; bool stoppingRule = false;
    for (int inner = 0;; ++inner) {
      if (trace) trace->innerIterations += 1;
      const LmParResult lp = lmPar(qr, dg, dl, par0, pars.maxStepLengthIter);
      if (trace) { if (lp.gaussNewton) trace->gaussNewtonSteps += 1; else trace->dampedSteps += 1; }
      const double par = lp.par; const Vec& pk = lp.x; const Vec& xDiag = lp.xDiag;
      const double xDiagNorm = norm(xDiag);
      x = sub(xin, pk);
      const double xN = norm(x);
      fNorm = std::sqrt(f.apply(x));                             // :164
      Vec predicted = zeros(dg.size());
      for (int j = 0; j < (int)predicted.size(); ++j) predicted = sub(predicted, mul(qr.Rcol(j), pk[qr.ipvt[j]]));   // :167-175
      const double scaledPredictedNorm2 = norm2(predicted) / (fN0 * fN0);
      const double scaledxDiagNorm2 = par * (xDiagNorm * xDiagNorm) / (fN0 * fN0);
      const double predictedReduction = scaledPredictedNorm2 + 2.0 * scaledxDiagNorm2;
      const double directionalDerivative = -(scaledPredictedNorm2 + scaledxDiagNorm2);
      const double actualReduction = (fNorm < fN0) ? 1.0 - (fNorm * fNorm) / (fN0 * fN0) : -1.0;      // :182
      const double ratio = (predictedReduction != 0.0) ? actualReduction / predictedReduction : 0.0;
      double parNew, delta;
      if (ratio > 0.25) {
        if (par != 0.0 && ratio < 0.75) { parNew = par; delta = dl; } else { parNew = 0.5 * par; delta = 2.0 * xDiagNorm; }
      } else {
        const double temp1 = (actualReduction >= 0.0) ? 0.5 : 0.5 * directionalDerivative / (directionalDerivative + 0.5 * actualReduction);
        const double temp2 = (0.1 * fNorm >= fN0 || temp1 < 0.1) ? 0.1 : temp1;
        parNew = par / temp2; delta = temp2 * std::min(dl, 10.0 * xDiagNorm);
      }
      const bool convergence1 = std::fabs(actualReduction) <= pars.tol && predictedReduction <= pars.tol && 0.5 * ratio <= 1.0;
      const bool convergence2 = delta <= pars.xTol * xN;
      const bool termination1 = std::fabs(actualReduction) <= pars.epsmch && predictedReduction <= pars.epsmch && 0.5 * ratio <= 1.0;
      const bool termination2 = delta <= pars.epsmch * xN;
      stoppingRule = convergence1 || convergence2 || termination1 || termination2;
      delta2 = delta;
      if (stoppingRule || inner >= pars.maxInnerIter || ratio >= 0.0001) break;                        // :218
      xin = x; dg = xDiag; dl = delta; par0 = parNew; fN0 = fNorm;                                     // :221
    }
    if (stoppingRule || outer >= pars.maxIter || gNorm <= pars.gTol || gNorm <= pars.epsmch) return x;  // :145
    x0 = x; diag0 = diag; xNorm0 = fNorm; delta0 = delta2;                                              // :148
  }
}

}  // namespace LevenbergMarquardt
}  // namespace scalaopt
