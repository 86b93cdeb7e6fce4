// gpu.hpp — the GPU-backed data set and objective functions: what a user of the reference switches to.
//
//   GpuDataSet      extends DataSet[DataPoint]   (DataSet.scala:27-145)  observations resident in HBM
//   GpuMSEFunction  extends MSEFunction          (function/MSEFunction.scala:12-76; precedent for a custom
//                   MSEFunction with analytic derivatives: std-apps/.../nnet/FFNeuralNetworkTrainer.scala:36-144)
//
// Every evaluation is one call into libscalaopt_cuda.so (include/scalaopt_cuda.h); nothing is computed on
// the host except O(n^3) algebra on the n x n results.  There is no CPU fallback: closure-taking DataSet methods
// stream a bounded number of rows back or throw UnsupportedOperationException.
#pragma once
#include <cstring>
#include <list>

#include "../../include/scalaopt_cuda.h"
#include "scalaopt_host.hpp"

namespace scalaopt {

struct GpuError : std::runtime_error { int code; GpuError(int c, const std::string& m) : std::runtime_error(m), code(c) {} };

class GpuDataSet : public DataSet<DataPoint> {
 public:
  static constexpr int64_t kCollectLimit = 1 << 20;      // rows a closure-taking method may stream back
  so_ctx* ctx; so_dataset* ds; int f; bool owns;
  GpuDataSet(so_ctx* c, so_dataset* d, int f_, bool owns_ = false) : ctx(c), ds(d), f(f_), owns(owns_) {}
  ~GpuDataSet() override { if (owns && ds) so_dataset_free(ds); }
  void check(int rc) const { if (rc != SO_OK) throw GpuError(rc, so_last_error(ctx)); }

  // ---- converters: SeqDataSet / any DataSet[DataPoint] -> GpuDataSet, raw fp64 files -> GpuDataSet ----
  static std::shared_ptr<GpuDataSet> from(so_ctx* ctx, const DataSet<DataPoint>& data) {
    const std::vector<DataPoint> rows = data.collect();
    if (rows.empty()) throw IllegalArgumentException("GpuDataSet.from: empty data set");
    const int f = (int)rows.front().x.size();
    so_dataset* nd = nullptr;
    if (int rc = so_dataset_create(ctx, (int64_t)rows.size(), f, &nd)) throw GpuError(rc, so_last_error(ctx));
    auto out = std::make_shared<GpuDataSet>(ctx, nd, f, true);
    std::vector<double> X(rows.size() * (size_t)f), y(rows.size());
    for (size_t i = 0; i < rows.size(); ++i) { std::copy(rows[i].x.begin(), rows[i].x.end(), X.begin() + i * f); y[i] = rows[i].y; }
    out->check(so_dataset_append(nd, X.data(), y.data(), (int64_t)rows.size()));
    return out;
  }
  static std::shared_ptr<GpuDataSet> fromRawFiles(so_ctx* ctx, const std::string& path_X, const std::string& path_y, int64_t rows, int f) {
    so_dataset* nd = nullptr;
    if (int rc = so_dataset_create(ctx, rows, f, &nd)) throw GpuError(rc, so_last_error(ctx));
    auto out = std::make_shared<GpuDataSet>(ctx, nd, f, true);
    out->check(so_dataset_load_raw(nd, path_X.c_str(), path_y.c_str(), rows, 0));
    return out;
  }

  int64_t size() const override { int64_t m = 0; check(so_dataset_size(ds, &m)); return m; }                 // DataSet.scala:126
  DataPoint head() const override { DataPoint p; p.x.resize(f); check(so_dataset_head(ds, p.x.data(), &p.y)); return p; }   // :82
  std::vector<DataPoint> collect() const override {                                                           // :51, bounded
    const int64_t m = size();
    if (m > kCollectLimit) throw UnsupportedOperationException("GpuDataSet.collect: data set too large to materialise on the host");
    std::vector<double> X((size_t)m * f), y((size_t)m);
    check(so_dataset_read(ds, 0, m, X.data(), y.data()));
    std::vector<DataPoint> out((size_t)m);
    for (int64_t i = 0; i < m; ++i) { out[i].x.assign(X.begin() + i * f, X.begin() + (i + 1) * f); out[i].y = y[i]; }
    return out;
  }
  // closures cannot run on the device: bounded host streaming or an exception (no silent CPU evaluation path)
  void aggregate_raw(const std::function<void(const DataPoint&)>& seqop) const override { for (const DataPoint& p : collect()) seqop(p); }
  void foreach (const std::function<void(const DataPoint&)>& fn) const override { for (const DataPoint& p : collect()) fn(p); }
  std::shared_ptr<DataSet<DataPoint>> filter(const std::function<bool(const DataPoint&)>& p) const override {
    return SeqDataSet<DataPoint>(collect()).filter(p);
  }
  std::shared_ptr<DataSet<DataPoint>> map(const std::function<DataPoint(const DataPoint&)>& fn) const override {
    return SeqDataSet<DataPoint>(collect()).map(fn);
  }
  std::shared_ptr<DataSet<DataPoint>> concat(const DataSet<DataPoint>& that) const override {                // ++ :36
    const auto rows = that.collect();
    const int64_t m0 = size(), m1 = (int64_t)rows.size();
    so_dataset* nd = nullptr;
    check(so_dataset_create(ctx, m0 + m1, f, &nd));
    auto out = std::make_shared<GpuDataSet>(ctx, nd, f, true);
    const int64_t CH = 1 << 20;
    std::vector<double> X, y;
    for (int64_t b = 0; b < m0; b += CH) {
      const int64_t e = std::min(m0, b + CH);
      X.resize((size_t)(e - b) * f); y.resize((size_t)(e - b));
      check(so_dataset_read(ds, b, e - b, X.data(), y.data()));
      check(so_dataset_append(nd, X.data(), y.data(), e - b));
    }
    X.resize((size_t)m1 * f); y.resize((size_t)m1);
    for (int64_t i = 0; i < m1; ++i) { std::copy(rows[i].x.begin(), rows[i].x.end(), X.begin() + i * f); y[i] = rows[i].y; }
    if (m1) check(so_dataset_append(nd, X.data(), y.data(), m1));
    return out;
  }
};

// Counters of what the objective asked the device to do (the reference has no metrics).
struct GpuCounters {
  int64_t loss_calls = 0, gradient_calls = 0, dirder_calls = 0, hessvec_calls = 0, jacobian_calls = 0;   // requests from the optimizer
  int64_t k1_passes = 0, k2_passes = 0, k3_passes = 0, k4_passes = 0;                                     // passes over the data
  int64_t cache_hits = 0, line_hits = 0;
  int64_t spec_passes = 0, spec_hits = 0;      // speculative K2 passes at LM trial points / Jacobians served from them
};

class GpuMSEFunction : public MSEFunction {
 public:
  std::shared_ptr<GpuDataSet> data; int family; int n;
  int line_batch = 3;               // step sizes scored per K3 pass once a line search needs a second trial point
  bool use_tsqr = false;            // jacobianAndResidualsMatrix from so_eval_qr (Householder on the rows) instead of K2 + Cholesky
  // Levenberg-Marquardt asks for the loss at a trial point and, when it accepts the step (the usual case), for the
  // Jacobian system at that same point next (LevenbergMarquardt.scala:124,164).  With `speculate` the trial-point loss
  // is taken from a K2 pass (loss = r^T r / 2m) whose J^T J, J^T r are kept: an accepted step costs one pass over the
  // data instead of two (K1 loss 2.6 ms + K2 3.9 ms -> K2 3.9 ms at 1e9 rows); a rejected one costs K2 instead of K1.
  // Only the FIRST trial point after a Jacobian is speculated on (later ones follow a rejection and are mostly
  // rejected again), and only where a K2 pass costs about what a K1 pass costs (n <= 16: the Gram is HBM-bound).
  bool speculate = true;
  static constexpr int kSpeculateMaxN = 16;
  // FFNeuralNetworkTrainer's weight decay (std-apps/.../nnet/FFNeuralNetworkTrainer.scala:59-69,100-108,147-153): the loss fold
  // starts from |w|^2 / 2 * decay, the gradient fold from w * decay (both divided by data.size), and the Jacobian/residual
  // matrix gains n rows sqrt(decay) e_i | 0.  Data-independent, O(n): applied here, around the device sums.
  double decay = 0.0;
  mutable GpuCounters counters;
  GpuMSEFunction(std::shared_ptr<GpuDataSet> d, int family_, int n_) : data(std::move(d)), family(family_), n(n_) {}

  // ---- ObjectiveFunction.apply: loss at x (MSEFunction.scala:33-36 = sum r^2/2 / m; logistic: mean cross-entropy) ----
  double apply(const Vec& x) const override {
    counters.loss_calls++;
    if (const Entry* e = find(x)) { counters.cache_hits++; return e->loss; }
    if (const LineEntry* le = find_on_ray(x)) { counters.line_hits++; return le->phi; }
    if (on_ray_beta(x)) return score_ray(x).phi;
    if (loss_only_mode && speculate && use_tsqr && loss_scale() > 0.0 && trials_since_jacobian == 0) {
      // same idea on the TSQR route: r^T r is the squared norm of R's last column
      ++trials_since_jacobian;
      spec.R.assign((size_t)(n + 1) * (n + 1), 0.0);
      check(so_eval_qr(data->ds, family, x.data(), n, spec.R.data()));
      spec.x = x; spec.valid = true;
      counters.k2_passes++; counters.spec_passes++;
      double rtr = 0.0;
      for (int i = 0; i <= n; ++i) rtr += spec.R[(size_t)i * (n + 1) + n] * spec.R[(size_t)i * (n + 1) + n];
      const double loss = rtr * loss_scale() / (double)data->size() + decay_loss(x);
      remember(x, loss, nullptr);
      return loss;
    }
    if (loss_only_mode && speculate && !use_tsqr && loss_scale() > 0.0 && n <= kSpeculateMaxN && trials_since_jacobian++ == 0) {
      spec.JtJ.assign((size_t)n * n, 0.0); spec.Jtr.assign(n, 0.0);
      check(so_eval_normal_eq(data->ds, family, x.data(), n, spec.JtJ.data(), spec.Jtr.data(), &spec.rtr));
      spec.x = x; spec.valid = true;
      counters.k2_passes++; counters.spec_passes++;
      const double loss = spec.rtr * loss_scale() / (double)data->size() + decay_loss(x);
      remember(x, loss, nullptr);
      return loss;
    }
    if (loss_only_mode) {
      double loss = 0.0;
      check(so_eval_loss_grad(data->ds, family, x.data(), n, &loss, nullptr));
      counters.k1_passes++;
      loss += decay_loss(x);
      remember(x, loss, nullptr);
      return loss;
    }
    return fused(x).loss;
  }
  Vec gradient(const Vec& x) const override {
    counters.gradient_calls++;
    if (const Entry* e = find(x)) if (e->has_grad) { counters.cache_hits++; return e->grad; }
    return fused(x).grad;
  }
  // dirder = gradient(x) dot d (ObjectiveFunctionWithGradient.scala:37-39); also tells the objective which ray
  // the line search is walking, so that further trial points on it can be scored in batches (K3)
  double dirder(const Vec& x, const Vec& d) const override {
    counters.dirder_calls++;
    if (const Entry* e = find(x)) if (e->has_grad) { counters.cache_hits++; note_ray(x, d); return dot(e->grad, d); }
    if (same_dir(d)) if (const LineEntry* le = find_on_ray(x)) { counters.line_hits++; return le->dphi; }
    const Vec g = fused(x).grad;
    note_ray(x, d);
    return dot(g, d);
  }
  Vec dirHessian(const Vec& x, const Vec& d) const override {   // exact product (K4), not the 2-gradient finite difference
    counters.hessvec_calls++;
    Vec out(n);
    check(so_eval_hess_vec(data->ds, family, x.data(), d.data(), n, out.data()));
    counters.k4_passes++;
    // the trainer's dirHessian differences gradients that contain w decay / m (Trainer.scala:64-69,76-80): + decay d / m
    if (decay > 0.0) { const double m = (double)data->size(); for (int i = 0; i < n; ++i) out[i] += decay * d[i] / m; }
    return out;
  }
  // The (n+1)-row system with the same R, Q^T b, column norms and ||b|| (up to row signs) as the m-row
  // Jacobian/residual matrix of MSEFunction.scala:132-136: rows (R_chol[i,:] | q_i), then (0 | rho).
  std::shared_ptr<DataSet<AugmentedRow>> jacobianAndResidualsMatrix(const Vec& p) const override {
    counters.jacobian_calls++;
    loss_only_mode = true;          // Levenberg-Marquardt only ever asks for the loss between Jacobians
    trials_since_jacobian = 0;
    if (use_tsqr) {
      // the rows of the R factor of [J | r] ARE the (n+1)-row system; no Gram matrix, no Cholesky
      std::vector<double> Rf((size_t)(n + 1) * (n + 1));
      if (spec.valid && same_bits(spec.x, p) && spec.R.size() == Rf.size()) {
        Rf = spec.R;
        counters.spec_hits++;
      } else {
        check(so_eval_qr(data->ds, family, p.data(), n, Rf.data()));
        counters.k2_passes++;
      }
      double rtr = 0.0;
      for (int i = 0; i <= n; ++i) rtr += Rf[(size_t)i * (n + 1) + n] * Rf[(size_t)i * (n + 1) + n];
      // the loss at p is a function of r^T r only for the MSE families (cross-entropy: K1 computes it when LM asks)
      if (loss_scale() > 0.0) { int64_t m = data->size(); remember(p, rtr * loss_scale() / (double)m + decay_loss(p), nullptr); }
      std::vector<AugmentedRow> rows;
      for (int i = 0; i <= n; ++i)
        rows.emplace_back(Vec(Rf.begin() + (size_t)i * (n + 1), Rf.begin() + (size_t)i * (n + 1) + n), Rf[(size_t)i * (n + 1) + n], (int64_t)i);
      append_penalty_rows(rows);
      return std::make_shared<SeqDataSet<AugmentedRow>>(std::move(rows));
    }
    std::vector<double> JtJ((size_t)n * n), Jtr(n);
    double rtr = 0.0;
    if (spec.valid && same_bits(spec.x, p) && spec.JtJ.size() == JtJ.size()) {   // the accepted trial point: its sums are already here
      JtJ = spec.JtJ; Jtr = spec.Jtr; rtr = spec.rtr;
      counters.spec_hits++;
    } else {
      check(so_eval_normal_eq(data->ds, family, p.data(), n, JtJ.data(), Jtr.data(), &rtr));
      counters.k2_passes++;
    }
    // the loss at p comes for free: sum r^2 / (2m) for the MSE families; FFNeuralNetwork's MSE loss is not halved
    // (FFNeuralNetwork.scala:70-72); the cross-entropy families' loss is not a function of r^T r
    if (loss_scale() > 0.0) { int64_t m = data->size(); remember(p, rtr * loss_scale() / (double)m + decay_loss(p), nullptr); }
    // weight decay: the penalty rows sqrt(decay) e_i | 0 of the trainer's matrix (Trainer.scala:147-153) add decay to the
    // diagonal of the Gram; folding them in BEFORE the factorisation matters — J alone is often rank-deficient for a
    // network (soft-max redundancy: cond(J) ~ 1e35 on the Iris-shaped spec case) while [J; sqrt(decay) I] never is
    if (decay > 0.0) for (int i = 0; i < n; ++i) JtJ[(size_t)i * n + i] += decay;
    // Cholesky JtJ = R^T R (upper R), q = R^-T Jtr, rho = sqrt(max(0, rtr - q.q)); O(n^3) on the host
    std::vector<Vec> R(n, Vec(n, 0.0));
    Vec q(n, 0.0);
    for (int i = 0; i < n; ++i) {
      double dsum = JtJ[(size_t)i * n + i];
      for (int k = 0; k < i; ++k) dsum -= R[k][i] * R[k][i];
      if (!(dsum > 0.0)) continue;                               // rank-deficient column: zero row -> rDiag = 0 downstream
      const double rii = std::sqrt(dsum);
      R[i][i] = rii;
      for (int j = i + 1; j < n; ++j) {
        double s = JtJ[(size_t)i * n + j];
        for (int k = 0; k < i; ++k) s -= R[k][i] * R[k][j];
        R[i][j] = s / rii;
      }
      double s = Jtr[i];
      for (int k = 0; k < i; ++k) s -= R[k][i] * q[k];
      q[i] = s / rii;
    }
    const double rho = std::sqrt(std::max(0.0, rtr - inner(q, q)));
    std::vector<AugmentedRow> rows;
    for (int i = 0; i < n; ++i) rows.emplace_back(R[i], q[i], (int64_t)i);
    rows.emplace_back(Vec(n, 0.0), rho, (int64_t)n);
    return std::make_shared<SeqDataSet<AugmentedRow>>(std::move(rows));
  }

 private:
  struct Entry { Vec x; double loss; bool has_grad; Vec grad; };
  struct Spec { bool valid = false; Vec x; std::vector<double> JtJ, Jtr, R; double rtr = 0.0; };
  mutable Spec spec;
  mutable int trials_since_jacobian = 0;
  // loss = loss_scale * r^T r / m where the loss is a function of r^T r (0: it is not)
  double loss_scale() const {
    if (family == SO_FAMILY_MLP_MSE) return 1.0;                  // FFNeuralNetwork.scala:70-72, not halved
    if (family == SO_FAMILY_LOGISTIC || family == SO_FAMILY_MLP_CE || family == SO_FAMILY_SOFTMAX) return 0.0;
    return 0.5;                                                   // MSEFunction.scala:34
  }
  double decay_loss(const Vec& x) const { return decay > 0.0 ? 0.5 * decay * dot(x, x) / (double)data->size() : 0.0; }
  void append_penalty_rows(std::vector<AugmentedRow>& rows) const {      // penaltyMatrix, Trainer.scala:147-153
    if (!(decay > 0.0)) return;
    const double pen = std::sqrt(decay);
    const int64_t base = (int64_t)rows.size();
    for (int i = 0; i < n; ++i) { Vec e(n, 0.0); e[i] = pen; rows.emplace_back(std::move(e), 0.0, base + i); }
  }
  struct LineEntry { double beta, phi, dphi; };
  mutable std::list<Entry> cache;                 // most recent first; keyed on the bit pattern of x
  mutable bool loss_only_mode = false;
  mutable bool have_ray = false;
  mutable Vec ray_origin, ray_d;
  mutable int ray_trials = 0;                      // distinct points evaluated on the current ray
  mutable std::vector<LineEntry> line;

  void check(int rc) const { if (rc != SO_OK) throw GpuError(rc, so_last_error(data->ctx)); }
  static bool same_bits(const Vec& a, const Vec& b) { return a.size() == b.size() && std::memcmp(a.data(), b.data(), a.size() * sizeof(double)) == 0; }
  const Entry* find(const Vec& x) const { for (const Entry& e : cache) if (same_bits(e.x, x)) return &e; return nullptr; }
  void remember(const Vec& x, double loss, const Vec* g) const {
    for (Entry& e : cache) if (same_bits(e.x, x)) { e.loss = loss; if (g) { e.has_grad = true; e.grad = *g; } return; }
    cache.push_front(Entry{x, loss, g != nullptr, g ? *g : Vec()});
    if (cache.size() > 8) cache.pop_back();
  }
  const Entry& fused(const Vec& x) const {          // K1: loss and gradient in one pass
    double loss = 0.0; Vec g(n);
    check(so_eval_loss_grad(data->ds, family, x.data(), n, &loss, g.data()));
    counters.k1_passes++;
    if (decay > 0.0) {
      const double m = (double)data->size();
      loss += decay_loss(x);
      for (int i = 0; i < n; ++i) g[i] += x[i] * decay / m;
    }
    remember(x, loss, &g);
    if (on_ray_beta(x)) ray_trials++;
    return *find(x);
  }
  bool same_dir(const Vec& d) const { return have_ray && same_bits(d, ray_d); }
  void note_ray(const Vec& x, const Vec& d) const {
    if (same_dir(d)) return;
    have_ray = true; ray_origin = x; ray_d = d; ray_trials = 1; line.clear();
  }
  // x = origin + beta d ?  (tolerance: a few ulp of the point's magnitude)
  std::optional<double> on_ray_beta(const Vec& x) const {
    if (!have_ray) return std::nullopt;
    const double dd = dot(ray_d, ray_d);
    if (!(dd > 0.0)) return std::nullopt;
    Vec diff = sub(x, ray_origin);
    const double beta = dot(diff, ray_d) / dd;
    double err = 0.0, scale = 0.0;
    for (int i = 0; i < n; ++i) { err = std::max(err, std::fabs(diff[i] - beta * ray_d[i])); scale = std::max(scale, std::fabs(x[i]) + std::fabs(beta * ray_d[i])); }
    if (err > 64.0 * std::numeric_limits<double>::epsilon() * std::max(scale, 1e-300)) return std::nullopt;
    return beta;
  }
  const LineEntry* find_on_ray(const Vec& x) const {
    if (line.empty()) return nullptr;
    const auto beta = on_ray_beta(x);
    if (!beta) return nullptr;
    for (const LineEntry& le : line) if (std::fabs(le.beta - *beta) <= 1e-12 * std::max(1.0, std::fabs(*beta))) return &le;
    return nullptr;
  }
  // second and later trial points of a line search: one K3 pass scores this step size together with the
  // continuation of the bracket schedule alpha <- c3 alpha (StrongWolfe.scala:114), i.e. beta <- 2 beta + 1
  const LineEntry& score_ray(const Vec& x) const {
    const double beta = *on_ray_beta(x);
    std::vector<double> al{beta};
    double b = beta;
    for (int i = 1; i < line_batch && ray_trials >= 1; ++i) { b = 2.0 * b + 1.0; al.push_back(b); }
    std::vector<double> phi(al.size()), dphi(al.size());
    check(so_eval_line(data->ds, family, ray_origin.data(), ray_d.data(), n, al.data(), (int)al.size(), phi.data(), dphi.data()));
    counters.k3_passes++;
    if (decay > 0.0) {
      // the device scores the data term only; the trainer's weight decay (loss |w|^2/2 decay / m, gradient w decay / m,
      // see fused) belongs to every point of the ray as well: phi += decay |x_a|^2 / (2m), phi' += decay (x_a . d) / m
      const double m = (double)data->size();
      for (size_t i = 0; i < al.size(); ++i) {
        const Vec xa = add(ray_origin, mul(ray_d, al[i]));
        phi[i] += decay_loss(xa);
        dphi[i] += decay * dot(xa, ray_d) / m;
      }
    }
    for (size_t i = 0; i < al.size(); ++i) line.push_back(LineEntry{al[i], phi[i], dphi[i]});
    ray_trials++;
    return line[line.size() - al.size()];
  }
};

// linear.lm(data, addOrigin) (linear/package.scala:54-63): the least-squares solution of a beta = y through
// QR(rows (a | y), n).solution, a = (1, x) with the origin.  On a GpuDataSet the m rows never reach the host: at p = 0 the
// linear family's Jacobian/residual rows are s_i (-a_i | y_i), s_i = sign(y_i), so the (n+1)-row system of
// GpuMSEFunction carries the Gram of (-a | y) and the reference's own QR(...).solution on it is -beta.  One pass over the
// data (K2, or the on-device Householder TSQR with use_tsqr) against the 2n+1 passes of QR.apply.
inline Vec gpu_lm(const std::shared_ptr<GpuDataSet>& data, bool add_origin = true, bool use_tsqr = false) {
  const int n = data->f + (add_origin ? 1 : 0);
  GpuMSEFunction fn(data, add_origin ? SO_FAMILY_LINEAR : SO_FAMILY_LINEAR_NOINT, n);
  fn.use_tsqr = use_tsqr;
  Vec beta = QR::solve(*fn.jacobianAndResidualsMatrix(Vec(n, 0.0)), n);
  for (double& b : beta) b = -b;
  return beta;
}

// negLogLikelihood(pdf, data) _ as an objective (std-apps/.../stats/package.scala:74-103).  The reference hands this plain
// closure to an optimizer through `toFunctionWoGradient` (core/.../package.scala:74-77), i.e. as an
// ObjectiveFunctionFiniteDiffDerivatives: derivatives are forward differences of the (GPU-evaluated) objective.
class GpuNllFunction : public MSEFunction {
 public:
  std::shared_ptr<GpuDataSet> data; int pdf; int n; Vec consts; ConfigPars config;
  mutable GpuCounters counters;
  GpuNllFunction(std::shared_ptr<GpuDataSet> d, int pdf_, int n_, Vec consts_) : data(std::move(d)), pdf(pdf_), n(n_), consts(std::move(consts_)) {}
  double apply(const Vec& x) const override {
    counters.loss_calls++; counters.k1_passes++;
    double out = 0.0;
    const int rc = so_eval_nll(data->ds, pdf, x.data(), n, consts.data(), (int)consts.size(), &out);
    if (rc != SO_OK) throw GpuError(rc, so_last_error(data->ctx));
    return out;
  }
  Vec gradient(const Vec& x) const override { counters.gradient_calls++; return fd_gradient([this](const Vec& v) { return apply(v); }, x, config.eps); }
  double dirder(const Vec& x, const Vec& d) const override { counters.dirder_calls++; return (apply(add(x, mul(d, config.eps))) - apply(x)) / config.eps; }
  Vec dirHessian(const Vec& x, const Vec& d) const override {
    counters.hessvec_calls++;
    const Vec gradx = gradient(x);
    const Vec gradxd = gradient(add(x, mul(d, config.eps)));
    return div(sub(gradxd, gradx), config.eps);
  }
  std::shared_ptr<DataSet<AugmentedRow>> jacobianAndResidualsMatrix(const Vec&) const override {
    throw IllegalArgumentException("a negative log-likelihood is not an MSEFunction");
  }
};

}  // namespace scalaopt
