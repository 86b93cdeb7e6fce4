// host_abi.cpp — C ABI (include/scalaopt_host.h) over the C++ mirror of the reference interface.
#include "../../include/scalaopt_host.h"

#include <cstring>

#include "gpu.hpp"
#include "scalaopt_host.hpp"

using namespace scalaopt;

namespace {
thread_local std::string g_err;

// objective built from C callbacks: ObjectiveFunctionWithGradient / ...FiniteDiffDerivatives / an MSEFunction
class CallbackObjective : public MSEFunction {
 public:
  int n; soh_apply_fn fa; soh_gradient_fn fg; soh_hessvec_fn fh; soh_jacobian_fn fj; int64_t m; void* user; double eps;
  mutable GpuCounters counters;
  CallbackObjective(int n_, soh_apply_fn a, soh_gradient_fn g, soh_hessvec_fn h, soh_jacobian_fn j, int64_t m_, void* u, double e)
      : n(n_), fa(a), fg(g), fh(h), fj(j), m(m_), user(u), eps(e) {}
  double apply(const Vec& x) const override { counters.loss_calls++; return fa(user, x.data(), n); }
  Vec gradient(const Vec& x) const override {
    counters.gradient_calls++;
    if (fg) { Vec g(n); fg(user, x.data(), n, g.data()); return g; }
    return fd_gradient([this](const Vec& v) { return fa(user, v.data(), n); }, x, eps);    // DenseVector.gradient
  }
  double dirder(const Vec& x, const Vec& d) const override {
    counters.dirder_calls++;
    if (fg) return dot(gradient(x), d);                                                    // ObjectiveFunctionWithGradient.scala:37-39
    return (fa(user, add(x, mul(d, eps)).data(), n) - fa(user, x.data(), n)) / eps;        // ...FiniteDiffDerivatives.scala:40-42
  }
  Vec dirHessian(const Vec& x, const Vec& d) const override {
    counters.hessvec_calls++;
    if (fh) { Vec out(n); fh(user, x.data(), d.data(), n, out.data()); return out; }
    const Vec gradx = gradient(x);                                                         // :41-47 / :44-50
    const Vec gradxd = gradient(add(x, mul(d, eps)));
    return div(sub(gradxd, gradx), eps);
  }
  std::shared_ptr<DataSet<AugmentedRow>> jacobianAndResidualsMatrix(const Vec& p) const override {
    counters.jacobian_calls++;
    if (!fj) throw IllegalArgumentException("objective has no jacobianAndResidualsMatrix (not an MSEFunction)");
    std::vector<double> rows((size_t)m * (n + 1));
    fj(user, p.data(), n, rows.data(), m);
    std::vector<AugmentedRow> v;
    v.reserve((size_t)m);
    for (int64_t i = 0; i < m; ++i) {
      const double* r = rows.data() + (size_t)i * (n + 1);
      v.emplace_back(Vec(r, r + n), r[n], i);                    // data.zipWithIndex map AugmentedRow(..., i), MSEFunction.scala:132-136
    }
    return std::make_shared<SeqDataSet<AugmentedRow>>(std::move(v));
  }
};
}  // namespace

struct soh_objective {
  std::unique_ptr<MSEFunction> fn;
  GpuCounters* counters = nullptr;
  int n = 0;
  bool has_jacobian = false;
};

namespace {
template <class F>
int guarded(F&& body) {
  try { body(); return SOH_OK; }
  catch (const MaxIterException& e) { g_err = std::string("MaxIterException: ") + e.what(); return SOH_MAX_ITER; }
  catch (const IllegalArgumentException& e) { g_err = std::string("IllegalArgumentException: ") + e.what(); return SOH_ILLEGAL_ARG; }
  catch (const GpuError& e) { g_err = std::string("GpuError ") + std::to_string(e.code) + ": " + e.what(); return SOH_GPU_ERROR; }
  catch (const std::exception& e) { g_err = e.what(); return SOH_ERROR; }
}
StrongWolfeConfig sw_of(const soh_config& c) { return StrongWolfeConfig{c.max_iter_line, c.max_iter_zoom, c.c1, c.c2, c.c3}; }
}  // namespace

extern "C" const char* soh_last_error(void) { return g_err.c_str(); }

extern "C" int soh_default_config(int method, soh_config* o) {
  if (!o) return SOH_ILLEGAL_ARG;
  std::memset(o, 0, sizeof *o);
  o->tol = 1.0e-5; o->max_iter = 200; o->eps = 1.0e-8;
  o->max_iter_line = 10; o->max_iter_zoom = 10; o->max_iter_first_time = -1;
  o->c1 = 1.0e-4; o->c2 = (method == SOH_BFGS) ? 0.9 : 0.4; o->c3 = 2.0; o->cg_method = 2;
  o->delta0 = 1.0; o->delta_max = 1.0e5; o->eta = 0.2;
  o->x_tol = 1.49012e-8; o->g_tol = 0.0; o->ratio_tol = 0.0001; o->step_bound = 100.0; o->epsmch = 2.22044604926e-16;
  o->max_inner_iter = 10; o->max_step_length_iter = 10; o->use_pivoting = 0;
  o->c_reflection = 2.0; o->c_expansion = 1.0; o->c_contraction = 0.5; o->c_shrink = 0.5; o->rel_delta = 0.05; o->abs_delta = 0.00025;
  if (method == SOH_LEVENBERG_MARQUARDT) { o->tol = 1.49012e-8; o->max_iter = 10; }
  return SOH_OK;
}

extern "C" soh_objective* soh_objective_from_callbacks(int n, soh_apply_fn apply, soh_gradient_fn gradient, soh_hessvec_fn hessvec,
                                                       soh_jacobian_fn jacobian, int64_t m, void* user, double eps) {
  if (n < 1 || !apply) { g_err = "soh_objective_from_callbacks: n >= 1 and apply required"; return nullptr; }
  auto* cb = new CallbackObjective(n, apply, gradient, hessvec, jacobian, m, user, eps > 0.0 ? eps : 1.0e-8);
  auto* o = new soh_objective();
  o->counters = &cb->counters; o->n = n; o->has_jacobian = jacobian != nullptr;
  o->fn.reset(cb);
  return o;
}

extern "C" soh_objective* soh_objective_gpu(so_ctx* ctx, so_dataset* ds, int f, int family, int n) {
  if (!ctx || !ds || n < 1) { g_err = "soh_objective_gpu: bad arguments"; return nullptr; }
  auto data = std::make_shared<GpuDataSet>(ctx, ds, f, false);
  auto* g = new GpuMSEFunction(data, family, n);
  auto* o = new soh_objective();
  o->counters = &g->counters; o->n = n; o->has_jacobian = true;
  o->fn.reset(g);
  return o;
}

extern "C" int soh_objective_use_tsqr(soh_objective* o, int on) {
  auto* g = o ? dynamic_cast<GpuMSEFunction*>(o->fn.get()) : nullptr;
  if (!g) { g_err = "soh_objective_use_tsqr: not a GPU least-squares objective"; return SOH_ILLEGAL_ARG; }
  g->use_tsqr = on != 0;
  return SOH_OK;
}

extern "C" int soh_objective_set_decay(soh_objective* o, double decay) {
  auto* g = o ? dynamic_cast<GpuMSEFunction*>(o->fn.get()) : nullptr;
  if (!g || !(decay >= 0.0)) { g_err = "soh_objective_set_decay: not a GPU least-squares objective, or decay < 0"; return SOH_ILLEGAL_ARG; }
  g->decay = decay;
  return SOH_OK;
}

extern "C" int soh_objective_speculate(soh_objective* o, int on) {
  auto* g = o ? dynamic_cast<GpuMSEFunction*>(o->fn.get()) : nullptr;
  if (!g) { g_err = "soh_objective_speculate: not a GPU least-squares objective"; return SOH_ILLEGAL_ARG; }
  g->speculate = on != 0;
  return SOH_OK;
}

extern "C" soh_objective* soh_objective_gpu_nll(so_ctx* ctx, so_dataset* ds, int pdf, int n, const double* consts, int nconsts) {
  if (!ctx || !ds || n < 1 || (nconsts > 0 && !consts)) { g_err = "soh_objective_gpu_nll: bad arguments"; return nullptr; }
  auto data = std::make_shared<GpuDataSet>(ctx, ds, 1, false);
  auto* g = new GpuNllFunction(data, pdf, n, Vec(consts, consts + nconsts));
  auto* o = new soh_objective();
  o->counters = &g->counters; o->n = n; o->has_jacobian = false;
  o->fn.reset(g);
  return o;
}

extern "C" void soh_objective_free(soh_objective* f) { delete f; }

extern "C" int soh_objective_counters(soh_objective* f, soh_counters* out) {
  if (!f || !out) return SOH_ILLEGAL_ARG;
  const GpuCounters& c = *f->counters;
  *out = soh_counters{c.loss_calls, c.gradient_calls, c.dirder_calls, c.hessvec_calls, c.jacobian_calls,
                      c.k1_passes, c.k2_passes, c.k3_passes, c.k4_passes, c.cache_hits, c.line_hits,
                      c.spec_passes, c.spec_hits};
  return SOH_OK;
}

extern "C" int soh_objective_apply(soh_objective* f, const double* x, double* out) {
  if (!f || !x || !out) return SOH_ILLEGAL_ARG;
  return guarded([&] { *out = f->fn->apply(Vec(x, x + f->n)); });
}
extern "C" int soh_objective_gradient(soh_objective* f, const double* x, double* g) {
  if (!f || !x || !g) return SOH_ILLEGAL_ARG;
  return guarded([&] { const Vec r = f->fn->gradient(Vec(x, x + f->n)); std::copy(r.begin(), r.end(), g); });
}
extern "C" int soh_objective_dirder(soh_objective* f, const double* x, const double* d, double* out) {
  if (!f || !x || !d || !out) return SOH_ILLEGAL_ARG;
  return guarded([&] { *out = f->fn->dirder(Vec(x, x + f->n), Vec(d, d + f->n)); });
}
extern "C" int soh_objective_dir_hessian(soh_objective* f, const double* x, const double* d, double* out) {
  if (!f || !x || !d || !out) return SOH_ILLEGAL_ARG;
  return guarded([&] { const Vec r = f->fn->dirHessian(Vec(x, x + f->n), Vec(d, d + f->n)); std::copy(r.begin(), r.end(), out); });
}
extern "C" int soh_objective_jacobian_rows(soh_objective* f, const double* p, double* rows, int64_t cap, int64_t* m_out) {
  if (!f || !p || !rows || !m_out) return SOH_ILLEGAL_ARG;
  return guarded([&] {
    const auto ds = f->fn->jacobianAndResidualsMatrix(Vec(p, p + f->n));
    const auto v = ds->collect();
    *m_out = (int64_t)v.size();
    if ((int64_t)v.size() > cap) throw IllegalArgumentException("rows buffer too small");
    for (size_t i = 0; i < v.size(); ++i) {
      std::copy(v[i].a.begin(), v[i].a.end(), rows + i * (f->n + 1));
      rows[i * (f->n + 1) + f->n] = v[i].b;
    }
  });
}

extern "C" int soh_minimize(int method, soh_objective* f, const double* x0, int n, const soh_config* cfg, double* x_out,
                            soh_result* res) {
  if (!f || !x0 || !x_out || n != f->n) { g_err = "soh_minimize: bad arguments"; return SOH_ILLEGAL_ARG; }
  soh_config c;
  if (cfg) c = *cfg; else soh_default_config(method, &c);
  soh_result r{};
  const Vec start(x0, x0 + n);
  Vec sol;
  IterationTrace tr;
  const int rc = guarded([&] {
    switch (method) {
      case SOH_LEVENBERG_MARQUARDT: {
        if (!f->has_jacobian) throw IllegalArgumentException("Levenberg-Marquardt needs an MSEFunction");
        LevenbergMarquardtConfig p;
        p.tol = c.tol; p.maxIter = c.max_iter; p.eps = c.eps; p.xTol = c.x_tol; p.gTol = c.g_tol; p.ratioTol = c.ratio_tol;
        p.stepBound = c.step_bound; p.epsmch = c.epsmch; p.maxInnerIter = c.max_inner_iter;
        p.maxStepLengthIter = c.max_step_length_iter; p.usePivoting = c.use_pivoting != 0;
        LevenbergMarquardt::Trace t;
        sol = LevenbergMarquardt::minimize(*f->fn, start, p, &t);
        r.iterations = t.outerIterations; r.inner_iterations = t.innerIterations;
        r.gauss_newton_steps = t.gaussNewtonSteps; r.damped_steps = t.dampedSteps;
        break;
      }
      case SOH_BFGS: {
        BFGSConfig p; p.tol = c.tol; p.maxIter = c.max_iter; p.eps = c.eps; p.maxIterLine = c.max_iter_line; p.maxIterZoom = c.max_iter_zoom;
        if (c.max_iter_first_time >= 0) p.maxIterFirstTime = c.max_iter_first_time;
        p.c1 = c.c1; p.c2 = c.c2; p.c3 = c.c3;
        sol = BFGS::minimize(*f->fn, start, p, &tr);
        break;
      }
      case SOH_CONJUGATE_GRADIENT: case SOH_NEWTON_CG: {
        CGConfig p; p.tol = c.tol; p.maxIter = c.max_iter; p.eps = c.eps; p.maxIterLine = c.max_iter_line; p.maxIterZoom = c.max_iter_zoom;
        p.c1 = c.c1; p.c2 = c.c2; p.c3 = c.c3; p.method = c.cg_method == 0 ? "FR" : (c.cg_method == 1 ? "PR" : "PR+");
        sol = (method == SOH_NEWTON_CG) ? NewtonCG::minimize(*f->fn, start, p, &tr) : ConjugateGradient::minimize(*f->fn, start, p, &tr);
        break;
      }
      case SOH_STEIHAUG_CG: {
        SteihaugCGConfig p; p.tol = c.tol; p.maxIter = c.max_iter; p.eps = c.eps; p.delta0 = c.delta0; p.deltaMax = c.delta_max; p.eta = c.eta;
        sol = SteihaugCG::minimize(*f->fn, start, p, &tr);
        break;
      }
      case SOH_NELDER_MEAD: {
        NelderMeadConfig p; p.tol = c.tol; p.maxIter = c.max_iter; p.eps = c.eps; p.cReflection = c.c_reflection; p.cExpansion = c.c_expansion;
        p.cContraction = c.c_contraction; p.cShrink = c.c_shrink; p.relDelta = c.rel_delta; p.absDelta = c.abs_delta;
        sol = NelderMead::minimize(*f->fn, start, p, &tr);
        break;
      }
      default: throw IllegalArgumentException("unknown method");
    }
  });
  if (method != SOH_LEVENBERG_MARQUARDT) r.iterations = tr.iterations;
  if (res) *res = r;
  if (rc == SOH_OK) std::copy(sol.begin(), sol.end(), x_out);
  return rc;
}

extern "C" int soh_zoom_step_length(soh_objective* f, const double* x, const double* d, int n, double a, double b,
                                    const soh_config* cfg, double* x_out) {
  if (!f || !x || !d || !x_out || n != f->n) return SOH_ILLEGAL_ARG;
  soh_config c;
  if (cfg) c = *cfg; else soh_default_config(SOH_BFGS, &c);
  return guarded([&] {
    const Vec xv(x, x + n), dv(d, d + n);
    LineSearchPoint ptInit(xv, f->fn.get(), dv);
    LineSearchPoint pta = (a == 0.0) ? ptInit : ptInit.copyX(add(xv, mul(dv, a)));
    LineSearchPoint ptb = ptInit.copyX(add(xv, mul(dv, b)));
    const LineSearchPoint out = StrongWolfe::zoomStepLength(a, pta, b, ptb, ptInit, sw_of(c));
    std::copy(out.x.begin(), out.x.end(), x_out);
  });
}

extern "C" int soh_gpu_lm(so_ctx* ctx, so_dataset* ds, int f, int add_origin, int use_tsqr, double* beta) {
  if (!ctx || !ds || f < 1 || !beta) { g_err = "soh_gpu_lm: bad arguments"; return SOH_ILLEGAL_ARG; }
  return guarded([&] {
    const Vec b = gpu_lm(std::make_shared<GpuDataSet>(ctx, ds, f, false), add_origin != 0, use_tsqr != 0);
    std::copy(b.begin(), b.end(), beta);
  });
}

extern "C" int soh_qr(const double* ab, int64_t m, int n, int pivoting, double* R, double* qtb, int32_t* ipvt, double* acnorms,
                      double* bnorm, double* solution) {
  if (n < 1 || m < 0 || (m > 0 && !ab)) return SOH_ILLEGAL_ARG;
  return guarded([&] {
    std::vector<AugmentedRow> rows;
    for (int64_t i = 0; i < m; ++i) { const double* r = ab + (size_t)i * (n + 1); rows.emplace_back(Vec(r, r + n), r[n], i); }
    SeqDataSet<AugmentedRow> ds(std::move(rows));
    const QR qr = QR::apply(ds, n, pivoting != 0);
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) if (R) R[i * n + j] = qr.R(i, j);
    for (int i = 0; i < n; ++i) { if (qtb) qtb[i] = qr.qtb[i]; if (ipvt) ipvt[i] = qr.ipvt[i]; if (acnorms) acnorms[i] = qr.acNorms[i]; }
    if (bnorm) *bnorm = qr.bNorm;
    if (solution) { const Vec s = qr.solution(); std::copy(s.begin(), s.end(), solution); }
  });
}
