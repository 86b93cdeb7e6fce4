/*
 * scalaopt_host.h — C ABI of libscalaopt_host.so: the host-side mirror of the reference's optimizer-facing
 * interface (scalaopt_b200/host/scalaopt_host.hpp, gpu.hpp) made callable from C / ctypes / JNI.
 *
 * The reference's callers of the hot path are Scala objects with
 *     minimize(f, x0)(implicit pars): Try[DenseVectorLike[A]]        core/.../Optimizer.scala:45
 * This ABI exposes the same five minimizers over an opaque objective handle that is either GPU-backed
 * (GpuMSEFunction over a so_dataset) or built from C callbacks (how the tests plug the CPU oracle or a
 * closure such as BFGSSpec's quadratic in).  Paths: core/... = core/src/main/scala/com/github/bruneli/scalaopt/core/...
 */
#ifndef SCALAOPT_HOST_H
#define SCALAOPT_HOST_H
#include <stdint.h>
#include "scalaopt_cuda.h"
#ifdef __cplusplus
extern "C" {
#endif

/* return codes of soh_minimize / soh_zoom_step_length */
#define SOH_OK            0   /* Success(x) */
#define SOH_MAX_ITER      1   /* MaxIterException was thrown (core/.../MaxIterException.scala) */
#define SOH_ILLEGAL_ARG   2   /* IllegalArgumentException (a failed require(...)) */
#define SOH_GPU_ERROR     3   /* an so_* call failed; see soh_last_error */
#define SOH_ERROR         4   /* anything else */

typedef enum soh_method {
  SOH_LEVENBERG_MARQUARDT = 0,  /* leastsquares/LevenbergMarquardt.scala:111-226 (needs an MSE objective) */
  SOH_BFGS = 1,                 /* gradient/BFGS.scala:57-110 */
  SOH_CONJUGATE_GRADIENT = 2,   /* gradient/ConjugateGradient.scala:53-90 */
  SOH_NEWTON_CG = 3,            /* gradient/NewtonCG.scala:55-82 */
  SOH_STEIHAUG_CG = 4,          /* gradient/SteihaugCG.scala:51-91 */
  SOH_NELDER_MEAD = 5           /* derivativefree/NelderMead.scala:45-65 (default method of stats.maxLikelihoodFit) */
} soh_method;

/* union of the reference's config case classes; soh_default_config fills the reference's defaults:
 * ConfigPars (Optimizer.scala:56-63), BFGSConfig (BFGS.scala:125-146), CGConfig (ConjugateGradient.scala:119-131),
 * SteihaugCGConfig (SteihaugCG.scala:162-168), LevenbergMarquardtConfig (LevenbergMarquardt.scala:497-508),
 * StrongWolfeConfig (StrongWolfe.scala:48-56) */
typedef struct soh_config {
  double tol; int32_t max_iter; double eps;
  int32_t max_iter_line, max_iter_zoom, max_iter_first_time /* < 0: None */;
  double c1, c2, c3;
  int32_t cg_method;              /* 0 = "FR", 1 = "PR", 2 = "PR+" */
  double delta0, delta_max, eta;  /* Steihaug */
  double x_tol, g_tol, ratio_tol, step_bound, epsmch; int32_t max_inner_iter, max_step_length_iter, use_pivoting;   /* LM */
  double c_reflection, c_expansion, c_contraction, c_shrink, rel_delta, abs_delta;   /* NelderMeadConfig (NelderMead.scala:307-314) */
} soh_config;
int soh_default_config(int method, soh_config* out);

typedef struct soh_result {
  int32_t iterations;        /* outer iterations performed */
  int32_t inner_iterations;  /* LM only */
  int32_t gauss_newton_steps, damped_steps;   /* LM only: which branch of lmPar each step took */
} soh_result;

/* what an objective was asked and what it made the device do */
typedef struct soh_counters {
  int64_t loss_calls, gradient_calls, dirder_calls, hessvec_calls, jacobian_calls;
  int64_t k1_passes, k2_passes, k3_passes, k4_passes;
  int64_t cache_hits, line_hits;
  int64_t spec_passes, spec_hits;   /* speculative K2 passes at LM trial points, Jacobians served from them */
} soh_counters;

typedef struct soh_objective soh_objective;

/* callbacks (user = the pointer given at creation) */
typedef double (*soh_apply_fn)(void* user, const double* x, int n);
typedef void   (*soh_gradient_fn)(void* user, const double* x, int n, double* grad_out);
typedef void   (*soh_hessvec_fn)(void* user, const double* x, const double* d, int n, double* out);
/* fills rows[m*(n+1)] with (J_i | r_i): MSEFunction.jacobianAndResidualsMatrix (function/MSEFunction.scala:74) */
typedef void   (*soh_jacobian_fn)(void* user, const double* p, int n, double* rows, int64_t m);

/* (f, df) closure pair -> ObjectiveFunctionWithGradient (function/ObjectiveFunctionWithGradient.scala:29-49; dirHessian is
 * its finite difference of gradients unless `hessvec` is given); gradient == NULL -> ObjectiveFunctionFiniteDiffDerivatives
 * (:30-52).  jacobian != NULL (with its row count m) additionally makes it an MSEFunction for Levenberg-Marquardt. */
soh_objective* soh_objective_from_callbacks(int n, soh_apply_fn apply, soh_gradient_fn gradient, soh_hessvec_fn hessvec,
                                            soh_jacobian_fn jacobian, int64_t m, void* user, double eps);

/* GpuMSEFunction(family, GpuDataSet(ds)): the drop-in for SimpleMSEFunction (function/MSEFunction.scala:85-147) and for
 * FFNeuralNetworkTrainer's logistic case.  The data set must outlive the objective. */
soh_objective* soh_objective_gpu(so_ctx* ctx, so_dataset* ds, int f, int family, int n);

/* GpuMSEFunction.jacobianAndResidualsMatrix from so_eval_qr (on-device Householder TSQR of [J | r])
 * instead of so_eval_normal_eq + Cholesky: one pass either way, about 2.5x the kernel time at n = 5, and the condition
 * number of J is not squared.  Single-input families; linear (+- intercept) and logistic rows up to n = 63. */
int soh_objective_use_tsqr(soh_objective* objective, int on);

/* Levenberg-Marquardt trial-point losses from a K2 pass whose sums are kept for the Jacobian request that follows an
 * accepted step (default on; off = one K1 loss pass per trial point + one K2 pass per outer iteration). */
/* FFNeuralNetworkTrainer.withDecay (std-apps/.../nnet/FFNeuralNetworkTrainer.scala:57): weight decay of a GPU network objective */
int soh_objective_set_decay(soh_objective* objective, double decay);
int soh_objective_speculate(soh_objective* objective, int on);

/* negLogLikelihood(pdf, GpuDataSet(ds)) _ (std-apps/.../stats/package.scala:97-103) with finite-difference derivatives
 * (what the reference's implicit `toFunctionWoGradient` gives a plain closure, core/.../package.scala:74-77). */
soh_objective* soh_objective_gpu_nll(so_ctx* ctx, so_dataset* ds, int pdf, int n, const double* consts, int nconsts);

void soh_objective_free(soh_objective* f);
int  soh_objective_counters(soh_objective* f, soh_counters* out);
/* direct access to the objective-function interface (DifferentiableObjectiveFunction.scala:28-60) */
int  soh_objective_apply(soh_objective* f, const double* x, double* out);
int  soh_objective_gradient(soh_objective* f, const double* x, double* grad_out);
int  soh_objective_dirder(soh_objective* f, const double* x, const double* d, double* out);
int  soh_objective_dir_hessian(soh_objective* f, const double* x, const double* d, double* out);
/* rows of jacobianAndResidualsMatrix(p): *m_out rows of n+1 doubles copied into rows (capacity in rows); n+1 rows for a GPU objective */
int  soh_objective_jacobian_rows(soh_objective* f, const double* p, double* rows, int64_t capacity_rows, int64_t* m_out);

int soh_minimize(int method, soh_objective* f, const double* x0, int n, const soh_config* cfg, double* x_out, soh_result* res);

/* StrongWolfe.zoomStepLength(a, ptInit.copy(x + a d), b, ptInit.copy(x + b d), ptInit) (linesearch/StrongWolfe.scala:132-230) */
int soh_zoom_step_length(soh_objective* f, const double* x, const double* d, int n, double a, double b,
                         const soh_config* cfg, double* x_out);

/* QR.apply over an in-memory augmented matrix ab[m*(n+1)] (linalg/QR.scala:131-152) + QR.solution (:97-114).
 * R is n*n row-major with R(i,j) of QR.scala:55-67. */
int soh_qr(const double* ab, int64_t m, int n, int pivoting, double* R, double* qtb, int32_t* ipvt, double* acnorms,
           double* bnorm, double* solution);

/* linear.lm(GpuDataSet(ds), addOrigin) (linear/package.scala:54-63): least-squares solution beta[f + addOrigin] of
 * (1, x) beta = y through the reference's QR(...).solution on the (n+1)-row system of one K2 pass (use_tsqr: of the
 * on-device Householder TSQR, f + addOrigin <= 63).  Failure(e) of the reference's Try = a non-zero return code. */
int soh_gpu_lm(so_ctx* ctx, so_dataset* ds, int f, int add_origin, int use_tsqr, double* beta);

const char* soh_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
