/*
 * oracle.h — CPU restatement of scalaopt's data-set evaluation path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this library.  The product (libscalaopt_cuda.so, libscalaopt_host.so, scalaopt_b200/) never does.
 *
 * Parity status: the reference is pure Scala and there is no JVM in this image or on the GPU box, so
 * the reference cannot be executed.  This restatement is pinned against the reference's own golden
 * vectors and known-answer tests where they exist (QRSpec.scala:33-43,66-71,120-125 via orc_qr, both pivoting
 * modes; java.util.Random(12345) draws; LinearSpec.scala:32-54; the LevenbergMarquardtSpec / BFGSSpec /
 * NewtonCGSpec / FFNeuralNetworkTrainerSpec / NelderMeadSpec recoveries with their iteration counts, driven through
 * the restated optimizers — tests/test_oracle.py, tests/test_host.py).
 * No reference test pins loss/gradient/JtJ/Hv to better than 1e-5; beyond that the parity target is
 * this file's line-by-line restatement, whose hand-derived derivative formulas are in turn checked
 * against independent 50-digit numerical differentiation of the models' definitions (tests/test_oracle_mp.py, 1e-13).
 *
 * All paths below are relative to /root/reference; core/... = core/src/main/scala/com/github/bruneli/scalaopt/core/...
 */
#ifndef SCALAOPT_ORACLE_H
#define SCALAOPT_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* same numbering as include/scalaopt_cuda.h (kept separate on purpose: the oracle shares no code
 * with the product) */
enum { ORC_LINEAR = 0, ORC_LINEAR_NOINT = 1, ORC_LOGISTIC = 2, ORC_EXPDECAY = 3, ORC_GAUSS = 4,
       ORC_GAUSS_EXPBKG = 5, ORC_POLY = 6,
       /* FFNeuralNetwork(Vector(f, h, 1)) with logistic inner neurons (FFNeuralNetwork.scala:114-189, Neuron.scala:40-73):
        * MSE loss + linear output / cross-entropy loss + logistic output; n = h (f + 1) + h + 1 weights in the
        * reference's order (hidden neurons, then the output neuron; bias first) */
       ORC_MLP_MSE = 7, ORC_MLP_CE = 8 };

/* fold order: parts <= 1 -> strict left fold (SeqDataSet.aggregate, core/.../SeqDataSetConverter.scala:32-35);
 * parts = N > 1 -> N contiguous partitions [i*m/N, (i+1)*m/N) folded independently from the zero element and
 * combined in partition order (deterministic model of rdd.aggregate on local[N],
 * spark-apps/.../SparkDataSetConverter.scala:42-45).  threads > 1 folds the partitions on that many pthreads. */
typedef struct orc_fold { int parts; int threads; } orc_fold;

/* ---- java.util.Random (48-bit LCG) so spec inputs with `new Random(12345)` can be regenerated ---- */
typedef struct orc_jrandom { uint64_t seed; int have_next; double next_gaussian; } orc_jrandom;
void   orc_jrandom_init(orc_jrandom* r, int64_t seed);
double orc_jrandom_next_double(orc_jrandom* r);
double orc_jrandom_next_gaussian(orc_jrandom* r);

/* ---- synthetic observations of SURVEY.md §8d (counter-based, same spec as so_dataset_generate) ---- */
void orc_generate(int family, int64_t m, int f, const double* p_true, int n, uint64_t seed, double noise_sigma,
                  int64_t row_offset, int64_t total_rows, double* X, double* y);

/* number of parameters of `family` on f inputs (-1 if the family takes a free n, i.e. POLY) */
int orc_family_nparams(int family, int f);

/* regression function f(p, x)  (RegressionFunction.scala:24-37; logistic: network output) */
double orc_model(int family, const double* p, int n, const double* x, int f);

/* MSEFunction.apply(p) = aggregate(0.0)(sum + loss)/size  (MSEFunction.scala:33-36,45-48,106-108);
 * logistic: FFNeuralNetworkTrainer.apply with decay 0 (FFNeuralNetworkTrainer.scala:59-62). */
double orc_apply(int family, const double* X, const double* y, int64_t m, int f, const double* p, int n, orc_fold fold);

/* analytic gradient as an aggregate of per-row gradients divided by size
 * (FFNeuralNetworkTrainer.scala:64-69,119-125 pattern). grad[n]. */
void orc_gradient(int family, const double* X, const double* y, int64_t m, int f, const double* p, int n,
                  orc_fold fold, double* grad);

/* SimpleMSEFunction.gradient: forward finite differences of apply, n+1 passes
 * (MSEFunction.scala:138-140, linalg/DenseVector.scala:269-275). */
void orc_gradient_fd(int family, const double* X, const double* y, int64_t m, int f, const double* p, int n,
                     double eps, orc_fold fold, double* grad);

/* one Jacobian row + residual, analytic: J = d residual / dp, residual = |y - f| (MSE families),
 * or the trainer's LM convention for the logistic family.  Jrow[n], returns residual. */
double orc_jacobian_row(int family, const double* p, int n, const double* x, int f, double y, double* Jrow);
/* same by the reference's forward finite difference (MSEFunction.scala:117-124). */
double orc_jacobian_row_fd(int family, const double* p, int n, const double* x, int f, double y, double eps, double* Jrow);

/* jacobianAndResidualsMatrix materialised: rows[m*(n+1)] = (J_i | r_i)  (MSEFunction.scala:132-136). fd != 0 -> FD rows */
void orc_jacobian_matrix(int family, const double* X, const double* y, int64_t m, int f, const double* p, int n,
                         int fd, double eps, double* rows);

/* sums that QR.apply extracts from that matrix: JtJ[n*n], Jtr[n], rtr (folded row by row in `fold` order). */
/* FFNeuralNetwork(Vector(sizes...)) with vector targets (Y is m x sizes[nlayers-1], row-major): one row, the trainer's loss /
 * gradient with weight decay, and its Jacobian/residual rows incl. the decay penalty rows.  -1: a shape the reference itself
 * refuses (mean squared error with several outputs is `???` there). */
int orc_net_row(const int* sizes, int nlayers, int cross_entropy, const double* w, const double* x, const double* y,
                double* loss, double* grad, double* jac, double* residual);
int orc_net_loss_grad(const int* sizes, int nlayers, int cross_entropy, const double* X, const double* Y, int64_t m,
                      const double* w, double decay, double* loss, double* grad);
int orc_net_jacobian_matrix(const int* sizes, int nlayers, int cross_entropy, const double* X, const double* Y, int64_t m,
                            const double* w, double decay, double* rows);
/* compensated (Neumaier + exact products) version of the same sums: hi + lo, n*n + n + 1 doubles each; test yardstick only */
void orc_normal_eq_comp(int family, const double* X, const double* y, int64_t m, int f, const double* p, int n, int threads,
                        double* out_hi, double* out_lo);
void orc_normal_eq(int family, const double* X, const double* y, int64_t m, int f, const double* p, int n,
                   orc_fold fold, double* JtJ, double* Jtr, double* rtr);

/* QR.apply restated pass for pass over a materialised augmented matrix ab[m*(n+1)] (linalg/QR.scala:131-312).
 * Outputs: r[n*n] (rows of the reduced matrix, strict upper part meaningful), rdiag[n], qtb[n], ipvt[n],
 * acnorms[n], *bnorm.  Returns 0, or -1 when the reference would throw IllegalArgumentException
 * (fewer than n rows: QR.scala:45-46). */
int orc_qr(const double* ab, int64_t m, int n, int pivoting, double* r, double* rdiag, double* qtb, int* ipvt,
           double* acnorms, double* bnorm);
/* QR.solution (QR.scala:97-114) */
void orc_qr_solution(const double* r, const double* rdiag, const double* qtb, const int* ipvt, int n, double* x);

/* LineSearchPoint.fx / dfx at x + alpha d (variable/Point.scala:54-57; StrongWolfe.scala:97-98):
 * phi[j] = apply(p + alpha_j d), dphi[j] = gradient(p + alpha_j d) . d  (ObjectiveFunctionWithGradient.scala:37-39) */
void orc_line(int family, const double* X, const double* y, int64_t m, int f, const double* p, const double* d, int n,
              const double* alpha, int k, orc_fold fold, double* phi, double* dphi);

/* dirHessian by the reference's finite difference of gradients (ObjectiveFunctionWithGradient.scala:41-47) */
void orc_hess_vec_fd(int family, const double* X, const double* y, int64_t m, int f, const double* p, const double* d,
                     int n, double eps, orc_fold fold, double* Hd);
/* exact Hessian-vector product (analytic second derivatives), aggregated row by row */
void orc_hess_vec(int family, const double* X, const double* y, int64_t m, int f, const double* p, const double* d,
                  int n, orc_fold fold, double* Hd);

/* ---- negative log-likelihood objective ----
 * negLogLikelihood(pdf, data)(p) = data.aggregate(0.0)((sum, x) => sum - 2.0 * Math.log(pdf(p, x)(0)), _ + _)
 * (std-apps/.../stats/package.scala:97-103; NOT divided by the data-set size), with the pdf of
 * std-apps/.../stats/example/ExSignalPlusBackgroundFit.scala:78-114:
 *   ORC_PDF_GAUSS_EXPBKG: p = (fSig, mu, sigma, lambda), consts = (xMin):
 *   pdf = N(x; mu, sigma) * fSig + lambda exp(-lambda (x - xMin)) * (1 - fSig) */
enum { ORC_PDF_GAUSS_EXPBKG = 0 };
double orc_pdf(int pdf, const double* p, int n, const double* consts, double x);
double orc_nll(int pdf, const double* x, int64_t m, const double* p, int n, const double* consts, orc_fold fold);
/* the example's own sample (ExSignalPlusBackgroundFit.scala:39-55): java.util.Random(seed), nbEvents variates */
void orc_nll_example_data(int64_t seed, int nb_events, double f_signal, double mu, double sigma, double lambda,
                          double x_min, double* x);

/* faithful multi-pass LM outer-iteration cost model for CPU timing: FD Jacobian matrix + 2n+1-pass QR + loss pass.
 * Returns bNorm. (LevenbergMarquardt.scala:124,164) */
double orc_lm_reference_passes(int family, const double* X, const double* y, int64_t m, int f, const double* p, int n,
                               int threads);

#ifdef __cplusplus
}
#endif
#endif
