/*
 * oracle.c — CPU restatement of scalaopt's data-set evaluation path.  TEST INFRASTRUCTURE ONLY
 * (see oracle.h for who may load it and for the parity status).
 *
 * Plain C99 + libm + pthreads, compiled -O2 without -ffast-math / -ffp-contract so that sums are
 * evaluated in the written order, like the JVM evaluates the reference.
 *
 * Reference paths are relative to /root/reference; core/... =
 * core/src/main/scala/com/github/bruneli/scalaopt/core/...
 */
#include "oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define ORC_MAXN 1024

/* =============================================================================================
 * java.util.Random — JDK LCG; scala.util.Random(12345) wraps it (every spec uses seed 12345,
 * e.g. core/src/test/.../leastsquares/LevenbergMarquardtSpec.scala:34)
 * =========================================================================================== */
static const uint64_t JR_MULT = 0x5DEECE66DULL;
static const uint64_t JR_MASK = (1ULL << 48) - 1;

void orc_jrandom_init(orc_jrandom* r, int64_t seed) {
  r->seed = ((uint64_t)seed ^ JR_MULT) & JR_MASK;
  r->have_next = 0;
  r->next_gaussian = 0.0;
}
static int32_t jr_next(orc_jrandom* r, int bits) {
  r->seed = (r->seed * JR_MULT + 0xBULL) & JR_MASK;
  return (int32_t)((int64_t)r->seed >> (48 - bits));
}
double orc_jrandom_next_double(orc_jrandom* r) {
  int64_t hi = (int64_t)jr_next(r, 26);
  int64_t lo = (int64_t)jr_next(r, 27);
  return (double)((hi << 27) + lo) * 0x1.0p-53;
}
double orc_jrandom_next_gaussian(orc_jrandom* r) {
  if (r->have_next) { r->have_next = 0; return r->next_gaussian; }
  double v1, v2, s;
  do {
    v1 = 2.0 * orc_jrandom_next_double(r) - 1.0;
    v2 = 2.0 * orc_jrandom_next_double(r) - 1.0;
    s = v1 * v1 + v2 * v2;
  } while (s >= 1.0 || s == 0.0);
  double mul = sqrt(-2.0 * log(s) / s);   /* StrictMath in the JDK; libm here */
  r->next_gaussian = v2 * mul;
  r->have_next = 1;
  return v1 * mul;
}

/* =============================================================================================
 * Model families (regression functions).  The reference passes these as Scala closures
 * (function/RegressionFunction.scala:24-37); the catalogue below follows the closures of its
 * own tests/examples where they exist.
 * =========================================================================================== */
int orc_family_nparams(int family, int f) {
  switch (family) {
    case ORC_LINEAR: case ORC_LOGISTIC: return f + 1;
    case ORC_LINEAR_NOINT: return f;
    case ORC_EXPDECAY: return 2;
    case ORC_GAUSS: return 3;
    case ORC_GAUSS_EXPBKG: return 5;
    default: return -1;
  }
}

/* DenseVectorLike.inner: acc from 0.0, left to right (linalg/DenseVectorLike.scala:68-76) */
static double inner(const double* a, const double* b, int len) {
  double acc = 0.0;
  for (int i = 0; i < len; ++i) acc += a[i] * b[i];
  return acc;
}

static void mlp_backprop(int family, const double* p, int n, const double* x, int f, double y,
                         double* grad, double* residual, double* loss, double* output);

double orc_model(int family, const double* p, int n, const double* x, int f) {
  if (family == ORC_MLP_MSE || family == ORC_MLP_CE) {
    double out;
    mlp_backprop(family, p, n, x, f, 0.0, NULL, NULL, NULL, &out);
    return out;
  }
  switch (family) {
    case ORC_LINEAR:        /* p(0) + p.tail.inner(x): LevenbergMarquardtSpec.scala:38-40 */
      return p[0] + inner(p + 1, x, f);
    case ORC_LINEAR_NOINT:
      return inner(p, x, f);
    case ORC_LOGISTIC: {    /* Neuron.activate: weights(0) + (inputs dot weights.tail), Neuron.scala:40-49;
                               LogisticFunction.apply, activation/LogisticFunction.scala:24 */
      double excitation = p[0] + inner(x, p + 1, f);
      return 1.0 / (1.0 + exp(-excitation));
    }
    case ORC_EXPDECAY:      /* x(0) * exp(x(1) * t(0)): LevenbergMarquardtSpec.scala:62-63 */
      return p[0] * exp(p[1] * x[0]);
    case ORC_GAUSS: {
      double d = x[0] - p[1];
      return p[0] * exp(-(d * d) / (2.0 * p[2] * p[2]));
    }
    case ORC_GAUSS_EXPBKG: { /* Gaussian signal + exponential background, ExSignalPlusBackgroundFit.scala:78-114 */
      double d = x[0] - p[1];
      return p[0] * exp(-(d * d) / (2.0 * p[2] * p[2])) + p[3] * exp(-p[4] * x[0]);
    }
    case ORC_POLY: {
      double acc = 0.0, tk = 1.0;
      for (int k = 0; k < n; ++k) { acc += p[k] * tk; tk *= x[0]; }
      return acc;
    }
    default: return NAN;
  }
}

/* d f / d p_j, analytic.  df[n] */
static void model_grad(int family, const double* p, int n, const double* x, int f, double* df) {
  switch (family) {
    case ORC_LINEAR:
    case ORC_LOGISTIC:   /* for the logistic family this is d excitation / dp = (1 +: inputs), Neuron.scala:65 */
      df[0] = 1.0;
      for (int j = 0; j < f; ++j) df[j + 1] = x[j];
      break;
    case ORC_LINEAR_NOINT:
      for (int j = 0; j < f; ++j) df[j] = x[j];
      break;
    case ORC_EXPDECAY: {
      double e = exp(p[1] * x[0]);
      df[0] = e;
      df[1] = p[0] * x[0] * e;
      break;
    }
    case ORC_GAUSS:
    case ORC_GAUSS_EXPBKG: {
      double A = p[0], mu = p[1], sg = p[2];
      double d = x[0] - mu;
      double g = exp(-(d * d) / (2.0 * sg * sg));
      df[0] = g;
      df[1] = A * g * d / (sg * sg);
      df[2] = A * g * d * d / (sg * sg * sg);
      if (family == ORC_GAUSS_EXPBKG) {
        double e = exp(-p[4] * x[0]);
        df[3] = e;
        df[4] = -p[3] * x[0] * e;
      }
      break;
    }
    case ORC_POLY: {
      double tk = 1.0;
      for (int k = 0; k < n; ++k) { df[k] = tk; tk *= x[0]; }
      break;
    }
    default: break;
  }
}

/* (d^2 f / dp^2) . d, analytic, through the explicit second-derivative matrix.  out[n] */
static void model_hess_dir(int family, const double* p, int n, const double* x, int f, const double* d, double* out) {
  (void)f;
  for (int j = 0; j < n; ++j) out[j] = 0.0;
  switch (family) {
    case ORC_EXPDECAY: {
      double t = x[0], e = exp(p[1] * t);
      double H01 = t * e, H11 = p[0] * t * t * e;
      out[0] = H01 * d[1];
      out[1] = H01 * d[0] + H11 * d[1];
      break;
    }
    case ORC_GAUSS:
    case ORC_GAUSS_EXPBKG: {
      double A = p[0], mu = p[1], s = p[2], t = x[0];
      double dl = t - mu;
      double g = exp(-(dl * dl) / (2.0 * s * s));
      double s2 = s * s, s3 = s2 * s, s4 = s2 * s2, s5 = s4 * s, s6 = s3 * s3;
      double HAm = g * dl / s2, HAs = g * dl * dl / s3;
      double Hmm = A * g * (dl * dl / s4 - 1.0 / s2);
      double Hms = A * g * (dl * dl * dl / s5 - 2.0 * dl / s3);
      double Hss = A * g * (dl * dl * dl * dl / s6 - 3.0 * dl * dl / s4);
      out[0] = HAm * d[1] + HAs * d[2];
      out[1] = HAm * d[0] + Hmm * d[1] + Hms * d[2];
      out[2] = HAs * d[0] + Hms * d[1] + Hss * d[2];
      if (family == ORC_GAUSS_EXPBKG) {
        double e = exp(-p[4] * t);
        double HBl = -t * e, Hll = p[3] * t * t * e;
        out[3] = HBl * d[4];
        out[4] = HBl * d[3] + Hll * d[4];
      }
      break;
    }
    default: break;   /* linear in the parameters: zero */
  }
}

/* SimpleMSEFunction.residual: (data.y - f(x, data.x)).norm, MSEFunction.scala:106-108 (scalar y: sqrt(d*d)) */
static double mse_residual(int family, const double* p, int n, const double* x, int f, double y) {
  double d = y - orc_model(family, p, n, x, f);
  return sqrt(d * d);
}

/* FFNeuralNetwork.loss for one output neuron, LossType.CrossEntropy, FFNeuralNetwork.scala:80-88 */
static double logistic_ce(double output, double target) {
  double entropy1 = (target > 0.0) ? -target * log(output / target) : 0.0;
  double entropy0 = (target < 1.0) ? -(1.0 - target) * log((1.0 - output) / (1.0 - target)) : 0.0;
  return entropy0 + entropy1;
}

/* =============================================================================================
 * DataSet.aggregate — fold machinery (DataSet.scala:49)
 * =========================================================================================== */
typedef void (*row_fn)(const void* ctx, const double* x, double y, int64_t i, double* acc);

typedef struct fold_job {
  row_fn fn; const void* ctx; const double* X; const double* y; int f; int Q;
  int64_t m; int parts; int first_part, last_part; double* accs;   /* accs[parts*Q] */
} fold_job;

static void* fold_worker(void* arg) {
  fold_job* jb = (fold_job*)arg;
  for (int pidx = jb->first_part; pidx < jb->last_part; ++pidx) {
    /* ParallelCollectionRDD.slice: [i*m/N, (i+1)*m/N) */
    int64_t b = (int64_t)(((__int128)pidx * jb->m) / jb->parts);
    int64_t e = (int64_t)(((__int128)(pidx + 1) * jb->m) / jb->parts);
    double* acc = jb->accs + (size_t)pidx * jb->Q;
    for (int q = 0; q < jb->Q; ++q) acc[q] = 0.0;
    for (int64_t i = b; i < e; ++i) jb->fn(jb->ctx, jb->X + (size_t)i * jb->f, jb->y[i], i, acc);
  }
  return NULL;
}

static void fold_rows(row_fn fn, const void* ctx, const double* X, const double* y, int64_t m, int f, int Q,
                      orc_fold fold, double* acc) {
  int parts = fold.parts < 1 ? 1 : fold.parts;
  int threads = fold.threads < 1 ? 1 : fold.threads;
  if (threads > parts) threads = parts;
  double* accs = (double*)malloc(sizeof(double) * (size_t)parts * Q);
  fold_job base = { fn, ctx, X, y, f, Q, m, parts, 0, parts, accs };
  if (threads == 1) {
    fold_worker(&base);
  } else {
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * threads);
    fold_job* jobs = (fold_job*)malloc(sizeof(fold_job) * threads);
    for (int t = 0; t < threads; ++t) {
      jobs[t] = base;
      jobs[t].first_part = (int)(((int64_t)t * parts) / threads);
      jobs[t].last_part = (int)(((int64_t)(t + 1) * parts) / threads);
      pthread_create(&th[t], NULL, fold_worker, &jobs[t]);
    }
    for (int t = 0; t < threads; ++t) pthread_join(th[t], NULL);
    free(th); free(jobs);
  }
  /* combop in partition order starting from the zero element */
  for (int q = 0; q < Q; ++q) acc[q] = 0.0;
  for (int pidx = 0; pidx < parts; ++pidx)
    for (int q = 0; q < Q; ++q) acc[q] = acc[q] + accs[(size_t)pidx * Q + q];
  free(accs);
}

/* ---------------------------------------------------------------------------------------------
 * FFNeuralNetwork(Vector(f, h, 1)).forward(x).backward(y), one row.
 * forward: Neuron.activate, excitation = weights(0) + (inputs dot weights.tail) (Neuron.scala:40-49);
 *   inner neurons LogisticFunction, output neuron LinearFunction (MSE) or LogisticFunction (cross-entropy).
 * backward, outer layer (FFNeuralNetwork.scala:159-181): MSE: Neuron.propagateError(target, f): lossDerivative =
 *   output - target, error = lossDerivative * f.derivative(output), residual = lossDerivative (Neuron.scala:51-56);
 *   cross-entropy, one output: residual = output - target, error = residual.
 * inner layer (Neuron.scala:58-63): lossDerivative = sum over next layer of error * weights(index + 1),
 *   error = lossDerivative * output (1 - output), residual = next.head.residual.
 * gradient = (1 +: inputs) * error per neuron, layer by layer (Neuron.scala:65, FFNeuralNetwork.scala:62-64);
 * loss = residual^2 (MSE: norm2 of the output residuals, :70-72) or the cross-entropy of :73-81.
 * ------------------------------------------------------------------------------------------- */
static int mlp_hidden(int n, int f) { return (n - 1) / (f + 2); }

static void mlp_backprop(int family, const double* p, int n, const double* x, int f, double y,
                         double* grad, double* residual, double* loss, double* output) {
  int h = mlp_hidden(n, f);
  double o[ORC_MAXN];
  for (int j = 0; j < h; ++j) {
    const double* w = p + (size_t)j * (f + 1);
    double excitation = w[0] + inner(x, w + 1, f);
    o[j] = 1.0 / (1.0 + exp(-excitation));
  }
  const double* v = p + (size_t)h * (f + 1);
  double excitation = v[0] + inner(o, v + 1, h);
  double out = (family == ORC_MLP_CE) ? 1.0 / (1.0 + exp(-excitation)) : excitation;
  double res = out - y;
  double error = (family == ORC_MLP_CE) ? res : res * 1.0;
  if (grad) {
    for (int j = 0; j < h; ++j) {
      double loss_derivative = error * v[j + 1];
      double ej = loss_derivative * (o[j] * (1.0 - o[j]));
      double* g = grad + (size_t)j * (f + 1);
      g[0] = 1.0 * ej;
      for (int i = 0; i < f; ++i) g[1 + i] = x[i] * ej;
    }
    double* g = grad + (size_t)h * (f + 1);
    g[0] = 1.0 * error;
    for (int j = 0; j < h; ++j) g[1 + j] = o[j] * error;
  }
  if (residual) *residual = res;
  if (output) *output = out;
  if (loss) *loss = (family == ORC_MLP_CE) ? logistic_ce(out, y) : res * res;
}

/* ---------------------------------------------------------------------------------------------
 * The general network: FFNeuralNetwork(Vector(sizes...)) with vector targets, LossType.MeanSquaredError + LinearFunction
 * output or LossType.CrossEntropy + LogisticFunction (one output) / SoftMaxFunction (several), inner neurons
 * LogisticFunction — what FFNeuralNetworkTrainerSpec.scala:222-260 (Vector(2, 3)) and :353-376 (Vector(4, 5, 3)) train.
 * Weights in the reference's order: layer by layer, neuron by neuron, bias first (FFNeuralNetwork.scala:66-68,252-263).
 *   forward  FFNeuralNetwork.forward / activateLayer (:112-158): soft-max with the maximum excitation subtracted
 *   backward propagateErrorsOuterLayer (:161-187), Neuron.propagateError (Neuron.scala:51-63)
 *   loss     :76-104 (the multi-class fold returns 0.0 when it meets an output that is exactly 0 — restated as is)
 *   residual :105-110; gradient Neuron.scala:65; jacobian Neuron.scala:67-73
 * Mean squared error with several outputs is `???` in the reference (:157): refused here too (returns -1).
 * ------------------------------------------------------------------------------------------- */
#define ORC_NET_MAXL 8
#define ORC_NET_MAXW 64
int orc_net_row(const int* sizes, int nlayers, int cross_entropy, const double* w, const double* x, const double* y,
                double* loss, double* grad, double* jac, double* residual) {
  double out[ORC_NET_MAXL][ORC_NET_MAXW], err[ORC_NET_MAXL][ORC_NET_MAXW], res[ORC_NET_MAXL][ORC_NET_MAXW];
  const double* wl[ORC_NET_MAXL];
  int K = sizes[nlayers - 1];
  if (!cross_entropy && K > 1) return -1;
  const double* wp = w;
  for (int l = 1; l < nlayers; ++l) { wl[l] = wp; wp += (size_t)sizes[l] * (sizes[l - 1] + 1); }
  for (int i = 0; i < sizes[0]; ++i) out[0][i] = x[i];
  for (int l = 1; l < nlayers; ++l) {
    int nin = sizes[l - 1], last = (l == nlayers - 1);
    double exc[ORC_NET_MAXW];
    for (int j = 0; j < sizes[l]; ++j) {
      const double* wj = wl[l] + (size_t)j * (nin + 1);
      exc[j] = wj[0] + inner(out[l - 1], wj + 1, nin);                       /* Neuron.activate */
    }
    if (last && K > 1) {                                                      /* activateLayer, CrossEntropy, several outputs */
      double mx = exc[0], sum = 0.0, probs[ORC_NET_MAXW];
      for (int j = 1; j < K; ++j) if (exc[j] > mx) mx = exc[j];
      for (int j = 0; j < K; ++j) { probs[j] = exp(exc[j] - mx); sum = sum + probs[j]; }
      for (int j = 0; j < K; ++j) out[l][j] = probs[j] / sum;
    } else {
      for (int j = 0; j < sizes[l]; ++j)
        out[l][j] = (last && !cross_entropy) ? exc[j] : 1.0 / (1.0 + exp(-exc[j]));
    }
  }
  int L = nlayers - 1;
  /* outer layer */
  double targets_sum = 0.0;
  for (int k = 0; k < K; ++k) targets_sum = targets_sum + y[k];
  for (int k = 0; k < K; ++k) {
    if (!cross_entropy) { res[L][k] = out[L][k] - y[k]; err[L][k] = res[L][k] * 1.0; }
    else if (K > 1) { res[L][k] = targets_sum * out[L][k] - y[k]; err[L][k] = res[L][k]; }
    else { res[L][k] = out[L][k] - y[k]; err[L][k] = res[L][k]; }
  }
  /* inner layers, last to first */
  for (int l = L - 1; l >= 1; --l)
    for (int j = 0; j < sizes[l]; ++j) {
      double ld = 0.0;
      for (int k = 0; k < sizes[l + 1]; ++k) ld = ld + err[l + 1][k] * wl[l + 1][(size_t)k * (sizes[l] + 1) + 1 + j];
      err[l][j] = ld * (out[l][j] * (1.0 - out[l][j]));
      res[l][j] = res[l + 1][0];                                              /* residual = neurons.head.residual */
    }
  if (loss) {
    if (!cross_entropy) { double s2 = 0.0; for (int k = 0; k < K; ++k) s2 = s2 + res[L][k] * res[L][k]; *loss = s2; }
    else if (K == 1) *loss = logistic_ce(out[L][0], y[0]);
    else {
      double osum = 0.0, entropy = 0.0;
      for (int k = 0; k < K; ++k) osum = osum + out[L][k];
      if (osum > 0.0) {
        for (int k = 0; k < K; ++k) entropy = (out[L][k] > 0.0) ? entropy - y[k] * log(out[L][k] / osum) : 0.0;
        *loss = entropy;
      } else *loss = 0.0;
    }
  }
  if (residual) {
    if (K == 1) *residual = res[L][0];
    else { double s2 = 0.0; for (int k = 0; k < K; ++k) s2 = s2 + res[L][k] * res[L][k]; *residual = sqrt(s2); }
  }
  if (grad || jac) {
    size_t q = 0;
    for (int l = 1; l < nlayers; ++l)
      for (int j = 0; j < sizes[l]; ++j)
        for (int i = 0; i <= sizes[l - 1]; ++i) {
          double g = (i == 0 ? 1.0 : out[l - 1][i - 1]) * err[l][j];
          if (grad) grad[q] = g;
          if (jac) jac[q] = (res[l][j] != res[l][j] || res[l][j] == 0.0) ? g : g / res[l][j];
          ++q;
        }
  }
  return 0;
}

static int net_nweights(const int* sizes, int nlayers) {
  int n = 0;
  for (int l = 1; l < nlayers; ++l) n += sizes[l] * (sizes[l - 1] + 1);
  return n;
}

/* FFNeuralNetworkTrainer.apply / gradient (Trainer.scala:59-69): the fold starts from the weight-decay term and is divided
 * by data.size; Y is m x K row-major */
int orc_net_loss_grad(const int* sizes, int nlayers, int cross_entropy, const double* X, const double* Y, int64_t m,
                      const double* w, double decay, double* loss, double* grad) {
  int n = net_nweights(sizes, nlayers), f = sizes[0], K = sizes[nlayers - 1];
  double w2 = 0.0;
  for (int i = 0; i < n; ++i) w2 = w2 + w[i] * w[i];
  double acc = w2 / 2.0 * decay, g[ORC_NET_MAXL * ORC_NET_MAXW * 4];
  if (n > (int)(sizeof g / sizeof g[0])) return -1;
  if (grad) for (int i = 0; i < n; ++i) grad[i] = w[i] * decay;
  for (int64_t r = 0; r < m; ++r) {
    double l;
    if (orc_net_row(sizes, nlayers, cross_entropy, w, X + (size_t)r * f, Y + (size_t)r * K, &l, grad ? g : NULL, NULL, NULL)) return -1;
    acc = acc + l;
    if (grad) for (int i = 0; i < n; ++i) grad[i] = grad[i] + g[i];
  }
  *loss = acc / (double)m;
  if (grad) for (int i = 0; i < n; ++i) grad[i] = grad[i] / (double)m;
  return 0;
}

/* jacobianAndResidualsMatrix (Trainer.scala:100-108,147-153): m rows (jacobian | residual), then, with decay > 0, n penalty rows
 * sqrt(decay) e_i | 0.  rows: (m [+ n]) x (n + 1) */
int orc_net_jacobian_matrix(const int* sizes, int nlayers, int cross_entropy, const double* X, const double* Y, int64_t m,
                            const double* w, double decay, double* rows) {
  int n = net_nweights(sizes, nlayers), f = sizes[0], K = sizes[nlayers - 1];
  for (int64_t r = 0; r < m; ++r)
    if (orc_net_row(sizes, nlayers, cross_entropy, w, X + (size_t)r * f, Y + (size_t)r * K, NULL, NULL, rows + (size_t)r * (n + 1),
                    rows + (size_t)r * (n + 1) + n)) return -1;
  if (decay > 0.0)
    for (int i = 0; i < n; ++i) {
      double* row = rows + (size_t)(m + i) * (n + 1);
      for (int k = 0; k <= n; ++k) row[k] = 0.0;
      row[i] = sqrt(decay);
    }
  return 0;
}

typedef struct eval_ctx { int family; const double* p; int n; int f; const double* d; double eps; } eval_ctx;

/* ---------------------------------------------------------------------------------------------
 * apply(p)
 * ------------------------------------------------------------------------------------------- */
static void loss_row(const void* vctx, const double* x, double y, int64_t i, double* acc) {
  const eval_ctx* c = (const eval_ctx*)vctx; (void)i;
  if (c->family == ORC_LOGISTIC) {
    /* FFNeuralNetworkTrainer.loss: zero + network.forward(x).backward(y).loss, Trainer.scala:133-135 */
    acc[0] = acc[0] + logistic_ce(orc_model(ORC_LOGISTIC, c->p, c->n, x, c->f), y);
  } else if (c->family == ORC_MLP_MSE || c->family == ORC_MLP_CE) {
    double l;
    mlp_backprop(c->family, c->p, c->n, x, c->f, y, NULL, NULL, &l, NULL);
    acc[0] = acc[0] + l;
  } else {
    /* lossSum(sum, xy) = sum + loss(p, xy); loss = res*res/2, MSEFunction.scala:34,45-48 */
    double res = mse_residual(c->family, c->p, c->n, x, c->f, y);
    acc[0] = acc[0] + res * res / 2;
  }
}

double orc_apply(int family, const double* X, const double* y, int64_t m, int f, const double* p, int n, orc_fold fold) {
  eval_ctx c = { family, p, n, f, NULL, 0.0 };
  double acc;
  fold_rows(loss_row, &c, X, y, m, f, 1, fold, &acc);
  return acc / (double)m;   /* ... / data.size, MSEFunction.scala:35 */
}

/* ---------------------------------------------------------------------------------------------
 * Jacobian rows
 * ------------------------------------------------------------------------------------------- */
double orc_jacobian_row(int family, const double* p, int n, const double* x, int f, double y, double* Jrow) {
  double df[ORC_MAXN];
  if (family == ORC_MLP_MSE || family == ORC_MLP_CE) {
    /* Neuron.jacobian = gradient / residual, or the gradient itself when residual is NaN or 0 (Neuron.scala:67-73);
     * FFNeuralNetworkTrainer.backpropagateRow: AugmentedRow(jacobian, residual), Trainer.scala:127-131 */
    double residual;
    mlp_backprop(family, p, n, x, f, y, df, &residual, NULL, NULL);
    for (int j = 0; j < n; ++j) Jrow[j] = (isnan(residual) || residual == 0.0) ? df[j] : df[j] / residual;
    return residual;
  }
  model_grad(family, p, n, x, f, df);
  if (family == ORC_LOGISTIC) {
    /* Neuron.propagateError / FFNeuralNetwork.propagateErrorsOuterLayer (CrossEntropy, one output):
     * residual = output - target, error = residual (FFNeuralNetwork.scala:174-179);
     * Neuron.jacobian = (1 +: inputs) * error / residual, or * error when residual is NaN/0 (Neuron.scala:67-73) */
    double residual = orc_model(family, p, n, x, f) - y;
    double error = residual;
    for (int j = 0; j < n; ++j)
      Jrow[j] = (isnan(residual) || residual == 0.0) ? df[j] * error : df[j] * error / residual;
    return residual;
  }
  /* residual = |y - f|; d residual/dp = -sign(y - f) df/dp: the eps -> 0 limit of the forward difference
   * in SimpleMSEFunction.jacobianAndResidual (MSEFunction.scala:117-124) */
  double diff = y - orc_model(family, p, n, x, f);
  double sgn = (diff >= 0.0) ? 1.0 : -1.0;
  for (int j = 0; j < n; ++j) Jrow[j] = -sgn * df[j];
  return sqrt(diff * diff);
}

double orc_jacobian_row_fd(int family, const double* p, int n, const double* x, int f, double y, double eps, double* Jrow) {
  /* DenseVector.gradient(f, x0): fx = f(x); (f(x.updated(i, xi + eps)) - fx)/eps, DenseVector.scala:269-275 */
  double pp[ORC_MAXN];
  memcpy(pp, p, sizeof(double) * n);
  double fx = mse_residual(family, p, n, x, f, y);
  for (int j = 0; j < n; ++j) {
    pp[j] = p[j] + eps;
    Jrow[j] = (mse_residual(family, pp, n, x, f, y) - fx) / eps;
    pp[j] = p[j];
  }
  return mse_residual(family, p, n, x, f, y);   /* residual0 = residual(x0, point), MSEFunction.scala:122 */
}

void orc_jacobian_matrix(int family, const double* X, const double* y, int64_t m, int f, const double* p, int n,
                         int fd, double eps, double* rows) {
  /* data.zipWithIndex map { (xy, i) => AugmentedRow(jacobianAndResidual(p, xy), i) }, MSEFunction.scala:132-136 */
  for (int64_t i = 0; i < m; ++i) {
    double* row = rows + (size_t)i * (n + 1);
    row[n] = fd ? orc_jacobian_row_fd(family, p, n, X + (size_t)i * f, f, y[i], eps, row)
                : orc_jacobian_row(family, p, n, X + (size_t)i * f, f, y[i], row);
  }
}

/* ---------------------------------------------------------------------------------------------
 * gradient
 * ------------------------------------------------------------------------------------------- */
static void grad_row(const void* vctx, const double* x, double y, int64_t i, double* acc) {
  /* backpropagate: (previous._1 + gradient, previous._2 + residual), FFNeuralNetworkTrainer.scala:119-125.
   * acc[0..n) = gradient sum, acc[n] = residual sum */
  const eval_ctx* c = (const eval_ctx*)vctx; (void)i;
  double J[ORC_MAXN];
  if (c->family == ORC_MLP_MSE || c->family == ORC_MLP_CE) {
    double residual;
    mlp_backprop(c->family, c->p, c->n, x, c->f, y, J, &residual, NULL, NULL);
    for (int j = 0; j < c->n; ++j) acc[j] = acc[j] + J[j];
    acc[c->n] = acc[c->n] + residual;
  } else if (c->family == ORC_LOGISTIC) {
    /* Neuron.gradient = (1 +: inputs) * error, error = output - target */
    double error = orc_model(ORC_LOGISTIC, c->p, c->n, x, c->f) - y;
    model_grad(ORC_LOGISTIC, c->p, c->n, x, c->f, J);
    for (int j = 0; j < c->n; ++j) acc[j] = acc[j] + J[j] * error;
    acc[c->n] = acc[c->n] + error;
  } else {
    /* d(res^2/2)/dp = res * d res/dp */
    double r = orc_jacobian_row(c->family, c->p, c->n, x, c->f, y, J);
    for (int j = 0; j < c->n; ++j) acc[j] = acc[j] + J[j] * r;
    acc[c->n] = acc[c->n] + r;
  }
}

void orc_gradient(int family, const double* X, const double* y, int64_t m, int f, const double* p, int n,
                  orc_fold fold, double* grad) {
  eval_ctx c = { family, p, n, f, NULL, 0.0 };
  double acc[ORC_MAXN + 1];
  fold_rows(grad_row, &c, X, y, m, f, n + 1, fold, acc);
  for (int j = 0; j < n; ++j) grad[j] = acc[j] / (double)m;   /* ._1 / data.size, Trainer.scala:68 */
}

void orc_gradient_fd(int family, const double* X, const double* y, int64_t m, int f, const double* p, int n,
                     double eps, orc_fold fold, double* grad) {
  /* DenseVector.gradient(this.apply, x), MSEFunction.scala:138-140 */
  double pp[ORC_MAXN];
  memcpy(pp, p, sizeof(double) * n);
  double fx = orc_apply(family, X, y, m, f, p, n, fold);
  for (int j = 0; j < n; ++j) {
    pp[j] = p[j] + eps;
    grad[j] = (orc_apply(family, X, y, m, f, pp, n, fold) - fx) / eps;
    pp[j] = p[j];
  }
}

/* ---------------------------------------------------------------------------------------------
 * normal equations
 * ------------------------------------------------------------------------------------------- */
static void normal_row(const void* vctx, const double* x, double y, int64_t i, double* acc) {
  const eval_ctx* c = (const eval_ctx*)vctx; (void)i;
  int n = c->n;
  double J[ORC_MAXN];
  double r = orc_jacobian_row(c->family, c->p, n, x, c->f, y, J);
  for (int a = 0; a < n; ++a)
    for (int b = 0; b < n; ++b) acc[a * n + b] = acc[a * n + b] + J[a] * J[b];
  for (int a = 0; a < n; ++a) acc[n * n + a] = acc[n * n + a] + J[a] * r;
  acc[n * n + n] = acc[n * n + n] + r * r;
}

void orc_normal_eq(int family, const double* X, const double* y, int64_t m, int f, const double* p, int n,
                   orc_fold fold, double* JtJ, double* Jtr, double* rtr) {
  eval_ctx c = { family, p, n, f, NULL, 0.0 };
  int Q = n * n + n + 1;
  double* acc = (double*)malloc(sizeof(double) * Q);
  fold_rows(normal_row, &c, X, y, m, f, Q, fold, acc);
  memcpy(JtJ, acc, sizeof(double) * n * n);
  memcpy(Jtr, acc + n * n, sizeof(double) * n);
  *rtr = acc[n * n + n];
  free(acc);
}

/* The same sums, COMPENSATED: exact products (TwoProduct through fma) added with Neumaier's running error term, per
 * partition and again across partitions; hi + lo is the sum of the m exactly-formed per-row terms to ~1e-32 relative.
 * Not a restatement of anything in the reference — the yardstick that both the reference's left fold and the GPU's tree
 * are measured against at m = 1e9 ("parity at scale").  out_hi / out_lo: n*n + n + 1 doubles. */
static inline void comp_add(double* s, double* c, double v) {
  double t = *s + v;
  if (fabs(*s) >= fabs(v)) *c += (*s - t) + v; else *c += (v - t) + *s;
  *s = t;
}
static inline void comp_add_prod(double* s, double* c, double a, double b) {
  double pr = a * b, er = fma(a, b, -pr);
  comp_add(s, c, pr);
  *c += er;
}
typedef struct comp_job { int family; const double* X; const double* y; int f; const double* p; int n; int64_t b, e; double* s; double* c; } comp_job;
static void* comp_worker(void* arg) {
  comp_job* jb = (comp_job*)arg;
  int n = jb->n, Q = n * n + n + 1;
  for (int q = 0; q < Q; ++q) { jb->s[q] = 0.0; jb->c[q] = 0.0; }
  double J[ORC_MAXN];
  for (int64_t i = jb->b; i < jb->e; ++i) {
    double r = orc_jacobian_row(jb->family, jb->p, n, jb->X + (size_t)i * jb->f, jb->f, jb->y[i], J);
    for (int a = 0; a < n; ++a)
      for (int b = a; b < n; ++b) comp_add_prod(&jb->s[a * n + b], &jb->c[a * n + b], J[a], J[b]);
    for (int a = 0; a < n; ++a) comp_add_prod(&jb->s[n * n + a], &jb->c[n * n + a], J[a], r);
    comp_add_prod(&jb->s[n * n + n], &jb->c[n * n + n], r, r);
  }
  return NULL;
}
void orc_normal_eq_comp(int family, const double* X, const double* y, int64_t m, int f, const double* p, int n, int threads,
                        double* out_hi, double* out_lo) {
  if (threads < 1) threads = 1;
  if (threads > 256) threads = 256;
  int Q = n * n + n + 1;
  pthread_t th[256]; comp_job jobs[256];
  double* buf = (double*)malloc(sizeof(double) * 2 * (size_t)threads * Q);
  for (int t = 0; t < threads; ++t) {
    comp_job jb = { family, X, y, f, p, n, (int64_t)(((__int128)t * m) / threads), (int64_t)(((__int128)(t + 1) * m) / threads),
                    buf + (size_t)(2 * t) * Q, buf + (size_t)(2 * t + 1) * Q };
    jobs[t] = jb;
    if (threads > 1) pthread_create(&th[t], NULL, comp_worker, &jobs[t]); else comp_worker(&jobs[t]);
  }
  if (threads > 1) for (int t = 0; t < threads; ++t) pthread_join(th[t], NULL);
  for (int q = 0; q < Q; ++q) {
    double s = 0.0, c = 0.0;
    for (int t = 0; t < threads; ++t) { comp_add(&s, &c, jobs[t].s[q]); comp_add(&s, &c, jobs[t].c[q]); }
    out_hi[q] = s; out_lo[q] = c;
  }
  for (int a = 0; a < n; ++a)
    for (int b = 0; b < a; ++b) { out_hi[a * n + b] = out_hi[b * n + a]; out_lo[a * n + b] = out_lo[b * n + a]; }
  free(buf);
}

/* ---------------------------------------------------------------------------------------------
 * line search scores
 * ------------------------------------------------------------------------------------------- */
void orc_line(int family, const double* X, const double* y, int64_t m, int f, const double* p, const double* d, int n,
              const double* alpha, int k, orc_fold fold, double* phi, double* dphi) {
  double x[ORC_MAXN], g[ORC_MAXN];
  for (int j = 0; j < k; ++j) {
    /* pt1 = pt0.copy(x = ptInit.x + ptInit.d * a1), StrongWolfe.scala:97 */
    for (int a = 0; a < n; ++a) x[a] = p[a] + d[a] * alpha[j];
    phi[j] = orc_apply(family, X, y, m, f, x, n, fold);          /* fx, Point.scala:54 */
    orc_gradient(family, X, y, m, f, x, n, fold, g);
    dphi[j] = inner(g, d, n);                                      /* gradient(x) dot d, ObjectiveFunctionWithGradient.scala:37-39 */
  }
}

/* ---------------------------------------------------------------------------------------------
 * Hessian-vector products
 * ------------------------------------------------------------------------------------------- */
void orc_hess_vec_fd(int family, const double* X, const double* y, int64_t m, int f, const double* p, const double* d,
                     int n, double eps, orc_fold fold, double* Hd) {
  /* gradx = gradient(x); gradxd = gradient(x + d*eps); (gradxd - gradx)/eps, ObjectiveFunctionWithGradient.scala:41-47 */
  double g0[ORC_MAXN], g1[ORC_MAXN], xd[ORC_MAXN];
  orc_gradient(family, X, y, m, f, p, n, fold, g0);
  for (int a = 0; a < n; ++a) xd[a] = p[a] + d[a] * eps;
  orc_gradient(family, X, y, m, f, xd, n, fold, g1);
  for (int a = 0; a < n; ++a) Hd[a] = (g1[a] - g0[a]) / eps;
}

static void hess_row(const void* vctx, const double* x, double y, int64_t i, double* acc) {
  const eval_ctx* c = (const eval_ctx*)vctx; (void)i;
  int n = c->n;
  double df[ORC_MAXN], h[ORC_MAXN];
  model_grad(c->family, c->p, n, x, c->f, df);
  double jd = inner(df, c->d, n);
  if (c->family == ORC_LOGISTIC) {
    /* d/dp [(o - y) (1 +: x)] . d = o (1 - o) ((1 +: x) . d) (1 +: x); LogisticFunction.derivative, :26 */
    double o = orc_model(ORC_LOGISTIC, c->p, n, x, c->f);
    double w = o * (1.0 - o) * jd;
    for (int j = 0; j < n; ++j) acc[j] = acc[j] + w * df[j];
  } else {
    /* loss = (y - f)^2/2  =>  H d = (df . d) df - (y - f) (d2f . d) */
    double rho = y - orc_model(c->family, c->p, n, x, c->f);
    model_hess_dir(c->family, c->p, n, x, c->f, c->d, h);
    for (int j = 0; j < n; ++j) acc[j] = acc[j] + (jd * df[j] - rho * h[j]);
  }
}

void orc_hess_vec(int family, const double* X, const double* y, int64_t m, int f, const double* p, const double* d,
                  int n, orc_fold fold, double* Hd) {
  eval_ctx c = { family, p, n, f, d, 0.0 };
  double acc[ORC_MAXN];
  fold_rows(hess_row, &c, X, y, m, f, n, fold, acc);
  for (int j = 0; j < n; ++j) Hd[j] = acc[j] / (double)m;
}

/* =============================================================================================
 * QR.apply restated pass for pass (linalg/QR.scala:131-312).  The data set is the materialised
 * augmented matrix; row index i of row k is k (SeqDataSet.zipWithIndex, SeqDataSetConverter.scala:57-59).
 * =========================================================================================== */
/* row-range parallel-for used by the QR passes: `parts` contiguous partitions on up to `parts` threads,
 * partial results combined in partition order (the par[N] model of oracle.h) */
typedef void (*range_fn)(void* ctx, int part, int64_t b, int64_t e);
typedef struct range_job { range_fn fn; void* ctx; int part; int64_t b, e; } range_job;
static void* range_worker(void* arg) { range_job* j = (range_job*)arg; j->fn(j->ctx, j->part, j->b, j->e); return NULL; }
static void parallel_ranges(range_fn fn, void* ctx, int64_t lo, int64_t hi, int parts) {
  if (parts <= 1 || hi - lo < 4096) { fn(ctx, 0, lo, hi); return; }
  pthread_t th[256]; range_job jobs[256];
  if (parts > 256) parts = 256;
  int64_t span = hi - lo;
  for (int t = 0; t < parts; ++t) {
    jobs[t].fn = fn; jobs[t].ctx = ctx; jobs[t].part = t;
    jobs[t].b = lo + (int64_t)(((__int128)t * span) / parts);
    jobs[t].e = lo + (int64_t)(((__int128)(t + 1) * span) / parts);
    pthread_create(&th[t], NULL, range_worker, &jobs[t]);
  }
  for (int t = 0; t < parts; ++t) pthread_join(th[t], NULL);
}

typedef struct qr_pass {
  double* ab; int n, w, j, kmax; double* partial;   /* partial[parts*w] */
  const double* dots; const double* rowj; double ajnorm, ajj;
} qr_pass;

static void qr_norm_range(void* vc, int part, int64_t b, int64_t e) {
  qr_pass* c = (qr_pass*)vc; double* acc = c->partial + (size_t)part * c->w;
  for (int k = 0; k < c->w; ++k) acc[k] = 0.0;
  for (int64_t i = b; i < e; ++i) {
    const double* row = c->ab + (size_t)i * c->w;
    for (int k = 0; k < c->w; ++k) acc[k] = acc[k] + row[k] * row[k];   /* columnsNorm, QR.scala:259-263 */
  }
}
static void qr_swap_range(void* vc, int part, int64_t b, int64_t e) {
  qr_pass* c = (qr_pass*)vc; (void)part;
  for (int64_t i = b; i < e; ++i) {
    double* row = c->ab + (size_t)i * c->w;
    double t = row[c->j]; row[c->j] = row[c->kmax]; row[c->kmax] = t;    /* swapColumns, :283-285 */
  }
}
static void qr_dot_range(void* vc, int part, int64_t b, int64_t e) {
  /* columnsDotProduct(j, j) over rows with i > j, QR.scala:297-312 */
  qr_pass* c = (qr_pass*)vc; double* acc = c->partial + (size_t)part * c->w; const int n = c->n, j = c->j;
  for (int k = 0; k < c->w; ++k) acc[k] = 0.0;
  for (int64_t i = b; i < e; ++i) {
    const double* row = c->ab + (size_t)i * c->w;
    double xj = row[j];
    for (int k = 0; k < n; ++k) acc[k] = acc[k] + ((k >= j) ? row[k] * xj : row[k]);
    acc[n] = acc[n] + xj * row[n];
  }
}
static void qr_update_range(void* vc, int part, int64_t b, int64_t e) {
  /* updateColumn(j) on rows with i >= j, QR.scala:201-224 */
  qr_pass* c = (qr_pass*)vc; (void)part; const int n = c->n, j = c->j;
  for (int64_t i = b; i < e; ++i) {
    double* row = c->ab + (size_t)i * c->w;
    double aij = (i == j) ? c->ajj : row[j] / c->ajnorm;
    double alphab = c->dots[n] / c->ajnorm / c->ajj + c->rowj[n];
    double newb = row[n] - alphab * aij;
    for (int k = 0; k < n; ++k) {
      if (k == j) row[k] = aij;
      else if (k > j) { double alpha = c->dots[k] / c->ajnorm / c->ajj + c->rowj[k]; row[k] = row[k] - alpha * aij; }
    }
    row[n] = newb;
  }
}

static int qr_impl(const double* ab_in, int64_t m, int n, int pivoting, int parts, double* r, double* rdiag,
                   double* qtb, int* ipvt, double* acnorms, double* bnorm) {
  const int w = n + 1;
  if (parts < 1) parts = 1;
  if (parts > 256) parts = 256;
  double* ab = (double*)malloc(sizeof(double) * (size_t)(m > 0 ? m : 1) * w);
  memcpy(ab, ab_in, sizeof(double) * (size_t)m * w);
  double* dots = (double*)malloc(sizeof(double) * w);    /* dotProducts.a[0..n), dotProducts.b */
  double* rowj = (double*)malloc(sizeof(double) * w);
  double* upd = (double*)malloc(sizeof(double) * w);
  double* partial = (double*)calloc((size_t)parts * w, sizeof(double));
  qr_pass pass = { ab, n, w, 0, 0, partial, dots, rowj, 0.0, 0.0 };
  const int eff_parts = (m < 4096) ? 1 : parts;

  /* abNorms = doSqrt(ab.aggregate(zeros(n+1))(columnsNorm, _ + _)), QR.scala:136-137,259-267 */
  parallel_ranges(qr_norm_range, &pass, 0, m, parts);
  for (int k = 0; k < w; ++k) {
    double s = 0.0;
    for (int t = 0; t < eff_parts; ++t) s = s + partial[(size_t)t * w + k];
    if (k < n) acnorms[k] = sqrt(s); else *bnorm = sqrt(s);
  }
  for (int k = 0; k < n; ++k) { ipvt[k] = k; rdiag[k] = acnorms[k]; }   /* QRObject(abNorms.a, ab, ipvt), :140-144 */

  for (int j = 0; j < n; ++j) {                                          /* houseHolder, :169-250 */
    pass.j = j;
    if (pivoting) {
      /* indexColumnLargestNorm: first strict maximum over k >= j, :269-273 */
      int kmax = j; double best = rdiag[j];
      for (int k = j; k < n; ++k) if (rdiag[k] > best) { best = rdiag[k]; kmax = k; }
      if (kmax != j) {
        /* swapColumns swaps the matrix columns and ipvt — NOT rDiag (reference behaviour), :278-285 */
        pass.kmax = kmax;
        parallel_ranges(qr_swap_range, &pass, 0, m, parts);
        int t = ipvt[j]; ipvt[j] = ipvt[kmax]; ipvt[kmax] = t;
      }
    }
    /* (dotProducts, rowj) = aggregate((zeros(n), zeros(n)))(columnsDotProduct(j, j), sumResults), :190-192.
     * Rows with i < j are skipped, row i == j is kept as rowj, rows i > j feed the sums. */
    for (int k = 0; k < w; ++k) { dots[k] = 0.0; rowj[k] = 0.0; }
    if (j < m) memcpy(rowj, ab + (size_t)j * w, sizeof(double) * w);
    if (m > j + 1) {
      parallel_ranges(qr_dot_range, &pass, j + 1, m, parts);
      int used = (m - (j + 1) < 4096) ? 1 : parts;
      for (int k = 0; k < w; ++k) {
        double s = 0.0;
        for (int t = 0; t < used; ++t) s = s + partial[(size_t)t * w + k];
        dots[k] = s;
      }
    }
    double rjj = rowj[j];
    double sgn = (rjj > 0.0) ? 1.0 : ((rjj < 0.0) ? -1.0 : 0.0);        /* Math.signum */
    double ajnorm = sqrt(dots[j] + rjj * rjj) * sgn;                     /* :195-196 */
    double ajj = (ajnorm != 0.0) ? rjj / ajnorm + 1.0 : 0.0;             /* :198 */
    if (ajnorm == 0.0) { rdiag[j] = 0.0; continue; }                     /* :237-238 */

    /* ajUpdatedRow = rowj.a.mapWithIndex(updateRow(j, j, ajj / ajNorm + 1.0)), :240-241 */
    {
      double aij = ajj / ajnorm + 1.0;
      for (int k = 0; k < n; ++k) {
        if (k == j) upd[k] = aij;
        else if (k > j) { double alpha = dots[k] / ajnorm / ajj + rowj[k]; upd[k] = rowj[k] - alpha * aij; }
        else upd[k] = rowj[k];
      }
    }
    /* rDiag1 = rDiag.mapWithIndex(rDiagUpdate(ajUpdatedRow)).updated(j, -ajNorm), :227-234,243-244 */
    for (int k = j + 1; k < n; ++k) {
      double akj = upd[k], rr = rdiag[k];
      rdiag[k] = rr * sqrt(fmax(0.0, 1 - akj * akj / rr / rr));
    }
    rdiag[j] = -ajnorm;
    /* ab0.map(updateColumn(j)), :214-224,246 */
    pass.ajnorm = ajnorm; pass.ajj = ajj;
    if (m > j) parallel_ranges(qr_update_range, &pass, j, m, parts);
  }
  /* reducedMat = ab.filter(_.i < n).collect().sortBy(_.i); r = rows' a; qtb = rows' b, :147-149 */
  int ok = (m >= n && n > 0) ? 0 : -1;    /* require(n == r.length), :46 */
  if (ok == 0) {
    for (int i = 0; i < n; ++i) {
      memcpy(r + (size_t)i * n, ab + (size_t)i * w, sizeof(double) * n);
      qtb[i] = ab[(size_t)i * w + n];
    }
  }
  free(ab); free(dots); free(rowj); free(upd); free(partial);
  return ok;
}

int orc_qr(const double* ab_in, int64_t m, int n, int pivoting, double* r, double* rdiag, double* qtb, int* ipvt,
           double* acnorms, double* bnorm) {
  return qr_impl(ab_in, m, n, pivoting, 1, r, rdiag, qtb, ipvt, acnorms, bnorm);
}

void orc_qr_solution(const double* r, const double* rdiag, const double* qtb, const int* ipvt, int n, double* x) {
  /* QR.solution, QR.scala:97-114 */
  int first_zero = -1;
  for (int j = 0; j < n; ++j) if (rdiag[j] == 0.0) { first_zero = j; break; }
  int nsing = (first_zero == -1) ? n - 1 : first_zero;
  double xs[ORC_MAXN];
  for (int j = 0; j < n; ++j) xs[j] = (j <= nsing) ? qtb[j] : 0.0;
  for (int j = nsing; j >= 0; --j) {
    if (j >= n) continue;
    double sum = 0.0;
    for (int k = j + 1; k < n; ++k) sum += xs[k] * r[(size_t)j * n + k];
    xs[j] = (xs[j] - sum) / rdiag[j];
  }
  for (int j = 0; j < n; ++j) x[ipvt[j]] = xs[j];   /* DenseVector.permute, DenseVector.scala:315-321 */
}

typedef struct jac_job { int family; const double* X; const double* y; int f; const double* p; int n; double* rows; } jac_job;
static void jac_fd_range(void* vc, int part, int64_t b, int64_t e) {
  jac_job* c = (jac_job*)vc; (void)part;
  for (int64_t i = b; i < e; ++i) {
    double* row = c->rows + (size_t)i * (c->n + 1);
    row[c->n] = orc_jacobian_row_fd(c->family, c->p, c->n, c->X + (size_t)i * c->f, c->f, c->y[i], 1.0e-8, row);
  }
}

double orc_lm_reference_passes(int family, const double* X, const double* y, int64_t m, int f, const double* p, int n,
                               int threads) {
  /* what one LevenbergMarquardt.iterate costs on a data set before the O(n^3) host part:
   * QR(f.jacobianAndResidualsMatrix(x0), n, usePivoting=false) (:124) + sqrt(f(x)) (:164).
   * threads > 1 = the par[N] model (Spark local[N]) of the same passes. */
  double* rows = (double*)malloc(sizeof(double) * (size_t)m * (n + 1));
  double* r = (double*)malloc(sizeof(double) * n * n);
  double rdiag[ORC_MAXN], qtb[ORC_MAXN], acn[ORC_MAXN], bnorm = 0.0;
  int ipvt[ORC_MAXN];
  jac_job jj = { family, X, y, f, p, n, rows };
  parallel_ranges(jac_fd_range, &jj, 0, m, threads);
  qr_impl(rows, m, n, 0, threads, r, rdiag, qtb, ipvt, acn, &bnorm);
  orc_fold fold = { threads < 1 ? 1 : threads, threads < 1 ? 1 : threads };
  double loss = orc_apply(family, X, y, m, f, p, n, fold);
  free(rows); free(r);
  return bnorm + 0.0 * loss;
}

/* =============================================================================================
 * Negative log-likelihood (std-apps/.../stats/package.scala:97-103)
 * =========================================================================================== */
double orc_pdf(int pdf, const double* p, int n, const double* consts, double x) {
  (void)n;
  if (pdf != ORC_PDF_GAUSS_EXPBKG) return NAN;
  /* ExSignalPlusBackgroundFit.pdf (:78-88): pdfSignal(signalPars, x) * fSig + pdfBkg(bkgPars, x) * (1.0 - fSig) */
  const double fSig = p[0], mu = p[1], sigma = p[2], lambda = p[3], xMin = consts[0];
  const double z = (x - mu) / sigma;                                                   /* :101 */
  const double sig = exp(-z * z / 2.0) / sqrt(2.0 * 3.14159265358979323846) / sigma;   /* :102 */
  const double bkg = lambda * exp(-lambda * (x - xMin));                               /* :114 */
  return sig * fSig + bkg * (1.0 - fSig);
}

typedef struct nll_ctx { int pdf; const double* p; int n; const double* consts; } nll_ctx;
static void nll_row(const void* vctx, const double* x, double y, int64_t i, double* acc) {
  const nll_ctx* c = (const nll_ctx*)vctx; (void)y; (void)i;
  acc[0] = acc[0] - 2.0 * log(orc_pdf(c->pdf, c->p, c->n, c->consts, x[0]));   /* sum - 2.0 * Math.log(pdf(p, x)(0)) */
}

double orc_nll(int pdf, const double* x, int64_t m, const double* p, int n, const double* consts, orc_fold fold) {
  nll_ctx c = { pdf, p, n, consts };
  double acc;
  fold_rows(nll_row, &c, x, x, m, 1, 1, fold, &acc);   /* the y stream is unused */
  return acc;
}

void orc_nll_example_data(int64_t seed, int nb_events, double f_signal, double mu, double sigma, double lambda,
                          double x_min, double* x) {
  orc_jrandom r;
  orc_jrandom_init(&r, seed);
  for (int i = 0; i < nb_events; ++i) {                                       /* :48-55 */
    if (orc_jrandom_next_double(&r) < f_signal) x[i] = orc_jrandom_next_gaussian(&r) * sigma + mu;
    else x[i] = x_min - log(orc_jrandom_next_double(&r)) / lambda;
  }
}

/* =============================================================================================
 * Synthetic observations — counter-based so any shard can be generated anywhere
 * =========================================================================================== */
static uint64_t splitmix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
static double u01(uint64_t seed, uint64_t ctr) { return (double)(splitmix64(seed ^ ctr) >> 11) * 0x1.0p-53; }
static double n01(uint64_t seed, uint64_t ctr) {
  double u1 = u01(seed, ctr), u2 = u01(seed, ctr ^ 0x8000000000000000ULL);
  return sqrt(-2.0 * log(1.0 - u1)) * cos(6.283185307179586476925 * u2);
}

void orc_generate(int family, int64_t m, int f, const double* p_true, int n, uint64_t seed, double noise_sigma,
                  int64_t row_offset, int64_t total_rows, double* X, double* y) {
  if (total_rows <= 0) total_rows = m;
  for (int64_t i = 0; i < m; ++i) {
    uint64_t gi = (uint64_t)(row_offset + i);
    uint64_t base = gi * (uint64_t)(f + 1);
    double* x = X + (size_t)i * f;
    switch (family) {
      case ORC_LINEAR: case ORC_LINEAR_NOINT:
        for (int j = 0; j < f; ++j) x[j] = u01(seed, base + j);     /* U[0,1): LevenbergMarquardtSpec.scala:83-86 */
        break;
      case ORC_LOGISTIC:
        for (int j = 0; j < f; ++j) x[j] = n01(seed, base + j);     /* N(0,1): FFNeuralNetworkTrainerSpec.scala:202-207 */
        break;
      default:
        x[0] = ((double)gi + 0.5) / (double)total_rows;             /* t grid on (0,1) */
        break;
    }
    double mu = orc_model(family, p_true, n, x, f);
    if (family == ORC_LOGISTIC) y[i] = (u01(seed, base + f) < mu) ? 1.0 : 0.0;
    else y[i] = mu + noise_sigma * n01(seed, base + f);
  }
}
