/*
 * scalaopt_cuda.h — C ABI of libscalaopt_cuda.so
 *
 * B200 (sm_100a) evaluation of scalaopt's data-set objective: loss, gradient J^T r,
 * Gauss-Newton Gram J^T J, Hessian-vector products and multi-alpha line-search scores over
 * fp64 observations resident in HBM.  This is the drop-in boundary behind the reference's
 * `DataSet` / `MSEFunction` / `DifferentiableObjectiveFunction` traits; the reference has no
 * FFI of its own (it is pure Scala), so every entry point below cites the reference
 * interface it replaces.  Paths are relative to the reference root, with
 *   core/... = core/src/main/scala/com/github/bruneli/scalaopt/core/...
 *
 * Conventions
 *   - blocking calls, `int` return: 0 = ok, <0 = one of SO_E*; message via so_last_error()
 *   - plain pointers and sizes only; caller owns every host buffer, the library owns device
 *     memory behind opaque handles
 *   - a context may be used by one host thread at a time (internal mutex)
 *   - observations: row-major X[m x f] and y[m] (scalar output), fp64, row-sharded in
 *     contiguous blocks across the context's GPUs
 *   - there is NO CPU fallback: every so_eval_* runs the sm_100a kernels or fails
 */
#ifndef SCALAOPT_CUDA_H
#define SCALAOPT_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SO_ABI_VERSION 2

/* error codes */
#define SO_OK            0
#define SO_EINVAL       -1   /* bad argument (IllegalArgumentException on the Scala side) */
#define SO_ECUDA        -2   /* CUDA runtime / driver error */
#define SO_ENOMEM       -3   /* device or pinned-host allocation failed */
#define SO_ENCCL        -4   /* NCCL error */
#define SO_EUNSUPPORTED -5   /* family / shape outside the catalogue */
#define SO_ESTATE       -6   /* e.g. evaluating a data set that has not been filled */

/*
 * Model-family catalogue.  Arbitrary Scala closures (core/.../function/RegressionFunction.scala:24-37)
 * cannot run on the device, so the regression function is one of these.
 *   p = parameter vector (n doubles), x = one observation's inputs (f doubles), y scalar.
 */
typedef enum so_family {
  /* y ~ p0 + sum_j p[j+1] x[j]; n = f+1.  core/src/test/.../leastsquares/LevenbergMarquardtSpec.scala:38-40 */
  SO_FAMILY_LINEAR = 0,
  /* y ~ sum_j p[j] x[j]; n = f (BASELINE config 4) */
  SO_FAMILY_LINEAR_NOINT = 1,
  /* o = 1/(1+exp(-(p0 + sum_j p[j+1] x[j]))), cross-entropy loss, y in [0,1]; n = f+1.
   * std-apps/.../learning/nnet/{Neuron.scala:40-49, activation/LogisticFunction.scala:24-26,
   * FFNeuralNetwork.scala:80-88,174-179, FFNeuralNetworkTrainer.scala:59-69} */
  SO_FAMILY_LOGISTIC = 2,
  /* y ~ p0 * exp(p1 * t); f = 1, n = 2.  LevenbergMarquardtSpec.scala:62-63 */
  SO_FAMILY_EXPDECAY = 3,
  /* y ~ A exp(-(t-mu)^2 / (2 sigma^2)); p = (A, mu, sigma); f = 1, n = 3 */
  SO_FAMILY_GAUSS = 4,
  /* y ~ A exp(-(t-mu)^2/(2 sigma^2)) + B exp(-lambda t); p = (A, mu, sigma, B, lambda); f = 1, n = 5.
   * functional form of std-apps/.../stats/example/ExSignalPlusBackgroundFit.scala:78-114 */
  SO_FAMILY_GAUSS_EXPBKG = 5,
  /* y ~ sum_k p[k] t^k; f = 1, 1 <= n <= SO_MAX_POLY_N */
  SO_FAMILY_POLY = 6,
  /* FFNeuralNetwork(Vector(f, h, 1)) with logistic inner neurons trained by FFNeuralNetworkTrainer (SURVEY.md 8f.3;
   * std-apps/.../nnet/FFNeuralNetwork.scala:114-189, Neuron.scala:40-73, FFNeuralNetworkTrainer.scala:59-131):
   * n = h (f + 1) + h + 1 weights in the reference's order (hidden neurons, then the output neuron; bias first), so h
   * follows from n and f.  MLP_MSE: LossType.MeanSquaredError with a LinearFunction output (loss = residual^2 — the
   * reference does not halve it here); MLP_CE: LossType.CrossEntropy with a LogisticFunction output, y in [0, 1].
   * Catalogue: f <= 4, h <= 4; so_eval_normal_eq additionally needs n <= 9 (Vector(2, 2, 1), the shape of the
   * reference's own spec).  so_eval_line / so_eval_hess_vec are composed from loss+gradient passes the way the
   * reference composes them (dirHessian = forward difference of gradients, eps = 1e-8, Trainer.scala:76-80). */
  SO_FAMILY_MLP_MSE = 7,
  SO_FAMILY_MLP_CE = 8,
  /* Round 2: the network families also take K = ny > 1 outputs per observation (so_dataset_create_multi), which is what the
   * reference's own spec trains — FFNeuralNetwork(Vector(Iris.nInputs, 5, Iris.nClasses)), FFNeuralNetworkTrainerSpec.scala:353-376:
   *   n = h (f + 1) + K (h + 1) weights, reference order (hidden neurons, then output neurons; bias first)
   *   MLP_MSE, K > 1: SO_EUNSUPPORTED — the reference has `???` there (FFNeuralNetwork.scala:157)
   *   MLP_CE,  K > 1: SoftMaxFunction outputs o_k = exp(z_k - max) / sum (:143-158), loss = -sum_k y_k log(o_k / sum o)
   *                   (:89-103), residual_k = (sum y) o_k - y_k (:168-173)
   *   LM convention (Neuron.jacobian, Neuron.scala:67-73): network residual = ||residual||_2, Jacobian row = gradient /
   *   residual of the neuron (hidden neurons: of the FIRST output neuron), as the reference has it.
   * SOFTMAX: no hidden layer — FFNeuralNetwork(Vector(f, K)) with CrossEntropy + SoftMaxFunction (spec :222-260), n = K (f + 1), K >= 2.
   * Shapes: f <= 16, h <= 16, K <= SO_MAX_OUTPUTS.  Weight decay (FFNeuralNetworkTrainer.scala:59-69,147-153) is data-independent
   * and is applied by the caller's objective (GpuMSEFunction::decay). */
  SO_FAMILY_SOFTMAX = 9,
  SO_FAMILY_COUNT = 10
} so_family;

#define SO_MAX_POLY_N   8      /* polynomial family: at most 8 coefficients */
#define SO_MAX_FEATURES 512    /* linear-predictor families: f <= 512 */
#define SO_MAX_ALPHAS   32     /* so_eval_line: k <= 32 step sizes per pass */
#define SO_MAX_OUTPUTS  8      /* outputs per observation (network families) */

typedef struct so_ctx so_ctx;
typedef struct so_dataset so_dataset;

/* ---------------------------------------------------------------------------------------------
 * Context
 * ------------------------------------------------------------------------------------------- */

/* Single-process mode: this process drives `ngpus` devices (device_ids may be NULL = 0..ngpus-1).
 * Replaces the SparkContext behind spark-apps/.../SparkDataSetConverter.scala:32.  With ngpus > 1 an
 * NCCL communicator is created with ncclCommInitAll. */
int so_init(int ngpus, const int* device_ids, so_ctx** out);

/* One-process-per-GPU mode (torchrun): rank `rank` of `world` drives device `device_id`.
 * `nccl_uid` is the SO_NCCL_UID_BYTES-byte id made by so_nccl_unique_id() on rank 0 and broadcast by the
 * caller (any channel).  world == 1 needs no id. */
#define SO_NCCL_UID_BYTES 128
int so_nccl_unique_id(void* out_uid, size_t bytes);
int so_init_rank(int device_id, int rank, int world, const void* nccl_uid, size_t uid_bytes, so_ctx** out);

/* Data sets of the context must be freed (so_dataset_free) BEFORE so_shutdown; handles are invalid afterwards. */
void so_shutdown(so_ctx* ctx);

/* Message of the last failing call on `ctx` (ctx may be NULL: last so_init* failure of this thread). */
const char* so_last_error(so_ctx* ctx);

int so_abi_version(void);

/* Pinned (page-locked) host memory so that so_dataset_append runs at full PCIe speed, e.g. as the backing
 * store of a JNI direct ByteBuffer.  Plain malloc'ed memory works too, only slower. */
int  so_host_alloc(size_t bytes, void** out);
void so_host_free(void* p);

/* ---------------------------------------------------------------------------------------------
 * Data set  —  core/.../DataSet.scala:27-145 (`trait DataSet[A]`, A = DataPoint, core/.../variable/Point.scala:29)
 * ------------------------------------------------------------------------------------------- */

/* Reserve HBM for `m` observations of `f` inputs + 1 output.  In single-process mode m is the whole data
 * set (sharded evenly, contiguous row blocks); in rank mode m is THIS rank's shard and the data-set size
 * used for the 1/m normalisation is the sum over ranks. */
int so_dataset_create(so_ctx* ctx, int64_t m, int f, so_dataset** out);
/* ... with `ny` outputs per observation: y is row-major [m x ny] in append / read / head / load_raw (DataPoint.y is a vector in
 * the reference, variable/Point.scala:29; every family but the network ones needs ny = 1). */
int so_dataset_create_multi(so_ctx* ctx, int64_t m, int f, int ny, so_dataset** out);

/* The row range [*begin, *begin + *rows) of shard `g` when `m` observations of `f` inputs are split over `G` GPUs: contiguous
 * blocks of ceil(m/G) rows, an even count for f == 1 (every shard then starts 16-byte aligned).  This is the split
 * so_dataset_create makes in single-process mode and the one a rank-mode caller should give rank g; the analogue of the
 * partitioner behind spark-apps/.../SparkDataSetConverter.scala:42-45.  Pure host arithmetic: callable without a GPU. */
int so_shard_plan(int64_t m, int f, int G, int g, int64_t* begin, int64_t* rows);

/* Append `rows` observations from host memory.  Pinned sources (so_host_alloc, cudaHostRegister) are copied directly;
 * pageable sources are staged through two pinned buffers so that the host-side memcpy overlaps the DMA.  `DataSet.++` (DataSet.scala:36) / Seq -> data-set conversion (SeqDataSetConverter.scala:28). */
int so_dataset_append(so_dataset* ds, const double* X_rowmajor, const double* y, int64_t rows);

/* Append `rows` observations from raw little-endian fp64 files: `path_X` holds row-major X (f doubles per row), `path_y`
 * one double per row; reading starts `file_row_offset` rows into both files.  Chunks are read into two pinned staging
 * buffers and copied to HBM asynchronously, so disk reads, H2D copies and (multi-GPU) the shards' copies overlap.
 * The reference has no on-disk format (its data sets are Scala collections or RDDs); this is the converter a
 * `SparkDataSet -> GpuDataSet` bridge would write partitions through. */
int so_dataset_load_raw(so_dataset* ds, const char* path_X, const char* path_y, int64_t rows, int64_t file_row_offset);

/* Forget the observations but keep the HBM reservation (re-ingest into the same data set). */
int so_dataset_clear(so_dataset* ds);

/* Fill the whole data set on the device with the synthetic observations of SURVEY.md §8d:
 * counter-based RNG u = splitmix64(seed ^ (global_row*(f+1)+col)), inputs per `family`,
 * y = f(p_true, x) + noise_sigma * N(0,1) (Bernoulli(o) labels for the logistic family).
 * `row_offset` is the global index of this data set's first row (rank mode: rank r passes r*m so the ranks
 * hold consecutive shards of one global data set of world*m rows). */
int so_dataset_generate(so_dataset* ds, int family, const double* p_true, int n,
                        uint64_t seed, double noise_sigma, int64_t row_offset);

/* Copy observations [row_begin, row_begin+rows) back to the host (`DataSet.collect`, DataSet.scala:51;
 * bounded: the caller chooses the window). Row indices are local to this process. */
int so_dataset_read(so_dataset* ds, int64_t row_begin, int64_t rows, double* X_rowmajor, double* y);

/* `DataSet.size` (DataSet.scala:126): number of observations filled so far, summed over ranks. */
int so_dataset_size(so_dataset* ds, int64_t* m);

/* `DataSet.head` (DataSet.scala:82): first observation. */
int so_dataset_head(so_dataset* ds, double* x, double* y);

/* Non-owning view of local rows [begin, end) — the data-set analogue of `filter` on the row index
 * (DataSet.scala:60, as used by QR.scala:147).  A view aliases its parent's device memory: rows appended to the parent later
 * do not join it, and once the parent has been freed or cleared every call on the view fails with SO_ESTATE (the library
 * keeps a generation per owning data set; the view itself must still be released with so_dataset_free). */
int so_dataset_slice(so_dataset* ds, int64_t begin, int64_t end, so_dataset** view);

void so_dataset_free(so_dataset* ds);

/* ---------------------------------------------------------------------------------------------
 * Evaluation  —  one pass over the resident observations per call; when the context spans several GPUs the small
 * result is summed across them by the evaluation kernel's folding block over NVLink peer memory, in rank order
 * (NCCL all-reduce only above 16384 doubles).  Replaces `DataSet.aggregate`, DataSet.scala:49
 * ------------------------------------------------------------------------------------------- */

/* K1. loss and gradient in one fused pass.
 *   MSE families:  loss = (1/m) sum_i r_i^2/2,  r_i = |y_i - f(p, x_i)|      (MSEFunction.scala:33-36,45-48,106-108)
 *                  grad = (1/m) sum_i r_i dr_i/dp  (analytic; replaces the n+1-pass finite difference
 *                         of MSEFunction.scala:138-140 / linalg/DenseVector.scala:269-275)
 *   logistic:      loss = (1/m) sum_i CE(o_i, y_i)                           (FFNeuralNetwork.scala:80-88, Trainer:59-62)
 *                  grad = (1/m) sum_i (o_i - y_i) [1, x_i]                   (Neuron.scala:65, Trainer:64-69)
 * grad may be NULL (loss-only pass, e.g. LevenbergMarquardt.scala:164). */
int so_eval_loss_grad(so_dataset* ds, int family, const double* p, int n, double* loss, double* grad);

/* K2. normal equations of the Jacobian/residual system that MSEFunction.jacobianAndResidualsMatrix
 * (MSEFunction.scala:74,132-136) would materialise and QR.apply (linalg/QR.scala:131-152) would factor:
 *   JtJ[a*n+b] = sum_i J_ia J_ib,  Jtr[a] = sum_i J_ia r_i,  rtr = sum_i r_i^2   (NOT divided by m)
 * with J_i = dr_i/dp = -sign(y_i - f) df/dp  (the analytic limit of MSEFunction.scala:117-124).
 * Logistic family: the reference trainer's LM convention J_i = [1, x_i], r_i = o_i - y_i
 * (Neuron.scala:67-73, FFNeuralNetwork.scala:174-179).  loss at p = rtr/(2m) for the MSE families. */
int so_eval_normal_eq(so_dataset* ds, int family, const double* p, int n,
                      double* JtJ /* n*n row-major, full symmetric */, double* Jtr /* n */, double* rtr);

/* K2 without the normal equations: the R factor of the m x (n+1) matrix [J | r] that
 * QR.apply (linalg/QR.scala:131-250) factors with 2n+1 passes over the data set — here one pass, Householder TSQR on
 * the rows themselves (per-thread triangles merged up a fixed tree, then across GPUs), so the condition number of J
 * is NOT squared.  R is (n+1) x (n+1) row-major upper triangular with
 *   R^T R = [ JtJ  Jtr ; Jtr^T  rtr ]   (same conventions and scaling as so_eval_normal_eq),
 * i.e. rows i < n are (R_J[i,:] | (Q^T r)_i) and R[n][n] = +-||r - J J^+ r||: exactly the (n+1)-row system
 * MSEFunction.jacobianAndResidualsMatrix hands to QR.apply in the drop-in (row signs are arbitrary, as with any QR).
 * Single-input families, and the linear / logistic families with n <= 63: per-thread triangles up to n = 8 (HBM- to
 * FP64-bound), one triangle per warp in shared memory beyond (issue-bound: the robust route, not the fast one).
 * SO_EUNSUPPORTED otherwise. */
int so_eval_qr(so_dataset* ds, int family, const double* p, int n, double* R /* (n+1)*(n+1) */);

/* K3. k step sizes scored in one pass:  phi[j] = loss(p + alpha[j] d),  dphi[j] = grad(p + alpha[j] d) . d
 * (LineSearchPoint.fx / dfx, variable/Point.scala:54-57, as requested by linesearch/StrongWolfe.scala:96-118,214-217). */
int so_eval_line(so_dataset* ds, int family, const double* p, const double* d, int n,
                 const double* alpha, int k, double* phi, double* dphi);

/* K4. Hessian-vector product Hd = (d/d eps) grad(p + eps d) at eps = 0, exact (analytic second
 * derivatives), replacing the two-gradient finite difference of
 * function/ObjectiveFunctionWithGradient.scala:41-47 consumed at gradient/NewtonCG.scala:91-96,
 * gradient/SteihaugCG.scala:101-111 and variable/Point.scala:75. */
int so_eval_hess_vec(so_dataset* ds, int family, const double* p, const double* d, int n, double* Hd);

/* Negative log-likelihood: *nll = sum_i -2 log pdf(p, x_i) over the data set's input column (f = 1; the
 * y column is not read), NOT divided by m — `negLogLikelihood(pdf, data)(p)`, std-apps/.../stats/package.scala:97-103, the
 * objective `maxLikelihoodFit` (:74-87) hands to Nelder-Mead / BFGS.  Probability-density catalogue:
 *   SO_PDF_GAUSS_EXPBKG  p = (fSig, mu, sigma, lambda), consts = (xMin):
 *                        N(x; mu, sigma) fSig + lambda exp(-lambda (x - xMin)) (1 - fSig)
 *                        (std-apps/.../stats/example/ExSignalPlusBackgroundFit.scala:78-114) */
typedef enum so_pdf { SO_PDF_GAUSS_EXPBKG = 0, SO_PDF_COUNT = 1 } so_pdf;
int so_eval_nll(so_dataset* ds, int pdf, const double* p, int n, const double* consts, int nconsts, double* nll);

/* ---------------------------------------------------------------------------------------------
 * Per-call counters (the reference has no metrics; new)
 * ------------------------------------------------------------------------------------------- */
typedef struct so_call_stats {
  int64_t rows;            /* observations streamed by the last so_eval_* (this process) */
  int64_t bytes;           /* algorithmic HBM bytes of that pass: rows * 8 * (f+1) */
  double  kernel_ms;       /* CUDA-event time of the evaluation kernel (max over this process's GPUs) */
  double  device_ms;       /* kernel incl. cross-GPU sum and host delivery (or + all-reduce + result copy), CUDA events on the evaluation stream */
  int64_t launches;        /* kernels launched by the last call (all GPUs of this process) */
  int64_t total_launches;  /* kernels launched since so_init */
  int64_t total_evals;     /* so_eval_* calls since so_init */
  /* ABI 2 */
  double  finish_ms;       /* last call: time the folding block spent in the cross-GPU exchange + host delivery (device clock;
                              includes waiting for the slowest GPU); part of kernel_ms when finished_on_device */
  double  sum_kernel_ms;   /* kernel_ms, device_ms, finish_ms summed over the calls since so_init / so_stats_reset, */
  double  sum_device_ms;   /*   so that a caller timing many evaluations need not ask after each one */
  double  sum_finish_ms;
  double  min_kernel_ms;   /* fastest kernel_ms since so_init / so_stats_reset (0: none yet) */
  int64_t finished_on_device; /* last call: 1 = cross-GPU sum and host delivery done by the evaluation kernel itself (results up to
                              16384 doubles), 0 = NCCL all-reduce + device-to-host copy */
} so_call_stats;
/* Event times are collected lazily: so_stats waits for the last evaluation's events (microseconds), the evaluation calls do not. */
int so_stats(so_ctx* ctx, so_call_stats* out);
int so_stats_reset(so_ctx* ctx);   /* zero the sum_* / min_* counters */

#ifdef __cplusplus
}
#endif
#endif /* SCALAOPT_CUDA_H */
