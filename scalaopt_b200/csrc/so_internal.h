// so_internal.h — launch interface between the C-ABI layer (so_api.cu) and the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/scalaopt_cuda.h"

namespace so {

// ---- control block of an evaluation stream (device memory; one per delivery mode per device) ------------------------
// Every evaluation kernel gets a pointer to it as its `ticket` argument: the first four words are the counters the
// kernels always used (last-block ticket, grid barrier of k_gram_ws); behind them sits the description of what the block
// that folds the per-block partials does with the result (so_finish, so_device.cuh):
//   mode 0  leave the sums in `out` (the host sums across GPUs with NCCL and copies them back — large results)
//   mode 1  the cross-GPU SUM and the delivery to the host happen inside the same kernel: the folding block stores its Q
//           sums into every rank's mailbox over NVLink (peer-mapped pointers: cudaDeviceEnablePeerAccess in one process,
//           cudaIpcOpenMemHandle between ranks), raises a flag there, waits for the flags of all ranks in its own
//           mailbox, adds the `world` slots in RANK ORDER (every rank gets the same bits) and writes the totals plus a
//           sequence number into mapped pinned host memory, where the calling thread is polling.  No collective launch,
//           no device-to-host copy, no stream synchronisation on the way (spark-apps/.../SparkDataSetConverter.scala:42-45
//           does this sum as rdd.aggregate on the driver).
// Mailboxes are double-buffered on the parity of the sequence number: a rank can be at most one call ahead of a peer
// (it cannot finish call c+1 before every peer has posted its sums for c+1, i.e. has finished reading call c).
constexpr int kMaxRanks = 8;
constexpr int kMailCap = 16384;         // doubles per (parity, source rank) slot: results up to this size take mode 1
struct Ctl {
  unsigned int ticket[4];
  int mode, world, rank, cap;
  unsigned long long* seq;              // device word shared by the control blocks of a device: exchanges finished
  double* host_out;                     // mapped pinned: cap doubles of result + 4 doubles of device timestamps
  unsigned long long* host_flag;        // mapped pinned: sequence number of the result host_out holds
  unsigned long long* host_err;         // mapped pinned: set when a peer did not show up within timeout_ns
  unsigned long long timeout_ns;
  double* slots[kMaxRanks];             // slots[r] = rank r's mailbox data  [2][kMaxRanks][cap]
  unsigned long long* flags[kMaxRanks]; // flags[r] = rank r's mailbox flags [2][kMaxRanks]
};

// One GPU's resident shard (or a slice of it) plus the workspace the kernels reduce through.
struct Shard {
  const double* X = nullptr;   // rows x f, row-major
  const double* y = nullptr;   // rows
  int64_t rows = 0;
  int f = 0;
  int ny = 1;                  // outputs per row (y is rows x ny, row-major); > 1 only for the network families
  // workspace owned by the context (per device)
  double* partials = nullptr;  // per-block partial sums
  size_t partials_doubles = 0;
  double* out = nullptr;       // reduced result of the running evaluation
  size_t out_doubles = 0;
  unsigned int* ticket = nullptr;   // -> Ctl (above)
  // set to false by a launcher whose kernel does not end in so_finish (k_gram_ws): the API layer then runs k_finish
  mutable bool device_finish = true;
  cudaStream_t stream = nullptr;
  int sm_count = 148;
  // f == 1 only: range of the inputs of this slice when it is known and all of them are finite (launch_range)
  bool has_range = false;
  double tmin = 0.0, tmax = 0.0;
};

constexpr int SO_QR_WIDE_MAX_N = 8;     // TSQR of the linear-predictor families: n + 1 <= 9 columns per thread (so_wide_qr.cu)
constexpr int SO_QR_COLS_MAX_N = 63;    // ... and up to 64 columns with one triangle per warp in shared memory (so_wide_qrw.cu)

enum Kind { kLossGrad = 1, kNormalEq = 2, kLine = 3, kHessVec = 4, kNll = 5, kQr = 6 };

// number of doubles each evaluation leaves in Shard::out (before the 1/m normalisation, which the
// API layer applies after the cross-GPU sum):
//   kLossGrad : [ loss_sum, g_0 .. g_{n-1} ]                          (loss_sum = sum r^2/2 or sum CE)
//   kNormalEq : [ JtJ upper triangle packed row-major (a<=b), Jtr_0..Jtr_{n-1}, rtr ]
//   kLine     : [ phi_sum_0..phi_sum_{k-1}, dphi_sum_0..dphi_sum_{k-1} ]
//   kHessVec  : [ Hd_sum_0 .. Hd_sum_{n-1} ]
//   kQr       : [ R, (n+1) x (n+1) row-major, upper triangular: the R factor of this shard's rows of [J | r] ]
inline int out_doubles(Kind kind, int n, int k) {
  switch (kind) {
    case kLossGrad: return 1 + n;
    case kNormalEq: return n * (n + 1) / 2 + n + 1;
    case kLine: return 2 * k;
    case kHessVec: return n;
    case kNll: return 1;
    case kQr: return (n + 1) * (n + 1);
  }
  return 0;
}

// required size of Shard::partials for an evaluation (doubles)
size_t scalar_partials_needed(Kind kind, int n, int k, int sm_count);
size_t wide_partials_needed(Kind kind, int family, int n, int f, int k, int sm_count);

// Each launcher enqueues its kernel(s) on shard.stream and returns the number of kernels launched
// (>0) or a negative SO_E* code.  `want_grad` = 0 makes kLossGrad a loss-only pass.
int launch_scalar(Kind kind, const Shard& s, int family, const double* p, const double* d, int n,
                  const double* alpha, int k, int want_grad);
int launch_wide(Kind kind, const Shard& s, int family, const double* p, const double* d, int n,
                const double* alpha, int k, int want_grad);

// FFNeuralNetwork(Vector(f, h, 1)) families (so_mlp.cu): kLossGrad and kNormalEq
int launch_mlp(Kind kind, const Shard& s, int family, const double* p, int n, int want_grad);
int launch_net(Kind kind, const Shard& s, int family, int h, const double* p, int n, int want_grad);   // so_net.cu: any (f, h, K)
size_t mlp_partials_needed(Kind kind, int n, int sm_count);

// min / max / non-finite count of the X stream of an f == 1 shard -> buf3 (device, order-preserving integer keys; the
// caller presets buf3 = {~0, 0, 0} and decodes with range_decode).  One pass at 8 B/row, once per data-set state.
int launch_range(const Shard& s, unsigned long long* buf3);
bool range_decode(const unsigned long long* host3, double* tmin, double* tmax);   // false: no rows or a non-finite input

// sum_i -2 log pdf(p, x_i) -> out[0]; reads the X stream only
int launch_nll(const Shard& s, int pdf, const double* p, int n, const double* consts, int nconsts);

// padded width KB of the kLine result of launch_scalar: phi at out[0,k), dphi at out[KB, KB+k)
int scalar_line_block(int k);

// the cross-GPU sum + host delivery of `q` doubles at s.out as a launch of its own (kernels that do not end in so_finish)
int launch_finish(const Shard& s, int q);

// synthetic data
int launch_generate(const Shard& s, double* X, double* y, int family, const double* p_true, int n, uint64_t seed,
                    double noise_sigma, int64_t row_offset, int64_t total_rows);

inline bool family_is_scalar_t(int family) {
  return family == SO_FAMILY_EXPDECAY || family == SO_FAMILY_GAUSS || family == SO_FAMILY_GAUSS_EXPBKG ||
         family == SO_FAMILY_POLY;
}
inline int family_nparams(int family, int f, int n_hint) {
  switch (family) {
    case SO_FAMILY_LINEAR: case SO_FAMILY_LOGISTIC: return f + 1;
    case SO_FAMILY_LINEAR_NOINT: return f;
    case SO_FAMILY_EXPDECAY: return 2;
    case SO_FAMILY_GAUSS: return 3;
    case SO_FAMILY_GAUSS_EXPBKG: return 5;
    case SO_FAMILY_POLY: return n_hint;
    case SO_FAMILY_MLP_MSE: case SO_FAMILY_MLP_CE: case SO_FAMILY_SOFTMAX: return n_hint;     // checked by net_hidden_units
  }
  return -1;
}
inline bool family_is_mlp(int family) { return family == SO_FAMILY_MLP_MSE || family == SO_FAMILY_MLP_CE; }
// the FFNeuralNetwork families (so_mlp.cu, so_net.cu): one hidden layer of h >= 1 logistic neurons, or none (SOFTMAX)
inline bool family_is_net(int family) { return family_is_mlp(family) || family == SO_FAMILY_SOFTMAX; }
// hidden neurons from the weight count: n = h (f + 1) + K (h + 1); SOFTMAX: n = K (f + 1) and h = 0.  -1: no such network
inline int net_hidden_units(int family, int f, int K, int n) {
  if (family == SO_FAMILY_SOFTMAX) return (K >= 2 && n == K * (f + 1)) ? 0 : -1;
  if (n <= K || (n - K) % (f + 1 + K) != 0) return -1;
  const int h = (n - K) / (f + 1 + K);
  return h >= 1 ? h : -1;
}

}  // namespace so
