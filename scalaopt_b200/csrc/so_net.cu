// so_net.cu — the reference's FFNeuralNetworkTrainer for ANY catalogue shape (f <= 16 inputs, h <= 16 hidden logistic neurons
// or none, K <= 8 outputs): multi-output targets, soft-max output layer, the shapes of the reference's own spec
// (FFNeuralNetwork(Vector(Iris.nInputs, 5, Iris.nClasses)) and Vector(2, 3), FFNeuralNetworkTrainerSpec.scala:222-260,353-376).
// The scalar-output shapes with f, h <= 4 keep their register-resident template kernel (so_mlp.cu); everything else comes here.
//
// What is restated per observation (std-apps/.../learning/nnet/):
//   forward   Neuron.activate (Neuron.scala:40-49): excitation = w_0 + inputs . w_tail; inner neurons LogisticFunction;
//             output layer: LinearFunction (MSE), LogisticFunction (CE, K = 1) or SoftMaxFunction (CE, K > 1:
//             o_k = exp(z_k - max z) / sum, FFNeuralNetwork.scala:143-158)
//   loss      MSE: sum_k residual_k^2 (:78-80, not halved); CE, K = 1: :81-88; CE, K > 1: -sum_k y_k log(o_k / sum o),
//             where an output that is exactly 0 resets the running sum to 0 (:89-103 — the fold returns 0.0 there; kept)
//   backward  outer layer (:161-181): MSE error_k = (o_k - y_k) act'(o_k), residual_k = o_k - y_k; CE: error_k = residual_k =
//             (sum y) o_k - y_k (K > 1) or o_k - y_k; inner layer (Neuron.scala:58-63): error_j = (sum_k error_k v_kj) o_j (1 - o_j)
//   gradient  Neuron.gradient (:65): (1, inputs) error              -> K1: loss and sum of gradients
//   jacobian  Neuron.jacobian (:67-73): (1, inputs) error / residual of the neuron unless that is 0 or NaN; an inner neuron's
//             "residual" is the FIRST output neuron's (:62 `residual = neurons.head.residual`); network residual (:105-110):
//             the output neuron's (K = 1) or ||residuals||_2                        -> K2: J^T J, J^T r, r^T r
// One thread per observation; a row's contribution to each of the Q sums is added over the warp's 32 rows with a
// butterfly and accumulated by lane 0 into the warp's accumulators in shared memory (no per-thread accumulator arrays:
// Vector(4, 5, 3) has 43 weights, its Gram 946 entries).  Deterministic: fixed row -> lane mapping, warp order, block order.
// This kernel is instruction-bound by construction (11 instructions per sum and 32 rows); it exists so that the catalogue
// covers what the reference trains, the data sets of which are small (Iris: 150 rows; the spec's synthetic sets: 10^4).
#include <algorithm>

#include "so_device.cuh"
#include "so_internal.h"

namespace so {
namespace {

constexpr int kNetThreads = 128, kNetWarps = kNetThreads / 32;
constexpr int kNetMaxF = 16, kNetMaxH = 16, kNetMaxK = SO_MAX_OUTPUTS;
constexpr int kNetMaxW = kNetMaxH * (kNetMaxF + 1) + kNetMaxK * (kNetMaxH + 1);

struct NetW { int f, h, K, ce, n; double w[kNetMaxW]; };

template <int MODE>      // 0: loss, 1: loss + gradient, 2: normal equations
__global__ void __launch_bounds__(kNetThreads)
k_net(const double* __restrict__ X, const double* __restrict__ Y, int64_t rows, const __grid_constant__ NetW W,
      double* __restrict__ partials, double* __restrict__ out, unsigned int* ticket) {
  extern __shared__ double net_smem[];              // weights | kNetWarps x Q accumulators
  const int f = W.f, h = W.h, K = W.K, n = W.n, ce = W.ce;
  const int T = n * (n + 1) / 2;
  const int Q = MODE == 0 ? 1 : (MODE == 1 ? 1 + n : T + n + 1);
  double* s_w = net_smem;
  double* s_acc = net_smem + kNetMaxW;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* acc = s_acc + (size_t)warp * Q;
  exp_table_init();
  log_table_init();
  for (int i = threadIdx.x; i < n; i += kNetThreads) s_w[i] = W.w[i];
  for (int i = threadIdx.x; i < kNetWarps * Q; i += kNetThreads) s_acc[i] = 0.0;
  __syncthreads();
  const int nin = h > 0 ? h : f;                     // inputs of the output layer
  const double* wo = s_w + h * (f + 1);              // output neurons: (bias, v_1 .. v_nin) each
  const int64_t nblk = (rows + 31) / 32;
  for (int64_t blk = (int64_t)blockIdx.x * kNetWarps + warp; blk < nblk; blk += (int64_t)gridDim.x * kNetWarps) {
    const int64_t row = blk * 32 + lane;
    const bool valid = row < rows;
    double x[kNetMaxF], o[kNetMaxH], z[kNetMaxK], res[kNetMaxK], err[kNetMaxK], eh[kNetMaxH];
    for (int i = 0; i < f; ++i) x[i] = valid ? ld_stream(X + row * f + i) : 0.0;
    // ---- forward ----
    for (int j = 0; j < h; ++j) {
      const double* wj = s_w + j * (f + 1);
      double e = 0.0;
      for (int i = 0; i < f; ++i) e = fma(x[i], wj[1 + i], e);
      e += wj[0];
      const double ex = so_exp(-fabs(e));
      const double inv = so_rcp_fast(1.0 + ex);
      o[j] = e >= 0.0 ? inv : ex * inv;
    }
    const double* in = h > 0 ? o : x;
    double zmax = -1.0e308;
    for (int k = 0; k < K; ++k) {
      const double* vk = wo + k * (nin + 1);
      double e = 0.0;
      for (int i = 0; i < nin; ++i) e = fma(in[i], vk[1 + i], e);
      e += vk[0];
      z[k] = e;
      zmax = fmax(zmax, e);
    }
    double loss = 0.0;
    if (!ce) {                                       // LinearFunction outputs, squared residuals
      for (int k = 0; k < K; ++k) {
        const double yk = valid ? ld_stream(Y + row * K + k) : 0.0;
        res[k] = z[k] - yk;
        err[k] = res[k];                             // (o - y) * LinearFunction.derivative = 1
        loss = fma(res[k], res[k], loss);
      }
    } else if (K == 1) {                             // LogisticFunction output, FFNeuralNetwork.scala:81-88
      const double yv = valid ? ld_stream(Y + row) : 0.0;
      const double E = z[0];
      const double ex = so_exp(-fabs(E));
      const double inv = so_rcp_fast(1.0 + ex);
      res[0] = (E >= 0.0 ? inv : ex * inv) - yv;
      err[0] = res[0];
      loss = ((ex >= 0.0 && ex <= 1.0) ? so_log1p_unit(ex) : log1p(ex)) + fmax(-E, 0.0) + (1.0 - yv) * E;
      if (yv > 0.0 && yv < 1.0) loss += yv * log(yv) + (1.0 - yv) * log(1.0 - yv);
    } else {                                         // SoftMaxFunction outputs
      double sum = 0.0;
      for (int k = 0; k < K; ++k) { z[k] = so_exp(z[k] - zmax); sum += z[k]; }     // probs, :150-152
      double osum = 0.0, ysum = 0.0;
      for (int k = 0; k < K; ++k) { z[k] = z[k] / sum; osum += z[k]; }              // outputs o_k, :153-156
      double entropy = 0.0;
      for (int k = 0; k < K; ++k) {
        const double yk = valid ? ld_stream(Y + row * K + k) : 0.0;
        ysum += yk;
        res[k] = yk;                                 // target for now
        if (z[k] > 0.0) entropy -= yk * log(z[k] / osum); else entropy = 0.0;       // the fold's `else 0.0`, :96-100
      }
      loss = osum > 0.0 ? entropy : 0.0;
      for (int k = 0; k < K; ++k) { res[k] = ysum * z[k] - res[k]; err[k] = res[k]; }   // :168-173
    }
    if (!valid) { loss = 0.0; for (int k = 0; k < K; ++k) { res[k] = 0.0; err[k] = 0.0; } }
    if (MODE == 0) {
      const double s = warp_sum(loss);
      if (lane == 0) acc[0] += s;
      continue;
    }
    // ---- backward: inner layer ----
    for (int j = 0; j < h; ++j) {
      double ld = 0.0;
      for (int k = 0; k < K; ++k) ld = fma(err[k], wo[k * (nin + 1) + 1 + j], ld);
      eh[j] = ld * (o[j] * (1.0 - o[j]));
    }
    if (MODE == 1) {
      double s = warp_sum(loss);
      if (lane == 0) acc[0] += s;
      int q = 1;
      for (int j = 0; j < h; ++j)
        for (int i = 0; i <= f; ++i) { s = warp_sum(eh[j] * (i == 0 ? 1.0 : x[i - 1])); if (lane == 0) acc[q] += s; ++q; }
      for (int k = 0; k < K; ++k)
        for (int i = 0; i <= nin; ++i) { s = warp_sum(err[k] * (i == 0 ? 1.0 : in[i - 1])); if (lane == 0) acc[q] += s; ++q; }
    } else {
      // Jacobian row (Neuron.jacobian): weight a of neuron N: (1, inputs)_a error_N / residual_N when that is neither 0 nor NaN
      double r2 = 0.0;
      for (int k = 0; k < K; ++k) r2 = fma(res[k], res[k], r2);
      const double rnet = K == 1 ? res[0] : sqrt(r2);                // FFNeuralNetwork.residual, :105-110
      auto jac = [&](int a) -> double {
        if (a < h * (f + 1)) {
          const int j = a / (f + 1), i = a - j * (f + 1);
          const double rr = res[0];                                  // neurons.head.residual, Neuron.scala:62
          const double g = eh[j] * (i == 0 ? 1.0 : x[i - 1]);
          return (rr != rr || rr == 0.0) ? g : g / rr;
        }
        const int b = a - h * (f + 1), k = b / (nin + 1), i = b - k * (nin + 1);
        const double g = err[k] * (i == 0 ? 1.0 : in[i - 1]);
        return (res[k] != res[k] || res[k] == 0.0) ? g : g / res[k];
      };
      double J[kNetMaxW];                                            // local memory: the row is formed once
      for (int a = 0; a < n; ++a) J[a] = jac(a);
      int q = 0;
      for (int a = 0; a < n; ++a) {
        const double ja = J[a];
        for (int b = a; b < n; ++b) { const double s = warp_sum(ja * J[b]); if (lane == 0) acc[q] += s; ++q; }
      }
      for (int a = 0; a < n; ++a) { const double s = warp_sum(J[a] * rnet); if (lane == 0) acc[T + a] += s; }
      { const double s = warp_sum(rnet * rnet); if (lane == 0) acc[T + n] += s; }
    }
  }
  __syncthreads();
  // block sums in warp order, per-block partials, last block in block order
  for (int q = threadIdx.x; q < Q; q += kNetThreads) {
    double s = 0.0;
    for (int w = 0; w < kNetWarps; ++w) s += s_acc[(size_t)w * Q + q];
    partials[(size_t)blockIdx.x * Q + q] = s;
  }
  __threadfence();
  __syncthreads();
  __shared__ unsigned int s_last_net;
  if (threadIdx.x == 0) s_last_net = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
  __syncthreads();
  if (s_last_net) {
    __threadfence();
    for (int q = threadIdx.x; q < Q; q += kNetThreads) {
      double s = 0.0;
      for (unsigned b = 0; b < gridDim.x; ++b) s += __ldcg(&partials[(size_t)b * Q + q]);
      out[q] = s;
    }
    if (threadIdx.x == 0) *ticket = 0u;
    so_finish(ticket, out, Q);
  }
}

template <int MODE>
int launch_net_mode(const Shard& s, const NetW& W) {
  auto kern = k_net<MODE>;
  const int n = W.n;
  const int Q = MODE == 0 ? 1 : (MODE == 1 ? 1 + n : n * (n + 1) / 2 + n + 1);
  const size_t smem = sizeof(double) * ((size_t)kNetMaxW + (size_t)kNetWarps * Q);
  if (smem > 200 * 1024) return SO_EUNSUPPORTED;
  static bool attr[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 64 && !attr[dev]) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) return SO_ECUDA;
    attr[dev] = true;
  }
  int per_sm = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kNetThreads, smem);
  if (per_sm < 1) return SO_EUNSUPPORTED;
  per_sm = std::min(per_sm, 8);
  const int64_t nblk = (s.rows + 31) / 32;
  const int64_t want = std::max<int64_t>(1, (nblk + kNetWarps - 1) / kNetWarps);
  const int grid = (int)std::min<int64_t>(want, (int64_t)per_sm * s.sm_count);
  if ((size_t)grid * Q > s.partials_doubles || (size_t)Q > s.out_doubles) return SO_ENOMEM;
  kern<<<grid, kNetThreads, smem, s.stream>>>(s.X, s.y, s.rows, W, s.partials, s.out, s.ticket);
  return cudaGetLastError() == cudaSuccess ? 1 : SO_ECUDA;
}

}  // namespace

int launch_net(Kind kind, const Shard& s, int family, int h, const double* p, int n, int want_grad) {
  if (s.f > kNetMaxF || h > kNetMaxH || s.ny > kNetMaxK || n > kNetMaxW) return SO_EUNSUPPORTED;
  if (family == SO_FAMILY_MLP_MSE && s.ny > 1) return SO_EUNSUPPORTED;   // `???` in the reference too (FFNeuralNetwork.scala:157)
  NetW W;
  W.f = s.f; W.h = h; W.K = s.ny; W.n = n;
  W.ce = family == SO_FAMILY_MLP_MSE ? 0 : 1;
  for (int i = 0; i < n; ++i) W.w[i] = p[i];
  if (kind == kLossGrad) return want_grad ? launch_net_mode<1>(s, W) : launch_net_mode<0>(s, W);
  if (kind == kNormalEq) return launch_net_mode<2>(s, W);
  return SO_EUNSUPPORTED;
}

}  // namespace so
