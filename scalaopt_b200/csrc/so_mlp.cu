// so_mlp.cu — the reference's FFNeuralNetworkTrainer as a device family: FFNeuralNetwork(Vector(f, h, 1))
// with logistic inner neurons, forward + backward per observation inside the kernel, one thread per row.
//   reference: std-apps/.../learning/nnet/FFNeuralNetwork.scala:114-189 (forward/backward), Neuron.scala:40-73,
//              FFNeuralNetworkTrainer.scala:59-69 (apply / gradient), :100-131 (jacobianAndResidualsMatrix)
//   MLP_MSE: linear output neuron, loss = residual^2 (FFNeuralNetwork.scala:70-72 — not halved, unlike MSEFunction),
//            error = residual = output - target
//   MLP_CE : logistic output neuron, cross-entropy loss (:73-81), error = residual = output - target (:174-179)
// Per row: e_j = w_j0 + x . w_j, o_j = sigma(e_j); E = v_0 + o . v; the derivative row
//   u = ( v_j o_j (1 - o_j) (1, x) for j < h | (1, o) )    (n = h (f + 1) + h + 1 entries, the reference's weight order)
// gives gradient_i = error u (Neuron.gradient) and the Jacobian row u = gradient / residual (Neuron.jacobian: for a
// zero residual the reference keeps the gradient itself, which is a zero row).  K1 accumulates (loss, error u), K2
// the Gram of (u | residual) in registers — n <= 9 for K2 (55 accumulators), e.g. the Vector(2, 2, 1) of the
// reference's own spec; the weights sit in the kernel-argument (constant) bank.  Shapes: f <= 4, h <= 4.
// K3 / K4 for these families are composed from K1 launches by the API layer, exactly as the reference composes them
// (LineSearchPoint: f(x + a d), grad . d; dirHessian: forward difference of gradients, Trainer.scala:76-80).
#include <algorithm>

#include "so_device.cuh"
#include "so_internal.h"

namespace so {
namespace {

constexpr int kMlpThreads = 256;

template <int F, int H> struct MlpW { double w[H][F + 1]; double v[H + 1]; };

template <int F, int H, bool CE, int MODE>      // MODE 0: loss, 1: loss + gradient, 2: normal equations
__global__ void __launch_bounds__(kMlpThreads, (MODE == 2 || H * (F + 1) + H + 1 > 16) ? 2 : 4)
k_mlp(const double* __restrict__ X, const double* __restrict__ y, int64_t rows, const __grid_constant__ MlpW<F, H> W,
      double* __restrict__ partials, double* __restrict__ out, unsigned int* ticket) {
  constexpr int N = H * (F + 1) + H + 1, T = N * (N + 1) / 2;
  constexpr int Q = MODE == 0 ? 1 : (MODE == 1 ? 1 + N : T + N + 1);
  __shared__ double s_red[(kMlpThreads / 32) * Q];
  exp_table_init();
  if (CE) log_table_init();
  __syncthreads();
  double acc[Q];
#pragma unroll
  for (int q = 0; q < Q; ++q) acc[q] = 0.0;
  for (int64_t row = (int64_t)blockIdx.x * kMlpThreads + threadIdx.x; row < rows; row += (int64_t)gridDim.x * kMlpThreads) {
    double x[F];
#pragma unroll
    for (int i = 0; i < F; ++i) x[i] = ld_stream(X + row * F + i);
    const double yv = ld_stream(y + row);
    double o[H], E = 0.0, ej[H], exj[H];
    bool ok = true;
#pragma unroll
    for (int j = 0; j < H; ++j) {
      double e = 0.0;
#pragma unroll
      for (int i = 0; i < F; ++i) e = fma(x[i], W.w[j][1 + i], e);
      e += W.w[j][0];                               // weights(0) + (inputs dot weights.tail), Neuron.scala:42-43
      ej[j] = e;
      ok = ok && exp_in_range(e);
    }
    if (ok) {                                       // one range check for the hidden layer: the H exponentials run branch-free
#pragma unroll
      for (int j = 0; j < H; ++j) exj[j] = so_exp_fast(-fabs(ej[j]));
    } else {
#pragma unroll
      for (int j = 0; j < H; ++j) exj[j] = exp_slow(-fabs(ej[j]));
    }
#pragma unroll
    for (int j = 0; j < H; ++j) {
      const double inv = so_rcp_fast(1.0 + exj[j]);
      o[j] = ej[j] >= 0.0 ? inv : exj[j] * inv;
      E = fma(o[j], W.v[1 + j], E);
    }
    E += W.v[0];
    double res, loss;
    if (CE) {
      const double ex = so_exp(-fabs(E));
      const double inv = so_rcp_fast(1.0 + ex);
      res = (E >= 0.0 ? inv : ex * inv) - yv;
      // -y log(o/y) - (1-y) log((1-o)/(1-y)) = softplus(-E) + (1-y) E + [y log y + (1-y) log(1-y)]
      loss = ((ex >= 0.0 && ex <= 1.0) ? so_log1p_unit(ex) : log1p(ex)) + fmax(-E, 0.0) + (1.0 - yv) * E;
      if (yv > 0.0 && yv < 1.0) loss += yv * log(yv) + (1.0 - yv) * log(1.0 - yv);
    } else {
      res = E - yv;
      loss = res * res;
    }
    if (MODE == 0) { acc[0] += loss; continue; }
    double u[N];
#pragma unroll
    for (int j = 0; j < H; ++j) {
      const double dj = W.v[1 + j] * (o[j] * (1.0 - o[j]));
      u[j * (F + 1)] = dj;
#pragma unroll
      for (int i = 0; i < F; ++i) u[j * (F + 1) + 1 + i] = dj * x[i];
    }
    u[H * (F + 1)] = 1.0;
#pragma unroll
    for (int j = 0; j < H; ++j) u[H * (F + 1) + 1 + j] = o[j];
    if (MODE == 1) {
      acc[0] += loss;
#pragma unroll
      for (int k = 0; k < N; ++k) acc[1 + k] = fma(res, u[k], acc[1 + k]);
    } else {
      const double live = res != 0.0 ? 1.0 : 0.0;   // Neuron.jacobian keeps the (zero) gradient when the residual is 0
      int q = 0;
#pragma unroll
      for (int a = 0; a < N; ++a) {
        const double ua = u[a] * live;
#pragma unroll
        for (int b = a; b < N; ++b) { acc[q] = fma(ua, u[b], acc[q]); ++q; }
      }
#pragma unroll
      for (int a = 0; a < N; ++a) acc[T + a] = fma(u[a], res, acc[T + a]);
      acc[T + N] = fma(res, res, acc[T + N]);
    }
  }
  grid_reduce<Q>(acc, s_red, partials, out, nullptr, ticket);
}

template <int F, int H, bool CE, int MODE>
int launch_one(const Shard& s, const double* p) {
  MlpW<F, H> W;
  for (int j = 0; j < H; ++j)
    for (int i = 0; i <= F; ++i) W.w[j][i] = p[j * (F + 1) + i];
  for (int j = 0; j <= H; ++j) W.v[j] = p[H * (F + 1) + j];
  constexpr int N = H * (F + 1) + H + 1;
  constexpr int Q = MODE == 0 ? 1 : (MODE == 1 ? 1 + N : N * (N + 1) / 2 + N + 1);
  auto kern = k_mlp<F, H, CE, MODE>;
  int per_sm = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kMlpThreads, 0);
  if (per_sm < 1) per_sm = 1;
  const int64_t want = std::max<int64_t>(1, (s.rows + kMlpThreads - 1) / kMlpThreads);
  const int grid = (int)std::min<int64_t>(want, (int64_t)per_sm * s.sm_count);
  if ((size_t)grid * Q > s.partials_doubles || (size_t)Q > s.out_doubles) return SO_ENOMEM;
  kern<<<grid, kMlpThreads, 0, s.stream>>>(s.X, s.y, s.rows, W, s.partials, s.out, s.ticket);
  return cudaGetLastError() == cudaSuccess ? 1 : SO_ECUDA;
}

template <int F, int H, bool CE>
int launch_shape(Kind kind, const Shard& s, const double* p, int want_grad) {
  constexpr int N = H * (F + 1) + H + 1;
  if (kind == kLossGrad) return want_grad ? launch_one<F, H, CE, 1>(s, p) : launch_one<F, H, CE, 0>(s, p);
  if (kind == kNormalEq) {
    if constexpr (N <= 9) return launch_one<F, H, CE, 2>(s, p);
    else return SO_EUNSUPPORTED;
  }
  return SO_EUNSUPPORTED;
}

template <bool CE>
int dispatch(Kind kind, const Shard& s, int h, const double* p, int want_grad) {
#define SO_MLP(FV, HV) if (s.f == FV && h == HV) return launch_shape<FV, HV, CE>(kind, s, p, want_grad);
  SO_MLP(1, 1) SO_MLP(1, 2) SO_MLP(1, 3) SO_MLP(1, 4)
  SO_MLP(2, 1) SO_MLP(2, 2) SO_MLP(2, 3) SO_MLP(2, 4)
  SO_MLP(3, 1) SO_MLP(3, 2) SO_MLP(3, 3) SO_MLP(3, 4)
  SO_MLP(4, 1) SO_MLP(4, 2) SO_MLP(4, 3) SO_MLP(4, 4)
#undef SO_MLP
  return SO_EUNSUPPORTED;
}

}  // namespace

size_t mlp_partials_needed(Kind kind, int n, int sm_count) {
  return (size_t)sm_count * 8 * (size_t)std::max(out_doubles(kind, n, 1), 16);
}

// scalar output and f, h <= 4: the register-resident template kernels above; any other catalogue shape (several outputs,
// soft-max, wider layers, no hidden layer): so_net.cu
int launch_mlp(Kind kind, const Shard& s, int family, const double* p, int n, int want_grad) {
  const int h = net_hidden_units(family, s.f, s.ny, n);
  if (h < 0) return SO_EINVAL;
  const bool small_k2 = kind != kNormalEq || n <= 9;
  if (s.ny == 1 && family != SO_FAMILY_SOFTMAX && s.f <= 4 && h >= 1 && h <= 4 && small_k2)
    return family == SO_FAMILY_MLP_CE ? dispatch<true>(kind, s, h, p, want_grad) : dispatch<false>(kind, s, h, p, want_grad);
  return launch_net(kind, s, family, h, p, n, want_grad);
}

}  // namespace so
