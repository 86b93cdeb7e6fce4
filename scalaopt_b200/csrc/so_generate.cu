// so_generate.cu — on-device synthetic observations, same counter-based specification as
// oracle/oracle.c:orc_generate:  u = splitmix64(seed ^ ctr) >> 11 * 2^-53, ctr = global_row*(f+1)+col,
// normals by Box-Muller with the second uniform drawn at ctr ^ 2^63.
//   linear families : x_j ~ U[0,1)                  (LevenbergMarquardtSpec.scala:83-86), y = f(p*, x) + sigma N(0,1)
//   logistic        : x_j ~ N(0,1), y ~ Bernoulli(o) (FFNeuralNetworkTrainerSpec.scala:202-207)
//   scalar-t        : t = (row + 1/2) / total_rows,  y = f(p*, t) + sigma N(0,1)
// Generation is not on the evaluation hot path; it uses libdevice exp/log/cos, not so_exp.
#include <algorithm>

#include "so_internal.h"

namespace so {
namespace {

__device__ __forceinline__ uint64_t splitmix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
__device__ __forceinline__ double u01(uint64_t seed, uint64_t ctr) {
  return (double)(splitmix64(seed ^ ctr) >> 11) * 0x1.0p-53;
}
__device__ __forceinline__ double n01(uint64_t seed, uint64_t ctr) {
  const double u1 = u01(seed, ctr), u2 = u01(seed, ctr ^ 0x8000000000000000ULL);
  return sqrt(-2.0 * log(1.0 - u1)) * cos(6.283185307179586476925 * u2);
}

struct GenParams { double p[SO_MAX_FEATURES + 1]; };

__global__ void k_gen_inputs(double* __restrict__ X, int64_t rows, int f, int family, uint64_t seed,
                             int64_t row_offset, int64_t total_rows) {
  const int64_t total = rows * f;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = idx / f;
    const int col = (int)(idx - row * f);
    const uint64_t gi = (uint64_t)(row_offset + row);
    const uint64_t ctr = gi * (uint64_t)(f + 1) + (uint64_t)col;
    double v;
    if (family == SO_FAMILY_LINEAR || family == SO_FAMILY_LINEAR_NOINT) v = u01(seed, ctr);
    else if (family == SO_FAMILY_LOGISTIC) v = n01(seed, ctr);
    else v = ((double)gi + 0.5) / (double)total_rows;
    X[idx] = v;
  }
}

__device__ double gen_model(int family, const double* p, int n, const double* x, int f) {
  switch (family) {
    case SO_FAMILY_LINEAR: { double a = 0.0; for (int j = 0; j < f; ++j) a += p[j + 1] * x[j]; return p[0] + a; }
    case SO_FAMILY_LINEAR_NOINT: { double a = 0.0; for (int j = 0; j < f; ++j) a += p[j] * x[j]; return a; }
    case SO_FAMILY_LOGISTIC: { double a = 0.0; for (int j = 0; j < f; ++j) a += x[j] * p[j + 1]; return 1.0 / (1.0 + exp(-(p[0] + a))); }
    case SO_FAMILY_EXPDECAY: return p[0] * exp(p[1] * x[0]);
    case SO_FAMILY_GAUSS: { double d = x[0] - p[1]; return p[0] * exp(-(d * d) / (2.0 * p[2] * p[2])); }
    case SO_FAMILY_GAUSS_EXPBKG: { double d = x[0] - p[1]; return p[0] * exp(-(d * d) / (2.0 * p[2] * p[2])) + p[3] * exp(-p[4] * x[0]); }
    case SO_FAMILY_POLY: { double a = 0.0, tk = 1.0; for (int k = 0; k < n; ++k) { a += p[k] * tk; tk *= x[0]; } return a; }
  }
  return 0.0;
}

__global__ void k_gen_outputs(const double* __restrict__ X, double* __restrict__ y, int64_t rows, int f, int family,
                              const __grid_constant__ GenParams prm, int n, uint64_t seed, double sigma,
                              int64_t row_offset) {
  for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < rows;
       row += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t gi = (uint64_t)(row_offset + row);
    const uint64_t ctr = gi * (uint64_t)(f + 1) + (uint64_t)f;
    const double mu = gen_model(family, prm.p, n, X + row * f, f);
    y[row] = (family == SO_FAMILY_LOGISTIC) ? ((u01(seed, ctr) < mu) ? 1.0 : 0.0) : mu + sigma * n01(seed, ctr);
  }
}

}  // namespace

int launch_generate(const Shard& s, double* X, double* y, int family, const double* p_true, int n, uint64_t seed,
                    double noise_sigma, int64_t row_offset, int64_t total_rows) {
  if (s.rows == 0) return 0;
  if (family_is_mlp(family)) return SO_EUNSUPPORTED;      // no synthetic generator for the MLP families
  if (n < 1 || n > SO_MAX_FEATURES + 1) return SO_EINVAL;
  GenParams prm;
  for (int i = 0; i < n; ++i) prm.p[i] = p_true[i];
  const int threads = 256;
  const int64_t elems = s.rows * s.f;
  int grid_x = (int)std::min<int64_t>((elems + threads - 1) / threads, (int64_t)s.sm_count * 16);
  k_gen_inputs<<<grid_x, threads, 0, s.stream>>>(X, s.rows, s.f, family, seed, row_offset, total_rows);
  int grid_y = (int)std::min<int64_t>((s.rows + threads - 1) / threads, (int64_t)s.sm_count * 16);
  k_gen_outputs<<<grid_y, threads, 0, s.stream>>>(X, y, s.rows, s.f, family, prm, n, seed, noise_sigma, row_offset);
  return cudaGetLastError() == cudaSuccess ? 2 : SO_ECUDA;
}

}  // namespace so
