// k2_ab.cu — A/B harness for the headline kernel (K2, Gaussian peak + exponential background, n = 5).
// Compiles scalaopt_b200/csrc/so_scalar.cu with SO_SCALAR_MINIMAL (+ the variant's -D flags) into ONE executable per
// variant; tools/k2_ab.sh builds them and runs them back to back on the same box.  Prints kernel ms (CUDA events,
// best and median of 20 launches after 3 warm-ups) and a checksum of the result.
#define SO_SCALAR_MINIMAL
#include "../scalaopt_b200/csrc/so_scalar.cu"
#include <cstdio>
#include <vector>

__global__ void k_fill(double* t, double* y, int64_t rows) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < rows; i += (int64_t)gridDim.x * blockDim.x) {
    uint64_t h = (uint64_t)i * 0x9E3779B97F4A7C15ull; h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
    const double u = (double)(h >> 11) * (1.0 / 9007199254740992.0);
    const double tt = 4.0 * u, z = (tt - 2.0) / 0.5;
    t[i] = tt;
    y[i] = 2.0 * exp(-0.5 * z * z) + 1.0 * exp(-0.5 * tt) + 0.01 * ((double)((h >> 3) & 1023) / 1023.0 - 0.5);
  }
}

int main(int argc, char** argv) {
  const int64_t rows = argc > 1 ? atoll(argv[1]) : 1000000000ll;
  const char* name = argc > 2 ? argv[2] : "variant";
  const int kind = argc > 3 ? atoi(argv[3]) : 2;          // 1 = K1 loss+gradient, 2 = K2 normal equations, 4 = K4 Hessian.vector
  double *t, *y, *partials, *out; unsigned int* ticket;
  cudaMalloc(&t, rows * 8); cudaMalloc(&y, rows * 8);
  cudaMalloc(&partials, 148 * 8 * 64 * 8); cudaMalloc(&out, 64 * 8); cudaMalloc(&ticket, 4); cudaMemset(ticket, 0, 4);
  k_fill<<<148 * 8, 256>>>(t, y, rows);
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("fill failed\n"); return 1; }
  so::Shard s; s.X = t; s.y = y; s.rows = rows; s.f = 1; s.partials = partials; s.partials_doubles = 148 * 8 * 64;
  s.out = out; s.out_doubles = 64; s.ticket = ticket;
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0); s.sm_count = prop.multiProcessorCount;
  if (argc > 4 && atoi(argv[4])) { s.has_range = true; s.tmin = 0.0; s.tmax = 4.0; }     // inputs of k_fill: check-free K2
  const double p[5] = {1.9, 2.05, 0.52, 1.1, 0.45};
  const double d[5] = {0.03, -0.02, 0.01, 0.025, -0.015};
  const double alphas[8] = {0.5, 1.0, 0.25, 2.0, 0.125, 1.5, 0.75, 3.0};
  const int nalpha = argc > 6 ? atoi(argv[6]) : 1;         // kind 3 (K3 line search): step sizes per pass
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  std::vector<float> ms;
  const int launches = argc > 5 ? atoi(argv[5]) : 23;      // > 23: sustained (power-capped) behaviour, mean of the last 2/3
  for (int r = 0; r < launches; ++r) {
    cudaEventRecord(e0);
    const double pn[4] = {0.3, 2.0, 0.5, 0.5}, cn[1] = {0.0};
    const int rc = kind == 5 ? so::launch_nll(s, SO_PDF_GAUSS_EXPBKG, pn, 4, cn, 1) : so::launch_scalar((so::Kind)kind, s, SO_FAMILY_GAUSS_EXPBKG, p, d, 5, alphas, kind == 3 ? nalpha : 0, 1);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    if (rc <= 0 || cudaGetLastError() != cudaSuccess) { printf("launch failed rc=%d\n", rc); return 1; }
    float m; cudaEventElapsedTime(&m, e0, e1); if (r >= (launches > 23 ? launches / 3 : 3)) ms.push_back(m);
  }
  double mean = 0; for (float m : ms) mean += m; mean /= ms.size();
  std::sort(ms.begin(), ms.end());
  const int Q = kind == 2 ? 21 : (kind == 1 ? 6 : (kind == 3 ? 2 * so::scalar_line_block(nalpha) : (kind == 5 ? 1 : (kind == 6 ? 36 : 5))));
  double h[64] = {0}; cudaMemcpy(h, out, sizeof(double) * Q, cudaMemcpyDeviceToHost);
  double cs = 0; for (int i = 0; i < Q; ++i) cs += h[i] * (i + 1);
  printf("{\"variant\": \"%s\", \"rows\": %lld, \"ms_best\": %.4f, \"ms_median\": %.4f, \"ms_mean\": %.4f, \"gbs_median\": %.1f, \"checksum\": %.17g, \"last\": %.17g}\n",
         name, (long long)rows, ms.front(), ms[ms.size() / 2], mean, rows * 16e-6 / ms[ms.size() / 2], cs, h[Q - 1]);
  return 0;
}
