// gram_ab.cu — A/B harness for the mid-size DMMA Gram kernel k_gram_tri (f = 32 or 16, linear / logistic).
// One executable per variant (-DSO_TRI_RING=<depth>, 0 = register prefetch); prints kernel ms and a checksum.
#define SO_GRAM_MINIMAL
#include "../scalaopt_b200/csrc/so_gram.cu"
#include <cstdio>
#include <vector>

__global__ void k_fill(double* X, double* y, int64_t rows, int f) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < rows * f; i += (int64_t)gridDim.x * blockDim.x) {
    uint64_t h = (uint64_t)i * 0x9E3779B97F4A7C15ull; h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
    X[i] = (double)(h >> 11) * (1.0 / 9007199254740992.0) - 0.5;
    if (i < rows) y[i] = (double)((h >> 5) & 1);
  }
}

int main(int argc, char** argv) {
  const int64_t rows = argc > 1 ? atoll(argv[1]) : 100000000ll;
  const char* name = argc > 2 ? argv[2] : "variant";
  const int f = argc > 3 ? atoi(argv[3]) : 32;
  const int fam = argc > 4 ? atoi(argv[4]) : 0;          // 0 linear (no intercept), 1 logistic (intercept)
  double *X, *y, *partials, *out; unsigned int* ticket;
  cudaMalloc(&X, rows * f * 8); cudaMalloc(&y, rows * 8);
  const size_t np = so::wide::gram_partials_needed(f + 1, f, 148) * 2;
  cudaMalloc(&partials, np * 8); cudaMalloc(&out, 4096 * 8); cudaMalloc(&ticket, 4); cudaMemset(ticket, 0, 4);
  k_fill<<<148 * 8, 256>>>(X, y, rows, f);
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("fill failed\n"); return 1; }
  so::Shard s; s.X = X; s.y = y; s.rows = rows; s.f = f; s.partials = partials; s.partials_doubles = np;
  s.out = out; s.out_doubles = 4096; s.ticket = ticket;
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0); s.sm_count = prop.multiProcessorCount;
  static so::wide::Params P;
  P.f = f; P.icpt = fam; P.n = f + fam; P.k = 0; P.scale0 = 1.0;
  for (int i = 0; i < P.n; ++i) P.p[i] = 0.1 * ((i % 7) - 3);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  std::vector<float> ms;
  for (int r = 0; r < 13; ++r) {
    cudaEventRecord(e0);
    const int rc = so::wide::launch_gram(s, fam, P);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    if (rc <= 0 || cudaGetLastError() != cudaSuccess) { printf("launch failed rc=%d %s\n", rc, cudaGetErrorString(cudaGetLastError())); return 1; }
    float m; cudaEventElapsedTime(&m, e0, e1); if (r >= 3) ms.push_back(m);
  }
  std::sort(ms.begin(), ms.end());
  const int Q = P.n * (P.n + 1) / 2 + P.n + 1;
  std::vector<double> h(Q); cudaMemcpy(h.data(), out, Q * 8, cudaMemcpyDeviceToHost);
  double cs = 0; for (int i = 0; i < Q; ++i) cs += h[i] * ((i % 13) + 1);
  printf("{\"variant\": \"%s\", \"rows\": %lld, \"f\": %d, \"fam\": %d, \"ms_best\": %.4f, \"ms_median\": %.4f, \"gbs_median\": %.1f, \"checksum\": %.17g, \"rtr\": %.17g}\n",
         name, (long long)rows, f, fam, ms.front(), ms[ms.size() / 2], rows * 8e-6 * (f + 1) / ms[ms.size() / 2], cs, h[Q - 1]);
  return 0;
}
