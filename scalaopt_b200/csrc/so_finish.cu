// so_finish.cu — the cross-GPU sum + host delivery (so_finish, so_device.cuh) as a launch of its own, for the
// evaluation kernels whose result is emitted by all blocks rather than by one folding block (k_gram_ws).
#include "so_device.cuh"
#include "so_internal.h"

namespace so {
namespace {

__global__ void __launch_bounds__(1024) k_finish(unsigned int* ticket, const double* __restrict__ out, int q) {
  so_finish(ticket, out, q);
}

}  // namespace

int launch_finish(const Shard& s, int q) {
  k_finish<<<1, 1024, 0, s.stream>>>(s.ticket, s.out, q);
  return cudaGetLastError() == cudaSuccess ? 1 : SO_ECUDA;
}

}  // namespace so
