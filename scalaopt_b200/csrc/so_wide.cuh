// so_wide.cuh — shared pieces of the linear-predictor ("wide row") kernels: linear regression with or
// without intercept and logistic regression, f inputs per observation, row-major X[m x f] + y[m].
//
// Thread mapping ("L lanes per row"): a warp covers 32/L consecutive rows per step; lane (r, q) = (lane / L,
// lane % L) holds the 16-byte units c*L + q, c < C, of row r in registers (V = 2 doubles per unit; odd row lengths
// as aligned units, see ODD below).  A row's dot product is finished with log2(L) shuffle steps inside the row's L
// lanes, the per-row scalar work (residual / sigmoid / loss) is done redundantly by those L lanes, and every lane
// accumulates the gradient components of its own columns: the only reuse of a row is register-level (dot product,
// then rank-1 update with the same registers).  How the units get into the registers: from 4 lanes per row on
// through a lane-private cp.async ring in shared memory (below), with 1-2 lanes per row by 128-bit streaming
// loads with the next row group prefetched in registers.
#pragma once
#include "so_device.cuh"
#include "so_internal.h"

namespace so {
namespace wide {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

// parameter block passed by value in the kernel argument buffer (CUDA 12.1+: up to 32 KB of arguments)
struct Params {
  double p[SO_MAX_FEATURES + 1];
  double d[SO_MAX_FEATURES + 1];
  double alpha[SO_MAX_ALPHAS];
  int n, f, icpt, k;
  double scale0;    // factor applied to out[0] (loss sum) on the way out
};

enum { kLinear = 0, kLogistic = 1 };

// loss and dloss/dz at linear predictor z.
//   linear   : loss = (z - y)^2 (x 1/2 on the way out), w = z - y        MSEFunction.scala:45-48
//   logistic : o = 1/(1+exp(-z)); loss = CE(o, y); w = o - y             FFNeuralNetwork.scala:80-88,174-179
template <int FAM>
__device__ __forceinline__ void link(double z, double y, double& loss, double& w) {
  if (FAM == kLinear) {
    w = z - y;
    loss = w * w;
  } else {
    const double e = so_exp(-fabs(z));                  // in [0, 1] (NaN for a NaN predictor: propagates)
    const double inv = so_rcp_fast(1.0 + e);
    const double o = (z >= 0.0) ? inv : e * inv;
    w = o - y;
    // -y log(o/y) - (1-y) log((1-o)/(1-y)) = softplus(-z) + (1-y) z + [y log y + (1-y) log(1-y)]
    double l = ((e >= 0.0 && e <= 1.0) ? so_log1p_unit(e) : log1p(e)) + fmax(-z, 0.0) + (1.0 - y) * z;
    if (y > 0.0 && y < 1.0) l += y * log(y) + (1.0 - y) * log(1.0 - y);
    loss = l;
  }
}

// curvature weight d2loss/dz2: 1 (linear) or o (1 - o) (logistic, LogisticFunction.derivative :26)
template <int FAM>
__device__ __forceinline__ double curvature(double z) {
  if (FAM == kLinear) return 1.0;
  const double e = so_exp(-fabs(z));
  const double inv = so_rcp_fast(1.0 + e);
  return e * inv * inv;
}

template <int L>
__device__ __forceinline__ double row_sum(double v) {       // sum over the L lanes of a row
#pragma unroll
  for (int o = L / 2; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <int L>
__device__ __forceinline__ double col_sum(double v) {       // sum over the 32/L rows held by a warp
#pragma unroll
  for (int o = L; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// registers of one lane for one row step
template <int C, int V>
struct Frag { double x[C * V]; };

// POLICY 0: streaming loads (read once, evict first); 1: L1-allocating loads — for rows that are not sector-aligned (odd f)
// the 64-byte pieces of neighbouring lane groups share 32-byte sectors, which then come from L1 instead of L2 twice
#ifndef SO_WIDE_ODD_LD
#define SO_WIDE_ODD_LD 0
#endif
template <int L, int C, int V, int POLICY = 0>
__device__ __forceinline__ void load_frag(Frag<C, V>& fr, const double* __restrict__ X, int64_t row, bool valid,
                                          int f, int q) {
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const int col = (c * L + q) * V;
    if (V == 2) {
      double2 v = make_double2(0.0, 0.0);
      if (valid && col < f) {
        const double2* src = reinterpret_cast<const double2*>(X + row * f + col);
        v = POLICY == 1 ? __ldg(src) : (POLICY == 2 ? *src : ld_stream2(src));
      }
      fr.x[c * V] = v.x;
      fr.x[c * V + 1] = v.y;
    } else {
      fr.x[c * V] = (valid && col < f) ? ld_stream(X + row * f + col) : 0.0;
    }
  }
}

// ---- odd f (k_wide / k_wide_line with ODD = true, V = 2) --------------------------------------------------------------
// A row of an odd number of doubles starts 16-byte aligned only every other row, and with 64-bit loads (V = 1) the mapping
// loses a third of the bandwidth (cfg5 as specified has f = n - 1 = 31 and 127: 0.58 and 0.62 of HBM in round 1).  But
// every row is COVERED by (f + 1)/2 aligned 16-byte units: a row that starts aligned ends with one foreign element (the
// first of the next row), a row that does not starts with one (the last of the previous row).  So a lane streams its row
// as f + 1 elements from the aligned address `row f - odd`, and its weights are shifted by `odd` with a zero on the
// foreign element.  The parity of a lane's rows never changes — rows per warp step and the group stride are even (or,
// with one row per warp, the warp index decides) — so the trick costs a pointer offset: the loop is instruction for
// instruction the even-f one.  The foreign element is data of the same data set (finite whenever the result is), except
// before the FIRST and after the LAST row of the shard, where it is zeroed in registers (a view's neighbours are not its
// data; X carries 16 bytes of slack for the load behind the last row, so_api.cu).
template <int L>
__device__ __forceinline__ int lane_row_parity(const double* X, int warp, int r) {     // 0: this lane's rows start 16-byte aligned
  return (((32 / L == 1) ? warp : r) + (int)((reinterpret_cast<uintptr_t>(X) >> 3) & 1)) & 1;
}
// Xs = X - odd
template <int L, int C, int V, bool ODD>
__device__ __forceinline__ void load_row(Frag<C, V>& fr, const double* __restrict__ Xs, int64_t row, int64_t rows, bool valid,
                                         int f, int q, int odd) {
  load_frag<L, C, V, ODD ? SO_WIDE_ODD_LD : 0>(fr, Xs, row, valid, f, q);
  if (ODD && ((row == rows - 1 && !odd) || (row == 0 && odd))) {
#pragma unroll
    for (int c = 0; c < C; ++c) {
      if (!odd && (c * L + q) * V + 1 == f) fr.x[c * V + 1] = 0.0;
      if (odd && c == 0 && q == 0) fr.x[0] = 0.0;
    }
  }
}

// ---- lane-private shared-memory ring (prefetch mode 2) ------------------------------------------------------------------
// The register prefetch keeps ONE row group in flight per warp: 16 warps x 2 KB = 32 KB per SM, and at ~1700 cycles of
// DRAM latency that is 5.4 TB/s — exactly what the odd-f K1 ran at (ncu: long_scoreboard 9.2 of 13.3 stall cycles per
// issue, DRAM bytes = algorithmic; profiles/r2_k1_odd_units_ncu.json).  Here the NEXT kRingDepth groups travel with cp.async
// (LDGSTS) into slots that belong to the lane that will consume them: no registers are held for data in flight, no
// barrier (a lane reads back only what it copied itself, cp.async.wait_group orders that; y is the one shared value of a row),
// and kRingDepth x 32 KB per SM are in flight.  A slot is refilled after its values were consumed (they are operands of the
// step's FMAs), so the asynchronous write cannot overtake the read.
// Measured (profiles/r2_wide_ring_*.jsonl, one box each).  First version, ring base wherever the reduction scratch ended
// (16-byte aligned): wins with 16/32 lanes per row for K1 (f = 127: 3.83 -> 3.28 ms per 2e7 rows) but erratic (0.85-0.94 of
// HBM depending on the scratch size), losses for K4, the even-f K3 and everything with 4 lanes per row.  With the base
// 128-byte aligned (a warp's 32 x 16-byte LDGSTS covers four shared-memory lines instead of five) it wins nearly
// everywhere: K1 f = 32: 4.21 -> 3.76 ms per 1e8 rows (0.96 -> 1.07 of the HBM copy figure; the read-only stream ceiling is
// 1.09), f = 31: 4.73 -> 4.38 (0.89), f = 47: 0.67 -> 0.93, f = 64: 0.95 -> 1.09, f = 128 / 256: 0.94 -> 1.10; K3 alike; K4
// f = 47: 0.77 -> 0.97, f = 64: 1.01 -> 1.13, f = 128: 1.00 -> 1.10, f = 32: 1.02 -> 1.08.  The one loser left is K4 on odd f
// with 4 lanes per row (f = 31: 4.13 -> 4.30), which stays on the register prefetch, as do the 1- and 2-lane mappings.
// Also measured and not kept (profiles/r2_wide_long_pieces_ab.jsonl): ONE 16-byte unit per lane and as many lanes per row
// as the row has units, so that a misaligned row costs one extra sector instead of one per 64-byte piece, with ring depths
// 2, 4, 8 for the bytes in flight — f = 31: 4.74 -> 7.1 ms, f = 32: 4.21 -> 6.5: the 16-lane shuffle tree per two rows costs
// more than the sectors saved.
#ifndef SO_RING_DEPTH
#define SO_RING_DEPTH 2      // power of two; 2 beat 4 and 8 on every shape it helps (profiles/r2_wide_ring_ab.jsonl)
#endif
constexpr int kRingDepth = SO_RING_DEPTH;
__device__ __forceinline__ void cpa16(void* dst, const void* src, bool valid) {
  const unsigned int sz = valid ? 16u : 0u;      // src-size 0: zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(smem_u32(dst)), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cpa8(void* dst, const void* src, bool valid) {
  const unsigned int sz = valid ? 8u : 0u;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" :: "r"(smem_u32(dst)), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cpa_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cpa_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }
#ifndef SO_RING_ALIGN
#define SO_RING_ALIGN 128    // bytes; the ring follows the reduction scratch in dynamic shared memory.  A warp's 32 x 16-byte
                             // LDGSTS should cover four shared-memory lines, not five: with a 16-byte aligned base K1 ran
                             // at 0.85-0.94 of HBM depending on the scratch size, with 128 bytes at 1.03-1.10
#endif
__host__ __device__ constexpr int ring_offset_doubles(int scratch_doubles) {
  return (scratch_doubles + SO_RING_ALIGN / 8 - 1) / (SO_RING_ALIGN / 8) * (SO_RING_ALIGN / 8);
}
constexpr size_t ring_bytes(int C) { return (size_t)kRingDepth * kThreads * ((size_t)C * 16 + 8); }
// this lane's slots: x[(stage * C + c) * kThreads], y[stage * kThreads]
template <int C> struct LaneRing { double2* x; double* y; };
template <int C>
__device__ __forceinline__ LaneRing<C> lane_ring(void* base) {
  LaneRing<C> rg;
  rg.x = reinterpret_cast<double2*>(base) + threadIdx.x;
  rg.y = reinterpret_cast<double*>(reinterpret_cast<double2*>(base) + (size_t)kRingDepth * C * kThreads) + threadIdx.x;
  return rg;
}
// NEEDY: the row's y travels too — copied by the row's first lane only, read by all of its lanes after a __syncwarp
template <int L, int C, bool NEEDY>
__device__ __forceinline__ void ring_issue(const LaneRing<C>& rg, int stage, const double* __restrict__ Xs, const double* __restrict__ y,
                                           int64_t row, bool valid, int f, int q) {
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const int col = (c * L + q) * 2;
    const bool live = valid && col < f;
    cpa16(rg.x + (size_t)(stage * C + c) * kThreads, live ? Xs + row * f + col : Xs, live);
  }
  if (NEEDY) {
    if (L > 1) __syncwarp();                                     // every lane of the row has read the y this slot held
    if (q == 0) cpa8(rg.y + (size_t)stage * kThreads, valid ? y + row : y, valid);
  }
}
template <int L, int C, bool ODD, bool NEEDY>
__device__ __forceinline__ void ring_read(Frag<C, 2>& fr, double& yv, const LaneRing<C>& rg, int stage, int64_t row, int64_t rows,
                                          int f, int q, int odd) {
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const double2 v = rg.x[(size_t)(stage * C + c) * kThreads];
    fr.x[c * 2] = v.x; fr.x[c * 2 + 1] = v.y;
  }
  yv = 0.0;
  if (NEEDY) {
    if (L > 1) __syncwarp();                                     // the first lane's wait_group has passed as well
    yv = *(rg.y + (size_t)stage * kThreads - q);                 // slot of lane (r, 0)
  }
  if (ODD && ((row == rows - 1 && !odd) || (row == 0 && odd))) {
#pragma unroll
    for (int c = 0; c < C; ++c) {
      if (!odd && (c * L + q) * 2 + 1 == f) fr.x[c * 2 + 1] = 0.0;
      if (odd && c == 0 && q == 0) fr.x[0] = 0.0;
    }
  }
}

// After the row loop: block-level sums in warp order, per-block partials, last block adds in block order.
// vals[Q] in shared memory holds this block's sums; out[0 .. nscaled) is scaled by scale0.
__device__ __forceinline__ void finish_grid(const double* vals, int Q, int nscaled, double scale0,
                                            double* __restrict__ partials, double* __restrict__ out,
                                            unsigned int* ticket) {
  for (int q = threadIdx.x; q < Q; q += blockDim.x) partials[(size_t)blockIdx.x * Q + q] = vals[q];
  __threadfence();
  __syncthreads();
  __shared__ unsigned int s_last;
  if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
  __syncthreads();
  if (s_last) {
    __threadfence();
    const int G = partials_lanes(Q, blockDim.x);              // lanes per output, see grid_reduce in so_device.cuh
    const int per_round = blockDim.x / G, sub = threadIdx.x % G;
    for (int q0 = 0; q0 < Q; q0 += per_round) {
      const int q = q0 + threadIdx.x / G;
      double s = 0.0;
      if (q < Q) for (unsigned b = sub; b < gridDim.x; b += G) s += __ldcg(&partials[(size_t)b * Q + q]);
      for (int o = G >> 1; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (q < Q && sub == 0) out[q] = (q < nscaled) ? s * scale0 : s;
    }
    if (threadIdx.x == 0) *ticket = 0u;
    so_finish(ticket, out, Q);
  }
}

#ifndef SO_MAP_4X5
#define SO_MAP_4X5 1
#endif
// lane-mapping choice for f inputs
struct Mapping { int L, C, V; };
inline Mapping choose_mapping(int f) {
  Mapping m;
  m.V = (f % 2 == 0) ? 2 : 1;
  const int units = (f + m.V - 1) / m.V;      // V-wide column units per row
  m.C = 4;
  m.L = 1;
  while (m.L < 32 && units > m.L * 4) m.L *= 2;
  if (units > m.L * m.C) m.C = 8;
  if (units > m.L * m.C) m.C = 16;
  // 17-20 units (f = 33 .. 40): 4 lanes x 5 units instead of 8 lanes x 4 with half of them idle (cfg2's shape plus an
  // intercept column or a few features more: 0.51 of HBM with the 8-lane mapping)
  if (SO_MAP_4X5 && m.V == 2 && units > 16 && units <= 20) { m.L = 4; m.C = 5; }
  // the same gap one and two octaves up: 33-40 units (f = 65 .. 80) -> 8 x 5, 65-80 units (f = 129 .. 160) -> 16 x 5,
  // 129-160 units (f = 257 .. 320) -> 32 x 5 instead of 32 x 8
  if (SO_MAP_4X5 && m.V == 2 && units > 32 && units <= 40) { m.L = 8; m.C = 5; }
  if (SO_MAP_4X5 && m.V == 2 && units > 64 && units <= 80) { m.L = 16; m.C = 5; }
  if (SO_MAP_4X5 && m.V == 2 && units > 128 && units <= 160) { m.L = 32; m.C = 5; }
  return m;
}

}  // namespace wide
}  // namespace so
