// so_gram.cu — K2 for the linear-predictor families: J^T J, J^T r and r^T r in one pass.
//
// For these families a Jacobian row is +-[1, x_i] (linear least squares: J = -sign(y - f) [1, x], r = |y - f|;
// logistic under Levenberg-Marquardt: the reference trainer's convention J = [1, x], r = o - y,
// Neuron.scala:67-73 / FFNeuralNetwork.scala:174-179).  With Z = [x | 1 | res] (the intercept column is moved
// behind the inputs so that the inputs stay contiguous) all three results are blocks of the Gram matrix Z^T Z:
// a SYRK over m rows.  This replaces MSEFunction.jacobianAndResidualsMatrix (MSEFunction.scala:132-136) plus the
// 2n+1 passes of QR.apply (linalg/QR.scala:131-250) that only ever needed these sums.
//
//   n <= 8      one thread per observation, packed upper triangle in registers (FP64 FMA pipe, HBM-bound)
//   n+1 <= 48   DMMA (mma.sync.m8n8k4.f64): a warp owns the whole upper triangle of 8x8 tiles and streams slabs
//               of 4 rows.  The A and B fragments of tile (ta, tb) are the SAME registers — lane l holds
//               Z[row0 + l%4][8t + l/4] — so a slab costs T 64-bit loads per lane, straight from global memory.
//               Measured on this pool's B200: DMMA 37.2 TFLOP/s vs 33.8 for register-resident DFMA
//               (profiles/r1_fp64_peak.json), and one DMMA replaces 8 DFMA issue slots + their operand traffic.
//   larger n    shared-memory staged DMMA SYRK over 4x4-tile patches (k_gram_big below), n <= 513
#include <algorithm>

#include "so_wide.cuh"

namespace so {
namespace wide {
namespace {

// ---------------------------------------------------------------------------------------------------
// n <= 8: thread per observation
// ---------------------------------------------------------------------------------------------------
template <int FAM, int NP, int ICPT>
__global__ void __launch_bounds__(kThreads, 2)
k_gram_small(const double* __restrict__ X, const double* __restrict__ y, int64_t rows, const __grid_constant__ Params P,
             double* __restrict__ partials, double* __restrict__ out, unsigned int* ticket) {
  constexpr int F = NP - ICPT, T = NP * (NP + 1) / 2, Q = T + NP + 1;
  __shared__ double s_red[kWarps * Q];
  __shared__ double s_scale[Q];
  if (FAM == kLogistic) exp_table_init();
  for (int q = threadIdx.x; q < Q; q += kThreads) s_scale[q] = (q >= T && q < T + NP && FAM == kLinear) ? -1.0 : 1.0;
  __syncthreads();
  double acc[Q];
#pragma unroll
  for (int q = 0; q < Q; ++q) acc[q] = 0.0;
  double p[NP];
#pragma unroll
  for (int i = 0; i < NP; ++i) p[i] = P.p[i];
  for (int64_t row = (int64_t)blockIdx.x * kThreads + threadIdx.x; row < rows; row += (int64_t)gridDim.x * kThreads) {
    double xt[NP];
    if (ICPT) xt[0] = 1.0;
#pragma unroll
    for (int j = 0; j < F; ++j) xt[ICPT + j] = ld_stream(X + row * F + j);
    const double yv = ld_stream(y + row);
    double z = 0.0;
#pragma unroll
    for (int i = 0; i < NP; ++i) z = fma(p[i], xt[i], z);
    double res;
    if (FAM == kLinear) {
      res = yv - z;                     // rho; J r = -x~ rho
    } else {
      const double e = so_exp(-fabs(z));
      const double inv = 1.0 / (1.0 + e);
      res = ((z >= 0.0) ? inv : e * inv) - yv;
    }
    int q = 0;
#pragma unroll
    for (int i = 0; i < NP; ++i)
#pragma unroll
      for (int j = i; j < NP; ++j) { acc[q] = fma(xt[i], xt[j], acc[q]); ++q; }
#pragma unroll
    for (int i = 0; i < NP; ++i) acc[T + i] = fma(xt[i], res, acc[T + i]);
    acc[T + NP] = fma(res, res, acc[T + NP]);
  }
  grid_reduce<Q>(acc, s_red, partials, out, s_scale, ticket);
}

template <int FAM, int NP, int ICPT>
int launch_small(const Shard& s, const Params& P) {
  auto kern = k_gram_small<FAM, NP, ICPT>;
  constexpr int Q = NP * (NP + 1) / 2 + NP + 1;
  int per_sm = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, 0);
  if (per_sm < 1) per_sm = 1;
  const int64_t want = std::max<int64_t>(1, (s.rows + kThreads - 1) / kThreads);
  const int grid = (int)std::min<int64_t>(want, (int64_t)per_sm * s.sm_count);
  if ((size_t)grid * Q > s.partials_doubles || (size_t)Q > s.out_doubles) return SO_ENOMEM;
  kern<<<grid, kThreads, 0, s.stream>>>(s.X, s.y, s.rows, P, s.partials, s.out, s.ticket);
  return cudaGetLastError() == cudaSuccess ? 1 : SO_ECUDA;
}

template <int FAM>
int dispatch_small(const Shard& s, const Params& P) {
#define SO_CASE(NPV) case NPV: return P.icpt ? launch_small<FAM, NPV, 1>(s, P) : launch_small<FAM, NPV, 0>(s, P);
  switch (P.n) {
    case 1: return P.icpt ? SO_EUNSUPPORTED : launch_small<FAM, 1, 0>(s, P);
    SO_CASE(2) SO_CASE(3) SO_CASE(4) SO_CASE(5) SO_CASE(6) SO_CASE(7) SO_CASE(8)
  }
#undef SO_CASE
  return SO_EUNSUPPORTED;
}

// ---------------------------------------------------------------------------------------------------
// 8 < n, n + 1 <= 8 T <= 48: DMMA over the upper triangle of 8x8 tiles
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int T> __host__ __device__ constexpr int tri_tiles() { return T * (T + 1) / 2; }

// EXTRA = true (f a multiple of 8): the tiles cover the f input columns only; the intercept's 1-column and the
// residual column are NOT padded into a sixth tile column — their sums (sum x_c, sum x_c res, sum res, sum res^2,
// row count) are accumulated with plain DADD/DFMA on the FP64 pipe, which is idle while the tensor pipe runs the
// DMMAs.  For f = 32 that is 10 DMMAs per 4-row slab instead of 15.
template <int FAM, int T, bool EXTRA>
__global__ void __launch_bounds__(kThreads, 2)
k_gram_tri(const double* __restrict__ X, const double* __restrict__ y, int64_t rows, const __grid_constant__ Params P,
           double* __restrict__ partials, double* __restrict__ out, unsigned int* ticket) {
  constexpr int NT = tri_tiles<T>();
  constexpr int QT = NT * 64 + (EXTRA ? 16 * T + 4 : 0);   // tiles, then [sum x_c] [sum x_c res] [cnt, sum res, sum res^2]
  extern __shared__ double smem[];            // QT doubles: this block's tiles (lane-major) and extras
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int rsel = lane & 3, csel = lane >> 2;
  const int f = P.f, icpt = P.icpt, n = P.n;
  const int col_one = (icpt && !EXTRA) ? f : -1, col_res = EXTRA ? -2 : f + icpt;
  if (FAM == kLogistic) { exp_table_init(); __syncthreads(); }

  double pz[T];
#pragma unroll
  for (int t = 0; t < T; ++t) pz[t] = (8 * t + csel < f) ? P.p[icpt + 8 * t + csel] : 0.0;
  const double p0 = icpt ? P.p[0] : 0.0;
  double acc[NT][2];
#pragma unroll
  for (int i = 0; i < NT; ++i) { acc[i][0] = 0.0; acc[i][1] = 0.0; }
  double s_one[EXTRA ? T : 1], s_res[EXTRA ? T : 1], cnt = 0.0, sr = 0.0, rr = 0.0;
#pragma unroll
  for (int t = 0; t < (EXTRA ? T : 1); ++t) { s_one[t] = 0.0; s_res[t] = 0.0; }

  const int64_t nslabs = (rows + 3) >> 2;
  const int64_t sstr = (int64_t)gridDim.x * kWarps;
  int64_t slab = (int64_t)blockIdx.x * kWarps + warp;

  double cur[T], nxt[T], ycur = 0.0, ynxt = 0.0;
  bool vcur = false, vnxt = false;
  auto load = [&](int64_t s, double (&fr)[T], double& yv, bool& valid) {
    const int64_t row = 4 * s + rsel;
    valid = row < rows;
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const int col = 8 * t + csel;
      fr[t] = (valid && col < f) ? ld_stream(X + row * f + col) : 0.0;
    }
    yv = valid ? ld_stream(y + row) : 0.0;
  };
  if (slab < nslabs) load(slab, cur, ycur, vcur);
  while (slab < nslabs) {
    const int64_t ns = slab + sstr;
    if (ns < nslabs) load(ns, nxt, ynxt, vnxt);
    if (EXTRA) {       // the tiles depend on the loads only: issue them first, the residual chain overlaps with them
#pragma unroll
      for (int ta = 0; ta < T; ++ta)
#pragma unroll
        for (int tb = ta; tb < T; ++tb) {
          const int idx = ta * T - ta * (ta - 1) / 2 + (tb - ta);
          dmma(acc[idx][0], acc[idx][1], cur[ta], cur[tb]);
        }
    }
    // linear predictor of the slab's 4 rows: lanes with equal lane%4 hold one row
    double dot = 0.0;
#pragma unroll
    for (int t = 0; t < T; ++t) dot = fma(pz[t], cur[t], dot);
    dot += __shfl_xor_sync(0xffffffffu, dot, 4);
    dot += __shfl_xor_sync(0xffffffffu, dot, 8);
    dot += __shfl_xor_sync(0xffffffffu, dot, 16);
    const double z = dot + p0;
    double res;
    if (FAM == kLinear) {
      res = ycur - z;
    } else {
      const double e = so_exp(-fabs(z));
      const double inv = 1.0 / (1.0 + e);
      res = ((z >= 0.0) ? inv : e * inv) - ycur;
    }
    if (!vcur) res = 0.0;
    if (EXTRA) {
#pragma unroll
      for (int t = 0; t < T; ++t) { s_one[t] += cur[t]; s_res[t] = fma(cur[t], res, s_res[t]); }
      if (csel == 0) { cnt += vcur ? 1.0 : 0.0; sr += res; rr = fma(res, res, rr); }
    } else {
      const double one = vcur ? 1.0 : 0.0;
#pragma unroll
      for (int t = 0; t < T; ++t) {
        const int col = 8 * t + csel;
        if (col == col_one) cur[t] = one;
        if (col == col_res) cur[t] = res;
      }
    }
    if (!EXTRA) {
#pragma unroll
      for (int ta = 0; ta < T; ++ta)
#pragma unroll
        for (int tb = ta; tb < T; ++tb) {
          const int idx = ta * T - ta * (ta - 1) / 2 + (tb - ta);
          dmma(acc[idx][0], acc[idx][1], cur[ta], cur[tb]);
        }
    }
#pragma unroll
    for (int t = 0; t < T; ++t) cur[t] = nxt[t];
    ycur = ynxt; vcur = vnxt;
    slab = ns;
  }

  if (EXTRA) {      // fold the 4 row slots of a column (lanes csel*4 + 0..3), and the per-row scalars over the warp
#pragma unroll
    for (int t = 0; t < T; ++t) {
      s_one[t] += __shfl_xor_sync(0xffffffffu, s_one[t], 1); s_one[t] += __shfl_xor_sync(0xffffffffu, s_one[t], 2);
      s_res[t] += __shfl_xor_sync(0xffffffffu, s_res[t], 1); s_res[t] += __shfl_xor_sync(0xffffffffu, s_res[t], 2);
    }
    cnt = warp_sum(cnt); sr = warp_sum(sr); rr = warp_sum(rr);
  }
  // block sum in warp order (deterministic), lane-major tile storage: tile[idx][lane*2 + reg]
  for (int w = 0; w < kWarps; ++w) {
    if (warp == w) {
#pragma unroll
      for (int i = 0; i < NT; ++i) {
        double* tile = smem + i * 64 + lane * 2;
        if (w == 0) { tile[0] = acc[i][0]; tile[1] = acc[i][1]; }
        else { tile[0] += acc[i][0]; tile[1] += acc[i][1]; }
      }
      if (EXTRA) {
        double* ex = smem + NT * 64;
        if (rsel == 0) {
#pragma unroll
          for (int t = 0; t < T; ++t) {
            if (w == 0) { ex[8 * t + csel] = s_one[t]; ex[8 * T + 8 * t + csel] = s_res[t]; }
            else { ex[8 * t + csel] += s_one[t]; ex[8 * T + 8 * t + csel] += s_res[t]; }
          }
        }
        if (lane == 0) {
          double* sc = ex + 16 * T;
          if (w == 0) { sc[0] = cnt; sc[1] = sr; sc[2] = rr; sc[3] = 0.0; }
          else { sc[0] += cnt; sc[1] += sr; sc[2] += rr; }
        }
      }
    }
    __syncthreads();
  }
  for (int q = threadIdx.x; q < QT; q += kThreads) partials[(size_t)blockIdx.x * QT + q] = smem[q];
  __threadfence();
  __syncthreads();
  __shared__ unsigned int s_last;
  if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
  __syncthreads();
  if (s_last) {
    __threadfence();
    // emit the API layout: packed upper triangle over parameters (a <= b), then Jtr, then rtr.
    // Z column of parameter a: (icpt ? (a == 0 ? ONE : a - 1) : a); ONE = the intercept's 1-column, RES = the residual.
    constexpr int ONE = -1, RES = -2;
    const int Tn = n * (n + 1) / 2, Qo = Tn + n + 1;
    const double jsgn = (FAM == kLinear) ? -1.0 : 1.0;         // J r = -x~ rho for least squares
    for (int q = threadIdx.x; q < Qo; q += kThreads) {
      int ca, cb;
      double sgn = 1.0;
      if (q < Tn) {
        int a = 0, rem = q;                     // unpack (a, b) from the row-major packed index
        while (rem >= n - a) { rem -= n - a; ++a; }
        const int b = a + rem;
        ca = icpt ? (a == 0 ? ONE : a - 1) : a;
        cb = icpt ? (b == 0 ? ONE : b - 1) : b;
      } else if (q < Tn + n) {
        const int a = q - Tn;
        ca = icpt ? (a == 0 ? ONE : a - 1) : a;
        cb = RES;
        sgn = jsgn;
      } else {
        ca = RES; cb = RES;
      }
      int off;
      if (EXTRA) {
        const int ex = NT * 64;
        if (ca >= 0 && cb >= 0) {
          if ((ca >> 3) > (cb >> 3)) { const int tmp = ca; ca = cb; cb = tmp; }
          const int ta = ca >> 3, tb = cb >> 3, i = ca & 7, j = cb & 7;
          off = (ta * T - ta * (ta - 1) / 2 + (tb - ta)) * 64 + (i * 4 + (j >> 1)) * 2 + (j & 1);
        } else if (ca == ONE && cb == ONE) off = ex + 16 * T;                 // row count
        else if (ca == RES && cb == RES) off = ex + 16 * T + 2;               // sum res^2
        else if ((ca == ONE && cb == RES) || (ca == RES && cb == ONE)) off = ex + 16 * T + 1;   // sum res
        else if (ca == ONE || cb == ONE) off = ex + (ca == ONE ? cb : ca);    // sum x_c
        else off = ex + 8 * T + (ca == RES ? cb : ca);                        // sum x_c res
      } else {
        if (ca == ONE) ca = f;
        if (cb == ONE) cb = f;
        if (ca == RES) ca = f + icpt;
        if (cb == RES) cb = f + icpt;
        if ((ca >> 3) > (cb >> 3)) { const int tmp = ca; ca = cb; cb = tmp; }
        const int ta = ca >> 3, tb = cb >> 3, i = ca & 7, j = cb & 7;
        off = (ta * T - ta * (ta - 1) / 2 + (tb - ta)) * 64 + (i * 4 + (j >> 1)) * 2 + (j & 1);
      }
      double s = 0.0;
      for (unsigned b2 = 0; b2 < gridDim.x; ++b2) s += __ldcg(&partials[(size_t)b2 * QT + off]);
      out[q] = s * sgn;
    }
    if (threadIdx.x == 0) *ticket = 0u;
  }
}

template <int FAM, int T, bool EXTRA>
int launch_tri(const Shard& s, const Params& P) {
  auto kern = k_gram_tri<FAM, T, EXTRA>;
  constexpr int QT = tri_tiles<T>() * 64 + (EXTRA ? 16 * T + 4 : 0);
  const size_t smem = sizeof(double) * QT;
  int per_sm = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, smem);
  if (per_sm < 1) per_sm = 1;
  const int64_t nslabs = (s.rows + 3) / 4;
  const int64_t want = std::max<int64_t>(1, (nslabs + kWarps - 1) / kWarps);
  const int grid = (int)std::min<int64_t>(want, (int64_t)per_sm * s.sm_count);
  const int Qo = P.n * (P.n + 1) / 2 + P.n + 1;
  if ((size_t)grid * QT > s.partials_doubles || (size_t)Qo > s.out_doubles) return SO_ENOMEM;
  kern<<<grid, kThreads, smem, s.stream>>>(s.X, s.y, s.rows, P, s.partials, s.out, s.ticket);
  return cudaGetLastError() == cudaSuccess ? 1 : SO_ECUDA;
}

template <int FAM>
int dispatch_tri(const Shard& s, const Params& P) {
  if (P.f % 8 == 0) {                      // tiles over the inputs only, extras on the FP64 pipe
    switch (P.f / 8) {
      case 1: return launch_tri<FAM, 1, true>(s, P);
      case 2: return launch_tri<FAM, 2, true>(s, P);
      case 3: return launch_tri<FAM, 3, true>(s, P);
      case 4: return launch_tri<FAM, 4, true>(s, P);
      case 5: return launch_tri<FAM, 5, true>(s, P);
    }
  }
  const int T = (P.n + 1 + 7) / 8;
  switch (T) {
    case 2: return launch_tri<FAM, 2, false>(s, P);
    case 3: return launch_tri<FAM, 3, false>(s, P);
    case 4: return launch_tri<FAM, 4, false>(s, P);
    case 5: return launch_tri<FAM, 5, false>(s, P);
    case 6: return launch_tri<FAM, 6, false>(s, P);
  }
  return SO_EUNSUPPORTED;
}

// ---------------------------------------------------------------------------------------------------
// n + 1 > 48 (up to 514 columns): shared-memory staged DMMA SYRK.
//   * the upper triangle of 8x8 tiles is cut into 4x4-tile patches (32x32 entries); a warp owns one patch
//     (16 DMMA accumulators = 64 registers), a 512-thread block owns up to 16 patches ("tile group");
//   * the grid is (tile groups) x (row workers); tile group is the fastest block index so that the blocks that
//     stream the same rows run together and share them through L2 (the kernel is FP64-bound by 2-6x: n = 256
//     needs 66 306 flop per 2 056-byte row);
//   * rows are staged 16 at a time into shared memory with cp.async (LDGSTS; 8-byte granules, so any f and any
//     alignment), double buffered, row pitch 8T+4 doubles so that the 4 rows x 8 columns of an MMA fragment hit
//     distinct banks; the intercept's 1-column and the residual column are appended in shared memory, after
//     which every fragment is one plain LDS.64;
//   * per 4-row slab a warp loads 4 + 4 fragments and issues 16 DMMAs (diagonal patches waste 6 of them).
// ---------------------------------------------------------------------------------------------------
constexpr int kBigThreads = 512, kBigWarps = 16, kPatch = 4, kBigStages = 3;

__device__ __forceinline__ void cp_async8(void* dst_smem, const void* src, bool valid) {
  const unsigned int sz = valid ? 8u : 0u;      // src-size 0: zero fill
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" :: "r"(smem_u32(dst_smem)), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

// RPW = rows per warp per chunk (chunk = 16 * RPW rows).  Warp w copies rows w, w+16, ... of a chunk itself and
// finishes their residual / intercept columns itself, so the only block barrier per chunk is the one that
// publishes the finished stage; with 3 stages that same barrier also protects the stage being refilled.
template <int FAM, int RPW>
__global__ void __launch_bounds__(kBigThreads, 1)
k_gram_big(const double* __restrict__ X, const double* __restrict__ y, int64_t rows, const __grid_constant__ Params P,
           int PT, int npatches, double* __restrict__ partials, double* __restrict__ out, unsigned int* ticket) {
  extern __shared__ __align__(16) double sm[];
  constexpr int kRows = kBigWarps * RPW;
  const int f = P.f, icpt = P.icpt, n = P.n;
  const int W = 8 * kPatch * PT;                       // padded width: whole patches, zero beyond f + icpt + 1
  const int PW = W + 4;                                // row pitch in doubles (4 rows x 8 cols of a fragment: distinct banks)
  double* ps = sm;                                     // [W] parameters in Z-column order (0 beyond f)
  double* stage0 = sm + W;                             // kBigStages stages of kRows x PW
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rsel = lane & 3, csel = lane >> 2;
  const int col_one = icpt ? f : -1, col_res = f + icpt;
  if (FAM == kLogistic) exp_table_init();
  for (int c = tid; c < W; c += kBigThreads) ps[c] = (c < f) ? P.p[icpt + c] : 0.0;
  for (int i = tid; i < kBigStages * kRows * PW; i += kBigThreads) stage0[i] = 0.0;     // padding columns stay zero
  __syncthreads();
  const double p0 = icpt ? P.p[0] : 0.0;

  // this warp's patch: the patch-th (pa <= pb) pair in row-major order over PT x PT
  const int patch = blockIdx.x * kBigWarps + warp;
  int pa = 0, pb = 0;
  const bool active = patch < npatches;
  if (active) { int rem = patch; while (rem >= PT - pa) { rem -= PT - pa; ++pa; } pb = pa + rem; }
  const int offa = 8 * kPatch * pa + csel, offb = 8 * kPatch * pb + csel;

  double acc[kPatch][kPatch][2];
#pragma unroll
  for (int i = 0; i < kPatch; ++i)
#pragma unroll
    for (int j = 0; j < kPatch; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

  const int64_t nchunks = (rows + kRows - 1) / kRows;
  auto load_stage = [&](int64_t chunk, int stage) {      // this warp's rows of the chunk; always commits a group
    if (chunk < nchunks) {
      double* st = stage0 + (size_t)stage * kRows * PW;
#pragma unroll
      for (int k = 0; k < RPW; ++k) {
        const int r = warp + k * kBigWarps;
        const int64_t row = chunk * kRows + r;
        const bool valid = row < rows;
        const double* src = X + (valid ? row : 0) * (int64_t)f;
        for (int c = lane; c < f; c += 32) cp_async8(st + r * PW + c, src + c, valid);
      }
    }
    cp_async_commit();
  };

  int64_t chunk = blockIdx.y;
  load_stage(chunk, 0);
  load_stage(chunk + gridDim.y, 1);
  int it = 0;
  for (; chunk < nchunks; chunk += gridDim.y, ++it) {
    const int stage = it % kBigStages;
    cp_async_wait<1>();                                  // this warp's copies of `chunk` have landed
    __syncwarp();
    double* st = stage0 + (size_t)stage * kRows * PW;
#pragma unroll
    for (int k = 0; k < RPW; ++k) {                      // residual and intercept columns of this warp's rows
      const int r = warp + k * kBigWarps;
      const int64_t row = chunk * kRows + r;
      const bool valid = row < rows;
      double dot = 0.0;
      for (int c = lane; c < f; c += 32) dot = fma(ps[c], st[r * PW + c], dot);
      dot = warp_sum(dot);
      if (lane == 0) {
        const double z = dot + p0;
        const double yv = valid ? ld_stream(y + row) : 0.0;
        double res;
        if (FAM == kLinear) {
          res = yv - z;
        } else {
          const double e = so_exp(-fabs(z));
          const double inv = 1.0 / (1.0 + e);
          res = ((z >= 0.0) ? inv : e * inv) - yv;
        }
        if (col_one >= 0) st[r * PW + col_one] = valid ? 1.0 : 0.0;
        st[r * PW + col_res] = valid ? res : 0.0;
      }
    }
    __syncthreads();     // stage complete for everyone; everyone has also finished computing on the previous chunk
    load_stage(chunk + 2 * (int64_t)gridDim.y, (it + 2) % kBigStages);
    if (active) {
#pragma unroll 2
      for (int slab = 0; slab < kRows / 4; ++slab) {
        const double* base = st + (slab * 4 + rsel) * PW;
        double fa[kPatch], fb[kPatch];
#pragma unroll
        for (int i = 0; i < kPatch; ++i) { fa[i] = base[offa + 8 * i]; fb[i] = base[offb + 8 * i]; }
        if (pa != pb) {
#pragma unroll
          for (int i = 0; i < kPatch; ++i)
#pragma unroll
            for (int j = 0; j < kPatch; ++j) dmma(acc[i][j][0], acc[i][j][1], fa[i], fb[j]);
        } else {            // diagonal patch: only tiles on or above the diagonal are ever read back
#pragma unroll
          for (int i = 0; i < kPatch; ++i)
#pragma unroll
            for (int j = i; j < kPatch; ++j) dmma(acc[i][j][0], acc[i][j][1], fa[i], fb[j]);
        }
      }
    }
  }
  cp_async_wait<0>();

  // per-(row worker, patch) partials, lane-major tiles: [worker][patch][i][j][lane*2 + reg]
  constexpr int kPatchDoubles = kPatch * kPatch * 64;
  if (active) {
    double* dst = partials + ((size_t)blockIdx.y * npatches + patch) * kPatchDoubles;
#pragma unroll
    for (int i = 0; i < kPatch; ++i)
#pragma unroll
      for (int j = 0; j < kPatch; ++j) {
        dst[(i * kPatch + j) * 64 + lane * 2] = acc[i][j][0];
        dst[(i * kPatch + j) * 64 + lane * 2 + 1] = acc[i][j][1];
      }
  }
  __threadfence();
  __syncthreads();
  __shared__ unsigned int s_last;
  const unsigned int nblocks = gridDim.x * gridDim.y;
  if (tid == 0) s_last = (atomicAdd(ticket, 1u) == nblocks - 1) ? 1u : 0u;
  __syncthreads();
  if (s_last) {
    __threadfence();
    const int Tn = n * (n + 1) / 2, Qo = Tn + n + 1;
    for (int q = tid; q < Qo; q += kBigThreads) {
      int ca, cb;
      double sgn = 1.0;
      if (q < Tn) {
        // row-major packed upper index -> (a, b): closed form, then fix up
        int a = (int)((2.0 * n + 1.0 - sqrt((2.0 * n + 1.0) * (2.0 * n + 1.0) - 8.0 * q)) * 0.5);
        while (a > 0 && a * n - a * (a - 1) / 2 > q) --a;
        while ((a + 1) * n - (a + 1) * a / 2 <= q) ++a;
        const int b = a + (q - (a * n - a * (a - 1) / 2));
        ca = icpt ? (a == 0 ? f : a - 1) : a;
        cb = icpt ? (b == 0 ? f : b - 1) : b;
      } else if (q < Tn + n) {
        const int a = q - Tn;
        ca = icpt ? (a == 0 ? f : a - 1) : a;
        cb = col_res;
        if (FAM == kLinear) sgn = -1.0;
      } else {
        ca = col_res; cb = col_res;
      }
      if ((ca >> 3) > (cb >> 3)) { const int tmp = ca; ca = cb; cb = tmp; }   // tile row <= tile col (hence patch row <= patch col)
      const int qa = ca / (8 * kPatch), qb = cb / (8 * kPatch);
      const int pidx = qa * PT - qa * (qa - 1) / 2 + (qb - qa);
      const int ti = (ca >> 3) - qa * kPatch, tj = (cb >> 3) - qb * kPatch, i = ca & 7, j = cb & 7;
      const size_t off = (size_t)pidx * kPatchDoubles + (ti * kPatch + tj) * 64 + (i * 4 + (j >> 1)) * 2 + (j & 1);
      double s2 = 0.0;
      for (unsigned w = 0; w < gridDim.y; ++w) s2 += __ldcg(&partials[(size_t)w * npatches * kPatchDoubles + off]);
      out[q] = s2 * sgn;
    }
    if (tid == 0) *ticket = 0u;
  }
}

struct BigPlan { int PT, npatches, groups, workers, rpw; size_t smem; };
inline BigPlan big_plan(int n, int sm_count) {
  BigPlan b;
  const int T = (n + 1 + 7) / 8;
  b.PT = (T + kPatch - 1) / kPatch;
  b.npatches = b.PT * (b.PT + 1) / 2;
  b.groups = (b.npatches + kBigWarps - 1) / kBigWarps;
  b.workers = std::max(1, sm_count / b.groups);
  const size_t W = (size_t)8 * kPatch * b.PT;
  auto bytes = [&](int rpw) { return sizeof(double) * (W + (size_t)kBigStages * kBigWarps * rpw * (W + 4)); };
  b.rpw = (bytes(2) <= 229000) ? 2 : 1;   // 227 KB per block minus ~3 KB of static shared memory
  b.smem = bytes(b.rpw);
  return b;
}

template <int FAM, int RPW>
int launch_big_rpw(const Shard& s, const Params& P, const BigPlan& b) {
  auto kern = k_gram_big<FAM, RPW>;
  static bool attr[64] = {false};
  int dev = 0; cudaGetDevice(&dev); dev &= 63;
  if (!attr[dev]) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 229000) != cudaSuccess) return SO_ECUDA;
    attr[dev] = true;
  }
  constexpr int kRows = kBigWarps * RPW;
  const int64_t nchunks = (s.rows + kRows - 1) / kRows;
  const int workers = (int)std::min<int64_t>(b.workers, std::max<int64_t>(1, nchunks));
  const size_t need = (size_t)workers * b.npatches * kPatch * kPatch * 64;
  const int Qo = P.n * (P.n + 1) / 2 + P.n + 1;
  if (need > s.partials_doubles || (size_t)Qo > s.out_doubles) return SO_ENOMEM;
  dim3 grid(b.groups, workers);
  kern<<<grid, kBigThreads, b.smem, s.stream>>>(s.X, s.y, s.rows, P, b.PT, b.npatches, s.partials, s.out, s.ticket);
  return cudaGetLastError() == cudaSuccess ? 1 : SO_ECUDA;
}

template <int FAM>
int launch_big(const Shard& s, const Params& P) {
  const BigPlan b = big_plan(P.n, s.sm_count);
  if (b.smem > 229000) return SO_EUNSUPPORTED;
  return b.rpw == 2 ? launch_big_rpw<FAM, 2>(s, P, b) : launch_big_rpw<FAM, 1>(s, P, b);
}

}  // namespace

size_t gram_partials_needed(int n, int f, int sm_count) {
  (void)f;
  if (n <= 8) return (size_t)sm_count * 8 * (size_t)(n * (n + 1) / 2 + n + 1);
  const int T = (n + 1 + 7) / 8;
  if (n + 1 <= 48) return (size_t)sm_count * 4 * ((size_t)(T * (T + 1) / 2) * 64 + 16 * T + 4);
  const BigPlan b = big_plan(n, sm_count);
  return (size_t)b.workers * b.npatches * kPatch * kPatch * 64;
}

int launch_gram(const Shard& s, int fam, const Params& P) {
  if (P.n <= 8) return fam == kLinear ? dispatch_small<kLinear>(s, P) : dispatch_small<kLogistic>(s, P);
  if (P.n + 1 <= 48) return fam == kLinear ? dispatch_tri<kLinear>(s, P) : dispatch_tri<kLogistic>(s, P);
  return fam == kLinear ? launch_big<kLinear>(s, P) : launch_big<kLogistic>(s, P);
}

}  // namespace wide
}  // namespace so
