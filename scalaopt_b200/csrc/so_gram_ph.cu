// so_gram_ph.cu — K2 (J^T J, J^T r, r^T r) for the linear-predictor families with 8 < n and f <= 104 inputs:
// shared-memory staged DMMA SYRK in BLOCK-WIDE PHASES.
//
// Why another Gram kernel.  Round 1's k_gram_tri (a warp streams 4-row slabs from global memory into DMMA fragments
// and runs the residual chain — dot product, three shuffle round trips, link function — in the same warp) sat at
// ~300 cycles per slab and sub-partition whatever the prefetch depth or instruction count, against 205 cycles of FP64
// pipe time (profiles/r1_gram_tri_*.json, DESIGN.md "Mid-size Gram"): with warps drifting out of step, the dependent
// FP64 instructions of a warp in its chain queue behind the DMMA bursts of the other three warps of its sub-partition
// (DFMA and DMMA share one pipe), the chain takes several hundred cycles, and then all four warps end up in their
// chains together with the pipe idle.  Here the two kinds of work never meet:
//   phase A  all 16 warps compute the residuals of a chunk of R rows (R = 256 or 128): TPR threads per row read the row
//            with LDS.128, 4T/TPR fused multiply-adds each, one or two shuffles, the link function, res -> shared
//            memory.  No DMMA is in flight, so the chain runs at DFMA latency, R rows side by side.
//   phase B  all 16 warps run DMMAs (mma.sync.m8n8k4.f64) on that chunk: the fragments of a 4-row slab are T LDS.64
//            per lane (A and B fragments of a tile pair are the same registers), the accumulations sum x_c and
//            sum x_c res are independent DFMAs on the same fragments.  No dependent FP64 chain in this phase.
// Rows travel global -> shared memory into a two-stage ring; chunk c+1 is in flight during both phases of chunk c and no
// thread ever waits for a load it issued.  Even f, aligned base: ONE cp.async.bulk.tensor.2d (TMA, SASS UTMALDG) per 16
// columns of the chunk, 128-byte swizzle (layout kTma below).  Otherwise cp.async (LDGSTS) into rows of pitch 8T+4 doubles.
// Both layouts make the fragment reads of phase B and the LDS.128 of phase A bank-conflict free.
// The upper triangle of 8x8 tiles is dealt round-robin to SPLIT warps (T <= 4: one warp holds all 10 tiles; T = 13: 91
// tiles over 8 warps), the other 16/SPLIT-way parallelism is over slabs.  As in k_gram_tri's EXTRA mode the tiles cover
// the f INPUT columns only; the intercept's 1-column and the residual column are sums on the FP64 pipe.
//
// Replaces, per row, MSEFunction.jacobianAndResidualsMatrix (MSEFunction.scala:132-136) + the 2n+1 passes of QR.apply
// (linalg/QR.scala:131-250); conventions as in so_gram.cu.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <type_traits>

#include <cuda.h>      // CUtensorMap (types only; the encoder is fetched with cudaGetDriverEntryPoint)

#include "so_wide.cuh"

namespace so {
namespace wide {
namespace {


__device__ __forceinline__ void ph_dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
[[maybe_unused]] __device__ __forceinline__ void ph_cp16(void* dst, const void* src, bool valid) {
  const unsigned int sz = valid ? 16u : 0u;      // src-size 0: zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(smem_u32(dst)), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void ph_cp8(void* dst, const void* src, bool valid) {
  const unsigned int sz = valid ? 8u : 0u;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" :: "r"(smem_u32(dst)), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void ph_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void ph_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Phase B is written in program order on purpose (volatile asm is not reordered against volatile asm): the loads of the
// NEXT slab, then this slab's independent sums, then its DMMAs.  Left to the scheduler, the second slab of an unrolled
// pair got its LDS six instructions before their first use and the sum x_c res chain landed in the middle of the DMMA
// burst: two instructions held a quarter of all stall samples (profiles/r2_gram_ph_v2_source.txt).
__device__ __forceinline__ double ph_lds(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void ph_fma(double& d, double a, double b) { asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(d) : "d"(a), "d"(b)); }
__device__ __forceinline__ void ph_add(double& d, double a) { asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d) : "d"(a)); }

__host__ __device__ constexpr int ph_tiles(int T) { return T * (T + 1) / 2; }
__host__ __device__ constexpr int ph_split(int T) { return T <= 4 ? 1 : T <= 6 ? 2 : T <= 8 ? 4 : 8; }
// rows per chunk: as many as two stages of shared memory hold (kTma stages have no padding but whole 16-column boxes)
__host__ __device__ constexpr int ph_rows(int T, int layout) { return layout == 1 ? (((T + 1) / 2) * 16 * 8 * 256 * 2 <= 200 * 1024 ? 256 : 128) : (T <= 6 ? 256 : 128); }
__host__ __device__ constexpr int ph_qt(int T) { return ph_tiles(T) * 64 + 16 * T + 4; }   // tiles | sum x_c | sum x_c res | cnt, sum res, sum res^2

// Block sums in a fixed order (deterministic: slab groups one after the other into shared memory), per-block partials, and
// the block that draws the last ticket emits the API layout and runs so_finish.  Called by all threads after the row loop.
template <int FAM, int T, int NTH>
__device__ __forceinline__ void ph_epilogue(double (&acc)[(ph_tiles(T) + ph_split(T) - 1) / ph_split(T)][2],
                                            double (&a_one)[(T + ph_split(T) - 1) / ph_split(T)],
                                            double (&a_res)[(T + ph_split(T) - 1) / ph_split(T)], double cnt, double sr, double rr,
                                            double* red, const Params& P, double* __restrict__ partials, double* __restrict__ out,
                                            unsigned int* ticket) {
  constexpr int kPhThreads = NTH, kPhWarps = NTH / 32;
  constexpr int SPLIT = ph_split(T), NG = kPhWarps / SPLIT;
  constexpr int NT = ph_tiles(T), NTS = (NT + SPLIT - 1) / SPLIT, TS = (T + SPLIT - 1) / SPLIT, QT = ph_qt(T);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rsel = lane & 3, csel = lane >> 2;
  const int sub = warp % SPLIT, grp = warp / SPLIT;
  const int icpt = P.icpt, n = P.n;
  for (int i = tid; i < QT; i += kPhThreads) red[i] = 0.0;
  __syncthreads();
  // fold the 4 row slots of a column (lanes csel*4 + 0..3)
#pragma unroll
  for (int i = 0; i < TS; ++i) {
    a_one[i] += __shfl_xor_sync(0xffffffffu, a_one[i], 1); a_one[i] += __shfl_xor_sync(0xffffffffu, a_one[i], 2);
    a_res[i] += __shfl_xor_sync(0xffffffffu, a_res[i], 1); a_res[i] += __shfl_xor_sync(0xffffffffu, a_res[i], 2);
  }
  for (int g = 0; g < NG; ++g) {
    if (grp == g) {
#pragma unroll
      for (int i = 0; i < NTS; ++i) {
        const int idx = i * SPLIT + sub;
        if (idx < NT) { double* tile = red + idx * 64 + lane * 2; tile[0] += acc[i][0]; tile[1] += acc[i][1]; }
      }
      if (rsel == 0) {
#pragma unroll
        for (int i = 0; i < TS; ++i) {
          const int t = i * SPLIT + sub;
          if (t < T) { red[NT * 64 + 8 * t + csel] += a_one[i]; red[NT * 64 + 8 * T + 8 * t + csel] += a_res[i]; }
        }
      }
    }
    __syncthreads();
  }
  cnt = warp_sum(cnt); sr = warp_sum(sr); rr = warp_sum(rr);
  for (int w = 0; w < kPhWarps; ++w) {
    if (warp == w && lane == 0) { double* sc = red + NT * 64 + 16 * T; sc[0] += cnt; sc[1] += sr; sc[2] += rr; }
    __syncthreads();
  }
  for (int q = tid; q < QT; q += kPhThreads) partials[(size_t)blockIdx.x * QT + q] = red[q];
  __threadfence();
  __syncthreads();
  __shared__ unsigned int s_last_ph;
  if (tid == 0) s_last_ph = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
  __syncthreads();
  if (s_last_ph) {
    __threadfence();
    // emit the API layout: packed upper triangle over parameters (a <= b), then Jtr, then rtr.
    // Z column of parameter a: (icpt ? (a == 0 ? ONE : a - 1) : a); ONE = the intercept's 1-column, RES = the residual.
    constexpr int ONE = -1, RES = -2;
    const int Tn = n * (n + 1) / 2, Qo = Tn + n + 1;
    const double jsgn = (FAM == kLinear) ? -1.0 : 1.0;         // J r = -x~ rho for least squares
    for (int q = tid; q < Qo; q += kPhThreads) {
      int ca, cb;
      double sgn = 1.0;
      if (q < Tn) {
        int a = (int)((2.0 * n + 1.0 - sqrt((2.0 * n + 1.0) * (2.0 * n + 1.0) - 8.0 * q)) * 0.5);   // packed index -> (a, b)
        while (a > 0 && a * n - a * (a - 1) / 2 > q) --a;
        while ((a + 1) * n - (a + 1) * a / 2 <= q) ++a;
        const int b = a + (q - (a * n - a * (a - 1) / 2));
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
      double s = 0.0;
      for (unsigned b2 = 0; b2 < gridDim.x; ++b2) s += __ldcg(&partials[(size_t)b2 * QT + off]);
      out[q] = s * sgn;
    }
    if (tid == 0) *ticket = 0u;
    so_finish(ticket, out, Qo);
  }
}

// Shared-memory layouts of a stage.
//   kPadded  R rows at a pitch of 8T+4 doubles, filled by cp.async (LDGSTS, 16- or 8-byte requests): any f, any alignment.
//            A slab is 4 consecutive rows.
//   kTma     ceil(f/16) boxes of R rows x 16 columns (128 bytes per row), each ONE cp.async.bulk.tensor.2d request with the
//            128-byte swizzle: the 16-byte chunk j of row r sits at chunk j ^ (r % 8).  A slab is rows {0,2,4,6} or
//            {1,3,5,7} of a group of 8, which makes the fragment reads bank-conflict free; phase A puts 8 consecutive rows
//            on 8 consecutive lanes, which makes its LDS.128 conflict free.  Columns past f and rows past the end of the
//            data arrive as zeros (out-of-bounds fill) — no tail code.  Needs even f, a 16-byte aligned X and < 2^31 rows.
// Per-row TMA bulk copies into the padded layout were tried too: 64-74 cycles per request and SM, 25.6 ms per 1e8 rows of
// 32 doubles (profiles/r2_gram_ph_v4.jsonl) against 8.3 ms with LDGSTS.
enum { kPadded = 0, kTma = 1 };

template <int FAM, int T, int LAYOUT, int NTH>
__global__ void __launch_bounds__(NTH, 512 / NTH)
k_gram_ph(const double* __restrict__ X, const double* __restrict__ y, int64_t rows, const __grid_constant__ Params P,
          const __grid_constant__ CUtensorMap tmap, int wide_copies, double* __restrict__ partials, double* __restrict__ out,
          unsigned int* ticket) {
  constexpr int kPhThreads = NTH, kPhWarps = NTH / 32;
  constexpr int SPLIT = ph_split(T), NG = kPhWarps / SPLIT, R = ph_rows(T, LAYOUT) * NTH / 512, TPR = kPhThreads / R;
  constexpr int NT = ph_tiles(T), NTS = (NT + SPLIT - 1) / SPLIT, TS = (T + SPLIT - 1) / SPLIT;
  constexpr int PITCH = 8 * T + 4, QT = ph_qt(T), NB = (T + 1) / 2;
  constexpr int STAGE = LAYOUT == kTma ? NB * R * 16 : R * PITCH;      // doubles per stage
  constexpr int UPT = 4 * T / TPR;                 // 16-byte units of a row per phase-A thread
  static_assert((4 * T) % TPR == 0, "row units must split evenly");
  extern __shared__ __align__(1024) double ph_smem[];
  double* s_x = ph_smem;                           // 2 stages (first: the TMA boxes want 1024-byte alignment)
  double* s_y = s_x + 2 * STAGE;                   // 2 stages x R
  double* s_res = s_y + 2 * R;                     // R residuals of the chunk in phase B
  double* s_p = s_res + R;                         // 8T parameters of the input columns, zero padded
  __shared__ __align__(8) uint64_t s_bar[2];       // one mbarrier per stage (kTma)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rsel = lane & 3, csel = lane >> 2;
  const int f = P.f, icpt = P.icpt, n = P.n;
  if (FAM == kLogistic) exp_table_init();
  for (int i = tid; i < 8 * T; i += kPhThreads) s_p[i] = i < f ? P.p[icpt + i] : 0.0;
  if (LAYOUT == kPadded) for (int i = tid; i < 2 * STAGE; i += kPhThreads) s_x[i] = 0.0;     // pad columns stay zero for the whole launch
  const double p0 = icpt ? P.p[0] : 0.0;
  if (LAYOUT == kTma && tid == 0) { mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1); mbar_fence_init(); }
  __syncthreads();

  const int64_t nchunks = (rows + R - 1) / R;
  // kPadded: the chunk is one contiguous span of global memory; piece i = tid + 512 k of it goes to (row, column piece) =
  // (i / ppr, i % ppr) of the stage.  Quotient and remainder advance by constants from one k to the next, so the loop is a
  // copy, three adds and a wrap test per piece (the first version recomputed 64-bit addresses and a zero-fill predicate per
  // piece: 88 of its 159 instructions per slab).  Rows past the end of the data are zeroed with plain stores.
  const int ppr = wide_copies ? f >> 1 : f;         // pieces per row (16 or 8 bytes each)
  const int pbytes = wide_copies ? 16 : 8;
  const int pr0 = tid / ppr, pc0 = tid - pr0 * ppr, pdr = kPhThreads / ppr, pdc = kPhThreads - pdr * ppr;
  const int dst0 = (pr0 * PITCH) * 8 + pc0 * pbytes, ddst = (pdr * PITCH) * 8 + pdc * pbytes, dwrap = (PITCH * 8) - ppr * pbytes;
  const int total = R * ppr;
  auto issue = [&](int64_t chunk, int st) {
    const int64_t row0 = chunk * R;
    const int64_t left = rows - row0;               // valid rows of this chunk (may exceed R)
    if constexpr (LAYOUT == kTma) {
      if (tid == 0) {
        const bool ybulk = left >= R;
        mbar_arrive_expect_tx(&s_bar[st], (uint32_t)(NB * R * 128) + (ybulk ? (uint32_t)R * 8u : 0u));
#pragma unroll
        for (int b = 0; b < NB; ++b)
          asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                       :: "r"(smem_u32(s_x + (size_t)st * STAGE + (size_t)b * R * 16)), "l"(&tmap), "r"(16 * b), "r"((int)row0),
                          "r"(smem_u32(&s_bar[st])) : "memory");
        if (ybulk) bulk_g2s(s_y + st * R, y + row0, (uint32_t)R * 8u, &s_bar[st]);
      }
      if (left < R && tid < R) s_y[st * R + tid] = tid < left ? y[row0 + tid] : 0.0;
    } else {
      const int limit = left >= R ? total : (int)left * ppr;
      const uint32_t xs_u = smem_u32(s_x + (size_t)st * STAGE);
      const char* src = reinterpret_cast<const char*>(X + row0 * f) + (size_t)tid * pbytes;
      uint32_t dst = xs_u + dst0;
      int c = pc0;
      if (wide_copies) {
        for (int i = tid; i < limit; i += kPhThreads) {
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(src) : "memory");
          src += kPhThreads * 16; dst += ddst; c += pdc;
          if (c >= ppr) { c -= ppr; dst += dwrap; }
        }
      } else {
        for (int i = tid; i < limit; i += kPhThreads) {
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(dst), "l"(src) : "memory");
          src += kPhThreads * 8; dst += ddst; c += pdc;
          if (c >= ppr) { c -= ppr; dst += dwrap; }
        }
      }
      if (left < R) {                                   // ragged last chunk: rows [left, R) contribute nothing
        double* xs = s_x + (size_t)st * STAGE;
        for (int i = (int)left * PITCH + tid; i < R * PITCH; i += kPhThreads) xs[i] = 0.0;
      }
      if (tid < R) ph_cp8(s_y + st * R + tid, tid < left ? y + row0 + tid : y, tid < left);
    }
  };

  double acc[NTS][2];
#pragma unroll
  for (int i = 0; i < NTS; ++i) { acc[i][0] = 0.0; acc[i][1] = 0.0; }
  double a_one[TS], a_res[TS];
#pragma unroll
  for (int i = 0; i < TS; ++i) { a_one[i] = 0.0; a_res[i] = 0.0; }
  double cnt = 0.0, sr = 0.0, rr = 0.0;
  const int sub = warp % SPLIT, grp = warp / SPLIT;

  // phase A: which row and which 16-byte units of it this thread owns.  kPadded: TPR neighbouring lanes share a row, units
  // interleaved over them.  kTma: 32/TPR consecutive rows on consecutive lanes, the row's TPR threads TPR-th of a warp apart.
  const int a_row = LAYOUT == kTma ? (warp * (32 / TPR) + (lane & (32 / TPR - 1))) : tid / TPR;
  const int a_h = LAYOUT == kTma ? lane / (32 / TPR) : tid % TPR;
  // the parameters a thread multiplies with in phase A are the same for every chunk: registers when there are few
  constexpr bool kPregs = UPT <= 8;
  double2 preg[kPregs ? UPT : 1];
  if constexpr (kPregs) {
#pragma unroll
    for (int j = 0; j < UPT; ++j) preg[j] = *reinterpret_cast<const double2*>(s_p + 2 * (j * TPR + a_h));
  }
  // phase B, kTma: byte offset of this lane's fragment inside a row line, for even and odd tile columns (t % 2), per slab parity
  uint32_t sw[2][2];
#pragma unroll
  for (int par = 0; par < 2; ++par)
#pragma unroll
    for (int th = 0; th < 2; ++th) sw[par][th] = (uint32_t)((((4 * th + (csel >> 1)) ^ (2 * rsel + par)) << 4) + ((csel & 1) << 3));

  int64_t chunk = blockIdx.x;
  int st = 0;
  if (chunk < nchunks) issue(chunk, 0);
  ph_commit();
  uint32_t par0 = 0, par1 = 0;                     // phase parity of the two stage barriers
  for (; chunk < nchunks; chunk += gridDim.x, st ^= 1) {
    if constexpr (LAYOUT == kTma) {
      if (st == 0) { mbar_wait(&s_bar[0], par0); par0 ^= 1; } else { mbar_wait(&s_bar[1], par1); par1 ^= 1; }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // our reads of the other stage precede its refill
    } else {
      ph_wait_all();
    }
    __syncthreads();          // chunk `chunk` has landed for everybody, and everybody is done with the previous chunk
    if (chunk + gridDim.x < nchunks) issue(chunk + gridDim.x, st ^ 1);
    ph_commit();
    const double* xs = s_x + (size_t)st * STAGE;
    // ---- phase A: residuals of the chunk's R rows ----
    {
      const int r = a_row, h = a_h;
      double d0 = 0.0, d1 = 0.0, d2 = 0.0, d3 = 0.0;       // four chains: the phase is latency-bound, not pipe-bound
#pragma unroll
      for (int j = 0; j < UPT; ++j) {
        const int u = j * TPR + h;
        double2 xv;
        if constexpr (LAYOUT == kTma) xv = *reinterpret_cast<const double2*>(xs + (u >> 3) * (R * 16) + r * 16 + (((u & 7) ^ (r & 7)) << 1));
        else xv = *reinterpret_cast<const double2*>(xs + r * PITCH + 2 * u);
        double2 pv;
        if constexpr (kPregs) pv = preg[j]; else pv = *reinterpret_cast<const double2*>(s_p + 2 * u);
        if (j & 1) { d2 = fma(xv.x, pv.x, d2); d3 = fma(xv.y, pv.y, d3); }
        else { d0 = fma(xv.x, pv.x, d0); d1 = fma(xv.y, pv.y, d1); }
      }
      double dot = (d0 + d1) + (d2 + d3);
      if constexpr (LAYOUT == kTma) {
#pragma unroll
        for (int o = 32 / TPR; o < 32; o <<= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
      } else {
#pragma unroll
        for (int o = 1; o < TPR; o <<= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
      }
      const double z = dot + p0;
      const double yv = s_y[st * R + r];
      const bool valid = chunk * R + r < rows;
      double res;
      if (FAM == kLinear) {
        res = yv - z;                       // rho; J r = -x~ rho
      } else {
        const double e = so_exp(-fabs(z));
        const double inv = so_rcp_fast(1.0 + e);
        res = ((z >= 0.0) ? inv : e * inv) - yv;
      }
      if (!valid) res = 0.0;
      if (h == 0) {
        s_res[r] = res;
        cnt += valid ? 1.0 : 0.0;
        sr += res;
        rr = fma(res, res, rr);
      }
    }
    __syncthreads();
    // ---- phase B: DMMAs; warp (sub, grp) takes tiles idx % SPLIT == sub of slabs grp, grp + NG, ... ----
    auto phase_b = [&](auto sub_c) {
      constexpr int SUB = decltype(sub_c)::value;
      constexpr int SL = R / 4 / NG;                 // slabs of a chunk per warp
      // byte address of tile column t of this lane's fragment of the warp's s2-th slab, and of that slab row's residual
      auto frag = [&](int s2, int t) -> uint32_t {
        const int sl = grp + s2 * NG;
        if constexpr (LAYOUT == kTma)
          return smem_u32(xs) + (uint32_t)((t >> 1) * (R * 128) + ((sl >> 1) * 8 + 2 * rsel + (sl & 1)) * 128) + sw[sl & 1][t & 1];
        else
          return smem_u32(xs) + (uint32_t)(((4 * sl + rsel) * PITCH + csel + 8 * t) * 8);
      };
      auto resa = [&](int s2) -> uint32_t {
        const int sl = grp + s2 * NG;
        if constexpr (LAYOUT == kTma) return smem_u32(s_res) + (uint32_t)(((sl >> 1) * 8 + 2 * rsel + (sl & 1)) * 8);
        else return smem_u32(s_res) + (uint32_t)((4 * sl + rsel) * 8);
      };
      double fr[2][T], rv[2];
#pragma unroll
      for (int t = 0; t < T; ++t) fr[0][t] = ph_lds(frag(0, t));
      rv[0] = ph_lds(resa(0));
#pragma unroll
      for (int s2 = 0; s2 < SL; ++s2) {
        const int cu = s2 & 1, nx = cu ^ 1;
        if (s2 + 1 < SL) {
#pragma unroll
          for (int t = 0; t < T; ++t) fr[nx][t] = ph_lds(frag(s2 + 1, t));
          rv[nx] = ph_lds(resa(s2 + 1));
        }
#pragma unroll
        for (int t = 0; t < T; ++t)
          if (t % SPLIT == SUB) { ph_add(a_one[t / SPLIT], fr[cu][t]); ph_fma(a_res[t / SPLIT], fr[cu][t], rv[cu]); }
#pragma unroll
        for (int ta = 0; ta < T; ++ta)
#pragma unroll
          for (int tb = ta; tb < T; ++tb) {
            const int idx = ta * T - ta * (ta - 1) / 2 + (tb - ta);
            if (idx % SPLIT == SUB) ph_dmma(acc[idx / SPLIT][0], acc[idx / SPLIT][1], fr[cu][ta], fr[cu][tb]);
          }
      }
    };
    if constexpr (SPLIT == 1) phase_b(std::integral_constant<int, 0>{});
    else if constexpr (SPLIT == 2) { if (sub == 0) phase_b(std::integral_constant<int, 0>{}); else phase_b(std::integral_constant<int, 1>{}); }
    else if constexpr (SPLIT == 4) {
      switch (sub) {
        case 0: phase_b(std::integral_constant<int, 0>{}); break;
        case 1: phase_b(std::integral_constant<int, 1>{}); break;
        case 2: phase_b(std::integral_constant<int, 2>{}); break;
        default: phase_b(std::integral_constant<int, 3>{}); break;
      }
    } else {
      switch (sub) {
        case 0: phase_b(std::integral_constant<int, 0>{}); break;
        case 1: phase_b(std::integral_constant<int, 1>{}); break;
        case 2: phase_b(std::integral_constant<int, 2>{}); break;
        case 3: phase_b(std::integral_constant<int, 3>{}); break;
        case 4: phase_b(std::integral_constant<int, 4>{}); break;
        case 5: phase_b(std::integral_constant<int, 5>{}); break;
        case 6: phase_b(std::integral_constant<int, 6>{}); break;
        default: phase_b(std::integral_constant<int, 7>{}); break;
      }
    }
  }
  ph_wait_all();
  __syncthreads();

  ph_epilogue<FAM, T, NTH>(acc, a_one, a_res, cnt, sr, rr, s_x, P, partials, out, ticket);
}

// cuTensorMapEncodeTiled through the runtime's driver entry point lookup (no link-time dependency on libcuda)
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) { cudaGetLastError(); p = nullptr; }
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

template <int FAM, int T, int LAYOUT, int NTH>
int launch_ph_layout(const Shard& s, const Params& P, const CUtensorMap& tmap, int wide_copies) {
  auto kern = k_gram_ph<FAM, T, LAYOUT, NTH>;
  constexpr int kPhThreads = NTH;
  constexpr int R = ph_rows(T, LAYOUT) * NTH / 512, PITCH = 8 * T + 4, QT = ph_qt(T), NB = (T + 1) / 2;
  constexpr size_t stage = LAYOUT == kTma ? (size_t)NB * R * 16 : (size_t)R * PITCH;
  const size_t smem = sizeof(double) * (2 * stage + 2 * R + R + (size_t)8 * T);
  static bool attr[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 64 && !attr[dev]) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return SO_ECUDA;
    attr[dev] = true;
  }
  const int64_t nchunks = (s.rows + R - 1) / R;
  const int grid = (int)std::min<int64_t>(std::max<int64_t>(1, nchunks), (int64_t)s.sm_count * (512 / NTH));
  const int Qo = P.n * (P.n + 1) / 2 + P.n + 1;
  if ((size_t)grid * QT > s.partials_doubles || (size_t)Qo > s.out_doubles) return SO_ENOMEM;
  kern<<<grid, kPhThreads, smem, s.stream>>>(s.X, s.y, s.rows, P, tmap, wide_copies, s.partials, s.out, s.ticket);
  return cudaGetLastError() == cudaSuccess ? 1 : SO_ECUDA;
}

// Threads per block: two blocks of 8 warps per SM up to 64 inputs (four of 4 warps up to 24) — their phases interleave,
// n = 32: 5.96 -> 5.70 ms, n = 16: 5.82 -> 4.80 ms per 1e8 / 2e8 rows — one block of 16 warps beyond, where the tile
// triangle is split over up to 8 warps (profiles/r2_gram_ph_v6_half.jsonl, r2_gram_ph_v7_nth.txt).
// Also measured and not kept: accumulating the next chunk's linear predictor inside the DMMA loop (three stages, one
// barrier per chunk, no separate phase A): n = 32 5.64 vs 5.48 ms, n = 40 7.47 vs 6.89 ms (profiles/r2_gram_ph_v8_fused.txt);
// and a three-stage ring with ONE barrier per chunk in which half of the warps (two per sub-partition) run "phase A of chunk
// c+1, then phase B of chunk c" and the other half the reverse, so that the latency-bound phase always has DMMA warps beside
// it: n = 32 5.34 -> 5.92 ms, logistic 5.81 -> 6.81, n = 16 4.79 -> 4.68, n = 24 3.85 -> 4.07 (profiles/r2_gram_ph_v9_pipe_roles.jsonl)
// — a phase-A warp's dependent DFMAs queue behind the DMMA bursts of its sub-partition, the k_gram_tri problem again.
__host__ __device__ constexpr int ph_threads(int T) { return T <= 3 ? 128 : T <= 8 ? 256 : 512; }

template <int FAM, int T>
int launch_ph(const Shard& s, const Params& P) {
  // 16-byte requests need 16-byte aligned rows: even f and an aligned first row (a view may start at an odd row)
  const bool aligned = (P.f % 2 == 0) && (reinterpret_cast<uintptr_t>(s.X) % 16 == 0) && (reinterpret_cast<uintptr_t>(s.y) % 16 == 0);
  alignas(64) CUtensorMap tmap;
  memset(&tmap, 0, sizeof tmap);
  static const bool no_tma = [] { const char* e = getenv("SCALAOPT_GRAM_NO_TMA"); return e && *e && *e != '0'; }();   // A/B switch
  constexpr int NTH = ph_threads(T);
  if constexpr (T <= 12) {
    if (aligned && !no_tma && s.rows < ((int64_t)1 << 31) && encode_tiled()) {
      constexpr int R = ph_rows(T, kTma) * NTH / 512;
      const cuuint64_t dims[2] = {(cuuint64_t)P.f, (cuuint64_t)s.rows};
      const cuuint64_t strides[1] = {(cuuint64_t)P.f * 8};
      const cuuint32_t box[2] = {16, (cuuint32_t)R};
      const cuuint32_t estr[2] = {1, 1};
      if (encode_tiled()(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(s.X), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS)
        return launch_ph_layout<FAM, T, kTma, NTH>(s, P, tmap, 1);
    }
  }
  return launch_ph_layout<FAM, T, kPadded, NTH>(s, P, tmap, aligned ? 1 : 0);
}

template <int FAM>
int dispatch_ph(const Shard& s, const Params& P) {
  switch ((P.f + 7) / 8) {
    case 1: return launch_ph<FAM, 1>(s, P);
    case 2: return launch_ph<FAM, 2>(s, P);
    case 3: return launch_ph<FAM, 3>(s, P);
    case 4: return launch_ph<FAM, 4>(s, P);
    case 5: return launch_ph<FAM, 5>(s, P);
    case 6: return launch_ph<FAM, 6>(s, P);
    case 7: return launch_ph<FAM, 7>(s, P);
    case 8: return launch_ph<FAM, 8>(s, P);
    case 9: return launch_ph<FAM, 9>(s, P);
    case 10: return launch_ph<FAM, 10>(s, P);
    case 11: return launch_ph<FAM, 11>(s, P);
    case 12: return launch_ph<FAM, 12>(s, P);
    case 13: return launch_ph<FAM, 13>(s, P);
  }
  return SO_EUNSUPPORTED;
}

}  // namespace

size_t gram_ph_partials_needed(int f, int sm_count) {
  const int T = (f + 7) / 8;
  return (size_t)sm_count * 4 * (size_t)ph_qt(T);
}

int launch_gram_ph(const Shard& s, int fam, const Params& P) {
  return fam == kLinear ? dispatch_ph<kLinear>(s, P) : dispatch_ph<kLogistic>(s, P);
}

}  // namespace wide
}  // namespace so
