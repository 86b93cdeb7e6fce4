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
//   f <= 104    k_gram_ph (so_gram_ph.cu): shared-memory staged DMMA SYRK in block-wide phases; DMMA = mma.sync.m8n8k4.f64,
//               the A and B fragments of tile (ta, tb) are the SAME registers — lane l holds Z[row0 + l%4][8t + l/4].
//               Measured on this pool's B200: DMMA 37.2 TFLOP/s vs 33.8 for register-resident DFMA
//               (profiles/r1_fp64_peak.json), and one DMMA replaces 8 DFMA issue slots + their operand traffic.
//               (Round 1's k_gram_tri — fragments straight from global memory, residual chain in the same warp — is gone:
//               slower at every width, profiles/r2_gram_legacy.jsonl vs r2_gram_ph_final.jsonl.)
//   larger n    warp-specialised shared-memory staged DMMA SYRK over 4x4-tile patches (k_gram_ws below), f <= 512
#include <algorithm>
#include <cstdlib>
#include <vector>

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
      const double inv = so_rcp_fast(1.0 + e);
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
// n + 1 > 48 (f up to 512 inputs): warp-specialised, shared-memory staged DMMA SYRK (k_gram_ws).
//   * The DMMA tiles cover the f INPUT columns only (T = ceil(f/8) tile columns).  The upper triangle of tiles is
//     cut into 4x4-tile patches (32x32 entries); a consumer warp owns one patch (16 DMMA accumulators = 64
//     registers).  4 producer warps + 12 consumer warps per 512-thread block, one block per SM.
//   * The intercept's 1-column and the residual column are NOT padded into a further tile column.  The sums
//     sum x_c and sum x_c res are 4 more DMMAs per slab on the DIAGONAL patches (X^T [1 res 0 ..] into their
//     unused below-diagonal accumulators), which brings those warps from 10 to 14 DMMAs per slab next to the 16 of
//     an off-diagonal patch.  n = 257: 560 DMMAs per 4-row slab instead of 666.
//   * The residuals themselves need the linear predictor, a matrix-vector product that DMMA would do at 1/8
//     efficiency, so it is DFMA work.  A DFMA of another warp does not get through a saturated DMMA stream
//     (measured: a producer warp's DFMA waited for the whole DMMA phase of its block, 1 300 cycles per
//     instruction), so it cannot be hidden behind the DMMAs; instead all 16 warps compute the residuals of the
//     NEXT chunk together as a short phase before the chunk barrier (warp w takes rows w, w+16, ...; a
//     transposing shuffle reduction), and chunks are made as long as shared memory allows (2-slot ring).
//   * Rows are streamed into a 3-stage ring, row pitch 8T+4 doubles so that the 4 rows x 8 columns of an MMA
//     fragment hit distinct banks.  f even: one TMA bulk copy per row (cp.async.bulk, 8f bytes), issued by a
//     single producer warp and completed on the stage's mbarrier (complete_tx).  f odd (rows only 8-byte
//     aligned): 8-byte cp.async by all producer threads, completed on the same mbarrier with
//     cp.async.mbarrier.arrive.noinc.  One block barrier per chunk releases the stage.
//   * grid = (tile groups) x (row workers); tile group is the fastest block index so that the blocks streaming
//     the same rows run together and share them through L2 (the kernel is FP64-bound: n = 257 needs 66 306 flop
//     per 2 056-byte row).  Spare consumer warps share a patch with another warp and split the slabs of a chunk
//     between them ("reps"); ws_plan() balances the DMMA count over the four SM sub-partitions.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
constexpr int kWsConsumers = 12, kWsProducers = 4, kWsThreads = 32 * (kWsConsumers + kWsProducers);
constexpr int kPatch = 4, kWsMaxStages = 3, kWsMaxRows = 128;
constexpr int kPatchDoubles = kPatch * kPatch * 64;
// below-diagonal accumulator slots of a diagonal patch that hold X^T [1 res] of its tile rows 0..3
__host__ __device__ constexpr int ws_extra_slot(int ti) { return ti == 0 ? 1 * kPatch + 0 : ti == 1 ? 2 * kPatch + 0 : ti == 2 ? 3 * kPatch + 0 : 2 * kPatch + 1; }

__device__ __forceinline__ void cp_async8(void* dst_smem, const void* src, bool valid) {
  const unsigned int sz = valid ? 8u : 0u;      // src-size 0: zero fill
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" :: "r"(smem_u32(dst_smem)), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void bar_block() { asm volatile("bar.sync 0;" ::: "memory"); }

constexpr int kWsMaxGroups = 12, kWsMaxPatches = 136;      // f <= 512: PT <= 16
struct WsSlot { unsigned char patch, rep, reps, pad; };     // patch 255: idle warp
struct WsPlan {
  int T, PT, npatches, groups, reps, workers, krows, stages, QE;     // reps = max over patches
  size_t smem;
  WsSlot slot[kWsMaxGroups][kWsConsumers];                   // consumer warp -> (patch, slab residue class)
  unsigned char preps[kWsMaxPatches];                        // warps sharing a patch
};

template <int FAM>
__global__ void __launch_bounds__(kWsThreads, 1)
k_gram_ws(const double* __restrict__ X, const double* __restrict__ y, int64_t rows, const __grid_constant__ Params P,
          const __grid_constant__ WsPlan L, double* __restrict__ partials, double* __restrict__ out, unsigned int* ticket) {
  extern __shared__ __align__(16) double sm[];
  __shared__ double s_sc[kWsThreads / 32][2];
  __shared__ __align__(8) uint64_t s_full[kWsMaxStages];
  const int f = P.f, icpt = P.icpt, n = P.n;
  const bool bulk = (f & 1) == 0;                        // rows 16-byte aligned: TMA bulk copies
  const int T = L.T, PT = L.PT, npatches = L.npatches, reps = L.reps, krows = L.krows, stages = L.stages;
  const int W = 8 * T, PW = W + 4;                       // row pitch in doubles
  double* ps = sm;                                       // [W] parameters of the input columns (0 beyond f)
  double* res_s = sm + W;                                // [2][kWsMaxRows] residuals of the chunk consumed / the next one
  double* y_s = res_s + 2 * kWsMaxRows;                  // [kWsMaxStages][kWsMaxRows] y of the chunks in the ring
  double* stage0 = y_s + kWsMaxStages * kWsMaxRows;      // L.stages ring slots of krows x PW
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int worker = blockIdx.y, nworkers = gridDim.y, group = blockIdx.x;
  if (FAM == kLogistic) exp_table_init();
  for (int c = tid; c < W; c += kWsThreads) ps[c] = (c < f) ? P.p[icpt + c] : 0.0;
  for (int i = tid; i < stages * krows * PW; i += kWsThreads) stage0[i] = 0.0;     // padding columns stay zero
  if (tid == 0) {
    for (int st = 0; st < stages; ++st) mbar_init(&s_full[st], (bulk ? 1 : kWsProducers * 32) + 1);   // + warp 1: y
    mbar_fence_init();
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // the zeros above are ordered before the bulk copies
  __syncthreads();
  const int64_t nchunks = (rows + krows - 1) / krows;
  double* tile_partials = partials;
  double* extra_partials = partials + (size_t)nworkers * reps * npatches * kPatchDoubles;

  const bool producer = warp < kWsProducers;
  const double p0 = icpt ? P.p[0] : 0.0;

  // ---- producers: stream chunk `chunk` (rows and y) into ring slot `stage`
  auto load_stage = [&](int64_t chunk, int stage) {
    if (chunk >= nchunks) return;
    double* st = stage0 + (size_t)stage * krows * PW;
    const int64_t row0 = chunk * krows;
    const int nvalid = (int)min((int64_t)krows, rows - row0);
    if (warp == 1) {                                       // y of the chunk: plain stores, released by an arrival
      for (int r = lane; r < krows; r += 32) y_s[stage * kWsMaxRows + r] = (r < nvalid) ? ld_stream(y + row0 + r) : 0.0;
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_full[stage]);
    }
    if (bulk) {
      if (warp != 0) return;
      for (int r = nvalid; r < krows; ++r)                 // ragged last chunk: rows past the end read as zero
        for (int c = lane; c < f; c += 32) st[r * PW + c] = 0.0;
      __syncwarp();
      if (lane == 0) mbar_arrive_expect_tx(&s_full[stage], (uint32_t)nvalid * (uint32_t)f * 8u);   // releases the zeros too
      __syncwarp();
      for (int r = lane; r < nvalid; r += 32) bulk_g2s(st + r * PW, X + (row0 + r) * (int64_t)f, (uint32_t)f * 8u, &s_full[stage]);
    } else {
      for (int r = warp; r < krows; r += kWsProducers) {
        const bool valid = r < nvalid;
        const double* src = X + (valid ? row0 + r : 0) * (int64_t)f + lane;
        double* dst = st + r * PW + lane;
        for (int c0 = 0; c0 < f; c0 += 128) {
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (c0 + 32 * u + lane < f) cp_async8(dst + c0 + 32 * u, src + c0 + 32 * u, valid);
        }
      }
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" :: "r"(smem_u32(&s_full[stage])) : "memory");
    }
  };

  // ---- all 16 warps: residuals of chunk `chunk` (ring slot it % stages) -> res_s[it & 1]; warp w takes rows
  //      w, w+16, w+32, w+48 of the chunk.  This is the only FP64-pipe (non-DMMA) work of the kernel and it runs as
  //      its own short phase between two DMMA phases: a DFMA of another warp does not get through a saturated
  //      DMMA stream.
  double sr = 0.0, rr = 0.0;                               // lanes 0..NR-1: the rows this warp finished
  const int nr = (krows - warp + 15) >> 4;                 // rows of a chunk this warp finishes (0..8)
  auto residual_rows = [&]<int NR>(int64_t chunk, int it, int slot) {
    const double* st = stage0 + (size_t)slot * krows * PW + warp * PW;
    double v[NR];
#pragma unroll
    for (int i = 0; i < NR; ++i) v[i] = 0.0;
#pragma unroll 2
    for (int c = lane; c < f; c += 32) {
      const double pc = ps[c];
#pragma unroll
      for (int i = 0; i < NR; ++i) v[i] = fma(pc, st[16 * i * PW + c], v[i]);
    }
    // transposing reduction (NR a power of two): every step halves the number of values a lane carries; at the
    // end lane l holds the full sum of row slot `slot_of(l)`; then slot i is fetched by lane i.
    double dot;
    if constexpr (NR == 8 || NR == 4 || NR == 2) {
      int width = 16;
      int myslot = 0;
#pragma unroll
      for (int n = NR; n > 1; n >>= 1) {
        const bool hi = (lane & width) != 0;
#pragma unroll
        for (int i = 0; i < n / 2; ++i) {
          const double send = hi ? v[i] : v[i + n / 2], keep = hi ? v[i + n / 2] : v[i];
          v[i] = keep + __shfl_xor_sync(0xffffffffu, send, width);
        }
        myslot += hi ? n / 2 : 0;
        width >>= 1;
      }
      dot = v[0];
#pragma unroll
      for (int o = 16 / NR; o >= 1; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
      // slot i lives in the lanes whose top log2(NR) bits spell i (bit 4 = NR/2, bit 3 = NR/4, ...)
      int src = 0;
#pragma unroll
      for (int n = NR, bit = 16; n > 1; n >>= 1, bit >>= 1) src |= (lane & (n / 2)) ? bit : 0;
      dot = __shfl_sync(0xffffffffu, dot, src);
      (void)myslot;
    } else {
#pragma unroll
      for (int i = 0; i < NR; ++i) v[i] = warp_sum(v[i]);
      dot = v[0];
#pragma unroll
      for (int i = 1; i < NR; ++i) if (lane == i) dot = v[i];
    }
    const int r = warp + 16 * lane;                        // lanes 0..NR-1 finish one row each
    if (lane < NR) {
      const double z = dot + p0, yv = y_s[slot * kWsMaxRows + r];
      double res;
      if (FAM == kLinear) {
        res = yv - z;
      } else {
        const double e = so_exp(-fabs(z));
        const double inv = so_rcp_fast(1.0 + e);
        res = ((z >= 0.0) ? inv : e * inv) - yv;
      }
      if (chunk * krows + r >= rows) res = 0.0;
      res_s[(it & 1) * kWsMaxRows + r] = res;
      sr += res;
      rr = fma(res, res, rr);
    }
  };
  auto residual_phase = [&](int64_t chunk, int it) {
    if (chunk >= nchunks) return;
    const int slot = it % stages;
    mbar_wait(&s_full[slot], (uint32_t)(it / stages) & 1u);
    switch (nr) {
      case 1: residual_rows.template operator()<1>(chunk, it, slot); break;
      case 2: residual_rows.template operator()<2>(chunk, it, slot); break;
      case 3: residual_rows.template operator()<3>(chunk, it, slot); break;
      case 4: residual_rows.template operator()<4>(chunk, it, slot); break;
      case 5: residual_rows.template operator()<5>(chunk, it, slot); break;
      case 6: residual_rows.template operator()<6>(chunk, it, slot); break;
      case 7: residual_rows.template operator()<7>(chunk, it, slot); break;
      case 8: residual_rows.template operator()<8>(chunk, it, slot); break;
      default: break;
    }
  };

  // ---- consumers: this warp's patch
  const int rsel = lane & 3, csel = lane >> 2;
  const int cw = warp - kWsProducers;
  int patch = 0, rep = 0, nrep = 1;
  bool active = false;
  if (!producer) {
    const WsSlot sl = L.slot[group][cw];
    patch = sl.patch; rep = sl.rep; nrep = sl.reps; active = sl.patch != 255;
  }
  int pa = 0, pb = 0;
  if (active) { int rem = patch; while (rem >= PT - pa) { rem -= PT - pa; ++pa; } pb = pa + rem; }
  const int na = min(kPatch, T - kPatch * pa), nb = min(kPatch, T - kPatch * pb);   // real tiles of the patch
  const int offa = 8 * kPatch * pa + csel, offb = 8 * kPatch * pb + csel;
  const int mode = !active ? 0 : (pa == pb ? (na == kPatch ? 1 : 3) : ((na == kPatch && nb == kPatch) ? 2 : 3));
  double acc[kPatch][kPatch][2];
#pragma unroll
  for (int i = 0; i < kPatch; ++i)
#pragma unroll
    for (int j = 0; j < kPatch; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
  const int nslabs = krows >> 2;

  int64_t chunk = worker;
  if (producer) for (int a = 0; a < stages - 1; ++a) load_stage(chunk + a * (int64_t)nworkers, a);
  residual_phase(chunk, 0);
  bar_block();
  for (int it = 0; chunk < nchunks; chunk += nworkers, ++it) {
    if (producer) {
      load_stage(chunk + (stages - 1) * (int64_t)nworkers, (it + stages - 1) % stages);   // the slot released by the last barrier
    } else if (mode != 0) {
      const double* st = stage0 + (size_t)(it % stages) * krows * PW + rsel * PW;
      const double* rs = res_s + (it & 1) * kWsMaxRows + rsel;
      mbar_wait(&s_full[it % stages], (uint32_t)(it / stages) & 1u);
      if (mode == 2) {                                     // full off-diagonal patch: 8 fragments, 16 DMMAs per slab
#pragma unroll 2
        for (int slab = rep; slab < nslabs; slab += nrep) {
          const double* base = st + slab * 4 * PW;
          double fa[kPatch], fb[kPatch];
#pragma unroll
          for (int i = 0; i < kPatch; ++i) { fa[i] = base[offa + 8 * i]; fb[i] = base[offb + 8 * i]; }
#pragma unroll
          for (int i = 0; i < kPatch; ++i)
#pragma unroll
            for (int j = 0; j < kPatch; ++j) dmma(acc[i][j][0], acc[i][j][1], fa[i], fb[j]);
        }
      } else if (mode == 1) {                              // full diagonal patch: 4 fragments, 10 + 4 DMMAs per slab
#pragma unroll 2
        for (int slab = rep; slab < nslabs; slab += nrep) {
          const double* base = st + slab * 4 * PW;
          double fa[kPatch];
#pragma unroll
          for (int i = 0; i < kPatch; ++i) fa[i] = base[offa + 8 * i];
          const double bx = (csel == 0) ? 1.0 : (csel == 1 ? rs[slab * 4] : 0.0);     // [1 res 0 0 0 0 0 0]
#pragma unroll
          for (int i = 0; i < kPatch; ++i)
#pragma unroll
            for (int j = i; j < kPatch; ++j) dmma(acc[i][j][0], acc[i][j][1], fa[i], fa[j]);
#pragma unroll
          for (int i = 0; i < kPatch; ++i) {
            constexpr int kSlot[4] = {ws_extra_slot(0), ws_extra_slot(1), ws_extra_slot(2), ws_extra_slot(3)};
            dmma(acc[kSlot[i] / kPatch][kSlot[i] % kPatch][0], acc[kSlot[i] / kPatch][kSlot[i] % kPatch][1], fa[i], bx);
          }
        }
      } else {                                             // ragged patch at the right / bottom edge
        for (int slab = rep; slab < nslabs; slab += nrep) {
          const double* base = st + slab * 4 * PW;
          double fa[kPatch], fb[kPatch];
#pragma unroll
          for (int i = 0; i < kPatch; ++i) {
            fa[i] = (i < na) ? base[offa + 8 * i] : 0.0;
            fb[i] = (i < nb) ? base[offb + 8 * i] : 0.0;
          }
#pragma unroll
          for (int i = 0; i < kPatch; ++i)
#pragma unroll
            for (int j = 0; j < kPatch; ++j)
              if (i < na && j < nb && (pa != pb || j >= i)) dmma(acc[i][j][0], acc[i][j][1], fa[i], fb[j]);
          if (pa == pb) {
            const double bx = (csel == 0) ? 1.0 : (csel == 1 ? rs[slab * 4] : 0.0);
#pragma unroll
            for (int i = 0; i < kPatch; ++i) {
              constexpr int kSlot[4] = {ws_extra_slot(0), ws_extra_slot(1), ws_extra_slot(2), ws_extra_slot(3)};
              if (i < na) dmma(acc[kSlot[i] / kPatch][kSlot[i] % kPatch][0], acc[kSlot[i] / kPatch][kSlot[i] % kPatch][1], fa[i], bx);
            }
          }
        }
      }
    }
    residual_phase(chunk + nworkers, it + 1);              // one chunk ahead of the DMMAs
    // generic-proxy reads of this slot are ordered before the bulk copies that refill it after the barrier (the loads have
    // completed by now — their values went into DMMAs — so the fence is free here; see k_scalar in so_scalar.cu)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    bar_block();                                           // everyone is done with this chunk
  }
  if (!bulk) asm volatile("cp.async.wait_all;" ::: "memory");

  // per-(row worker, rep, patch) partials, lane-major tiles: [worker*reps + rep][patch][i][j][lane*2 + reg]
  if (active) {
    double* dst = tile_partials + (((size_t)worker * reps + rep) * npatches + patch) * kPatchDoubles;
#pragma unroll
    for (int i = 0; i < kPatch; ++i)
#pragma unroll
      for (int j = 0; j < kPatch; ++j) {
        dst[(i * kPatch + j) * 64 + lane * 2] = acc[i][j][0];
        dst[(i * kPatch + j) * 64 + lane * 2 + 1] = acc[i][j][1];
      }
  }
  if (group == 0) {                                        // row count, sum res, sum res^2 of this worker
    sr = warp_sum(sr); rr = warp_sum(rr);
    if (lane == 0) { s_sc[warp][0] = sr; s_sc[warp][1] = rr; }
    __syncthreads();
    if (tid == 0) {
      double a = 0.0, b = 0.0;
      for (int w = 0; w < kWsThreads / 32; ++w) { a += s_sc[w][0]; b += s_sc[w][1]; }
      int64_t cnt = 0;
      for (int64_t c = worker; c < nchunks; c += nworkers) cnt += min((int64_t)krows, rows - c * krows);
      double* ex = extra_partials + (size_t)worker * L.QE;
      ex[0] = (double)cnt; ex[1] = a; ex[2] = b;
    }
  }
  // ---- grid barrier (cooperative launch: every block is resident), then ALL blocks fold the partials: block b emits
  //      outputs b*512 + tid, + nblocks*512, ...  The order of the sum over (worker, rep) is fixed: deterministic.
  const unsigned int nblocks = gridDim.x * gridDim.y, bid = blockIdx.y * gridDim.x + blockIdx.x;
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    atomicAdd(&ticket[1], 1u);
    unsigned int seen;
    do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(&ticket[1]) : "memory"); } while (seen < nblocks);
  }
  __syncthreads();
  {
    // emit the API layout: packed upper triangle over parameters (a <= b), then Jtr, then rtr.
    constexpr int ONE = -1, RES = -2;
    const int Tn = n * (n + 1) / 2, Qo = Tn + n + 1;
    for (int q = bid * kWsThreads + tid; q < Qo; q += nblocks * kWsThreads) {
      int ca, cb;
      double sgn = 1.0;
      if (q < Tn) {
        // row-major packed upper index -> (a, b): closed form, then fix up
        int a = (int)((2.0 * n + 1.0 - sqrt((2.0 * n + 1.0) * (2.0 * n + 1.0) - 8.0 * q)) * 0.5);
        while (a > 0 && a * n - a * (a - 1) / 2 > q) --a;
        while ((a + 1) * n - (a + 1) * a / 2 <= q) ++a;
        const int b = a + (q - (a * n - a * (a - 1) / 2));
        ca = icpt ? (a == 0 ? ONE : a - 1) : a;
        cb = icpt ? (b == 0 ? ONE : b - 1) : b;
      } else if (q < Tn + n) {
        const int a = q - Tn;
        ca = icpt ? (a == 0 ? ONE : a - 1) : a;
        cb = RES;
        if (FAM == kLinear) sgn = -1.0;                    // J r = -x~ rho for least squares
      } else {
        ca = RES; cb = RES;
      }
      double s2 = 0.0;
      if (ca < 0 && cb < 0) {                              // row count, sum res, sum res^2: the producers' scalars
        const int off = (ca == ONE && cb == ONE) ? 0 : ((ca == RES && cb == RES) ? 2 : 1);
        for (int w = 0; w < nworkers; ++w) s2 += __ldcg(&extra_partials[(size_t)w * L.QE + off]);
      } else {
        size_t off;
        if (ca >= 0 && cb >= 0) {
          if ((ca >> 3) > (cb >> 3)) { const int tmp = ca; ca = cb; cb = tmp; }   // tile row <= tile col
          const int qa = ca / (8 * kPatch), qb = cb / (8 * kPatch);
          const int pidx = qa * PT - qa * (qa - 1) / 2 + (qb - qa);
          const int ti = (ca >> 3) - qa * kPatch, tj = (cb >> 3) - qb * kPatch, i = ca & 7, j = cb & 7;
          off = (size_t)pidx * kPatchDoubles + (ti * kPatch + tj) * 64 + (i * 4 + (j >> 1)) * 2 + (j & 1);
        } else {                                           // sum x_c (j = 0) or sum x_c res (j = 1): the diagonal patch of c
          const int c = ca >= 0 ? ca : cb, j = (ca == ONE || cb == ONE) ? 0 : 1;
          const int qa = c / (8 * kPatch), pidx = qa * PT - qa * (qa - 1) / 2;
          const int ti = (c >> 3) - qa * kPatch, i = c & 7;
          off = (size_t)pidx * kPatchDoubles + ws_extra_slot(ti) * 64 + (i * 4) * 2 + j;
        }
        const int pr = L.preps[off / kPatchDoubles];
#pragma unroll 4
        for (int w = 0; w < nworkers; ++w)
          for (int e = 0; e < pr; ++e) s2 += __ldcg(&tile_partials[((size_t)w * reps + e) * npatches * kPatchDoubles + off]);
      }
      out[q] = s2 * sgn;
    }
  }
  __syncthreads();
  if (tid == 0 && atomicAdd(&ticket[2], 1u) == nblocks - 1) {    // last block out: everyone is past the barrier
    ticket[1] = 0u; ticket[2] = 0u;
    __threadfence();
  }
}

// Which warp works on what.  Patches are dealt to the tile groups by weight (DMMAs per slab: 16 for a full
// off-diagonal patch, 10 + 4 for a diagonal one, less at the ragged edge); inside a group spare warps take over slab
// residue classes of the heaviest patches, and the 12 consumer warps are placed so that the four SM sub-partitions
// (warp id mod 4: each has its own FP64/DMMA pipe) carry the same number of DMMAs.
inline WsPlan ws_plan(int f, int sm_count) {
  WsPlan b{};
  b.T = (f + 7) / 8;
  b.PT = (b.T + kPatch - 1) / kPatch;
  b.npatches = b.PT * (b.PT + 1) / 2;
  b.groups = (b.npatches + kWsConsumers - 1) / kWsConsumers;
  b.workers = std::max(1, sm_count / b.groups);
  const size_t W = (size_t)8 * b.T, PW = W + 4;
  const size_t fixed = sizeof(double) * (W + (2 + kWsMaxStages) * kWsMaxRows), budget = 227 * 1024 - 4096;   // static shared: exp table, flags
  b.stages = 2;
  int krows = (int)((budget - fixed) / (sizeof(double) * b.stages * PW));
  krows = std::min(kWsMaxRows, krows) & ~3;
  b.krows = krows;
  b.QE = 4;
  b.smem = fixed + sizeof(double) * b.stages * (size_t)std::max(krows, 0) * PW;

  // patch weights
  std::vector<int> weight(b.npatches);
  {
    int p = 0;
    for (int pa = 0; pa < b.PT; ++pa)
      for (int pb = pa; pb < b.PT; ++pb, ++p) {
        const int na = std::min(kPatch, b.T - kPatch * pa), nb = std::min(kPatch, b.T - kPatch * pb);
        weight[p] = pa == pb ? na * (na + 1) / 2 + na : na * nb;
      }
  }
  // patches -> groups: heaviest first into the lightest group that still has a free warp
  std::vector<int> order(b.npatches);
  for (int p = 0; p < b.npatches; ++p) order[p] = p;
  std::stable_sort(order.begin(), order.end(), [&](int x, int y2) { return weight[x] > weight[y2]; });
  std::vector<std::vector<int>> members(b.groups);
  std::vector<int> gload(b.groups, 0);
  for (int p : order) {
    int best = -1;
    for (int g = 0; g < b.groups; ++g)
      if ((int)members[g].size() < kWsConsumers && (best < 0 || gload[g] < gload[best])) best = g;
    members[best].push_back(p);
    gload[best] += weight[p];
  }
  const int nslabs = std::max(1, krows / 4);
  b.reps = 1;
  for (int g = 0; g < b.groups; ++g) {
    for (int cw = 0; cw < kWsConsumers; ++cw) b.slot[g][cw] = WsSlot{255, 0, 1, 0};
    std::vector<int> reps(members[g].size(), 1);
    int warps = (int)members[g].size();
    while (warps < kWsConsumers) {                       // spare warps split the heaviest share
      int best = -1;
      for (size_t i = 0; i < members[g].size(); ++i)
        if (reps[i] < nslabs && (best < 0 || weight[members[g][i]] * reps[best] > weight[members[g][best]] * reps[i])) best = (int)i;
      if (best < 0) break;
      ++reps[best]; ++warps;
    }
    struct Item { int patch, rep, reps; double w; };
    std::vector<Item> items;
    for (size_t i = 0; i < members[g].size(); ++i) {
      b.preps[members[g][i]] = (unsigned char)reps[i];
      b.reps = std::max(b.reps, reps[i]);
      for (int e = 0; e < reps[i]; ++e) items.push_back(Item{members[g][i], e, reps[i], (double)weight[members[g][i]] / reps[i]});
    }
    std::stable_sort(items.begin(), items.end(), [](const Item& x, const Item& y2) { return x.w > y2.w; });
    double load[4] = {0, 0, 0, 0};
    int used[4] = {0, 0, 0, 0};
    for (const Item& it : items) {                       // heaviest first into the lightest sub-partition with a free warp
      int best = -1;
      for (int k = 0; k < 4; ++k)
        if (used[k] < kWsConsumers / 4 && (best < 0 || load[k] < load[best])) best = k;
      const int cw = best + 4 * used[best];              // consumer warp cw runs on sub-partition cw % 4 (4 producer warps first)
      b.slot[g][cw] = WsSlot{(unsigned char)it.patch, (unsigned char)it.rep, (unsigned char)it.reps, 0};
      load[best] += it.w;
      ++used[best];
    }
  }
  return b;
}

template <int FAM>
int launch_ws(const Shard& s, const Params& P) {
  if (P.f > SO_MAX_FEATURES) return SO_EUNSUPPORTED;
  WsPlan b = ws_plan(P.f, s.sm_count);
  if (b.krows < 4) return SO_EUNSUPPORTED;
  auto kern = k_gram_ws<FAM>;
  static bool attr[64] = {false};
  int dev = 0; cudaGetDevice(&dev); dev &= 63;
  if (!attr[dev]) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 4096) != cudaSuccess) return SO_ECUDA;
    attr[dev] = true;
  }
  const int64_t nchunks = (s.rows + b.krows - 1) / b.krows;
  b.workers = (int)std::min<int64_t>(b.workers, std::max<int64_t>(1, nchunks));
  const size_t need = (size_t)b.workers * b.reps * b.npatches * kPatchDoubles + (size_t)b.workers * b.QE;
  const int Qo = P.n * (P.n + 1) / 2 + P.n + 1;
  if (need > s.partials_doubles || (size_t)Qo > s.out_doubles) return SO_ENOMEM;
  dim3 grid(b.groups, b.workers);                        // <= sm_count blocks, one per SM: co-resident (grid barrier)
  const double* Xp = s.X; const double* yp = s.y; int64_t rows = s.rows;
  double* partials = s.partials; double* outp = s.out; unsigned int* ticket = s.ticket;
  void* args[] = {&Xp, &yp, &rows, const_cast<Params*>(&P), &b, &partials, &outp, &ticket};
  if (cudaLaunchCooperativeKernel((const void*)kern, grid, dim3(kWsThreads), args, b.smem, s.stream) != cudaSuccess) return SO_ECUDA;
  s.device_finish = false;      // every block emits a slice of the result: the cross-GPU sum is a launch of its own (k_finish)
  return 1;
}

}  // namespace

// so_gram_ph.cu
size_t gram_ph_partials_needed(int f, int sm_count);
int launch_gram_ph(const Shard& s, int fam, const Params& P);
constexpr int kPhMaxF = 104;

size_t gram_partials_needed(int n, int f, int sm_count) {
  if (n <= 8) return (size_t)sm_count * 8 * (size_t)(n * (n + 1) / 2 + n + 1);
  if (f <= kPhMaxF) return gram_ph_partials_needed(f, sm_count);
  const WsPlan b = ws_plan(f, sm_count);
  return (size_t)b.workers * b.reps * b.npatches * kPatchDoubles + (size_t)b.workers * b.QE;
}

int launch_gram(const Shard& s, int fam, const Params& P) {
  if (P.n <= 8) return fam == kLinear ? dispatch_small<kLinear>(s, P) : dispatch_small<kLogistic>(s, P);
  if (P.f <= kPhMaxF) return launch_gram_ph(s, fam, P);
  return fam == kLinear ? launch_ws<kLinear>(s, P) : launch_ws<kLogistic>(s, P);
}

}  // namespace wide
}  // namespace so
