// so_device.cuh — device-side building blocks shared by the evaluation kernels (sm_100a).
//
//  * so_exp_fast: fp64 exp with a 256-entry 2^(j/256) table in shared memory and a degree-4 polynomial:
//    8 FP64-pipe instructions + 4 others instead of ~25 + ~15 for libdevice exp.  The Gaussian-peak / exponential
//    families are FP64-issue-bound on B200 (64 DFMA/clk/SM vs 16 B/row at HBM speed), so the
//    transcendental is the first thing to shave.  Error vs libm on |x| < 708: <= 1.5 ulp + 3.4e-17 |x| relative
//    (the second term = a relative perturbation of the ARGUMENT by 0.3 ulp: less than the rounding error the
//    argument already carries).
//  * deterministic two-level reduction: fixed thread->row mapping, butterfly shuffles inside a warp,
//    shared memory across warps in warp order, per-block partials in global memory, and the block that
//    draws the last ticket adds the partials in block order.  The result does not depend on scheduling.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "so_internal.h"

namespace so {

constexpr int kExpTableSize = 256;

// 2^(j/256), j = 0..255, one copy per block in static shared memory (2 KB).  File scope on purpose: LDS then
// addresses it with an immediate offset instead of rebuilding the shared-window base at every call.
__shared__ double s_exp_tab[kExpTableSize];

// Constants live in the constant bank so that DFMA/DADD take them as c[bank][offset] operands: a 64-bit
// literal that is not representable as a 32-bit immediate otherwise costs two UMOV/IMAD.MOV per use.
struct ExpConsts { double inv, shift, nln2hi, c4; };
static __constant__ ExpConsts kExp = {
    0x1.71547652b82fep+8,      // 256 / ln 2
    6755399441055744.0,        // 1.5 * 2^52: round-to-nearest-integer by addition
    -0x1.62e42fefa39efp-9,     // -(ln 2 / 256) rounded to double (the 9.1e-20 it misses is the 3.4e-17 |x| above)
    0x1.5555555555555p-5,      // 1/24
};
// 1/6 with the low word zero: DFMA takes it as a 32-bit immediate (no register, no constant-bank slot next to c4).
// It is 2.4e-7 low; on |r| <= ln2/512 that moves the result by <= 1e-16.
#define SO_EXP_C3 __longlong_as_double(0x3FC5555500000000LL)

// |x| >= 708 (result would be subnormal / overflow), inf and NaN take the libdevice path.
constexpr int kExpHiLimit = 0x40862000;
__device__ __forceinline__ int exp_abs_hi(double x) { return __double2hiint(x) & 0x7fffffff; }
// The same test on the FP32 pipe: the high word of a double read as a float orders like the double's magnitude, and the
// high word of an inf/NaN double is a NaN float, so !(|hi| < limit) is "out of range".  One FSETP with a free |.| modifier
// per argument (predicates OR-ed) instead of LOP3 + integer max + ISETP; FP32 instructions are the cheapest thing to
// issue next to DFMAs (tools/dfma_mix.cu).
__device__ __forceinline__ bool exp_in_range(double x) { return fabsf(__int_as_float(__double2hiint(x))) < __int_as_float(kExpHiLimit); }

// Fill the table.  Call from every thread of the block, then __syncthreads().
// Entry j holds 2^(j/256) with (j << 12) subtracted from its high word: so_exp_fast then adds (k << 12) to the high
// word, and since k - j = 256 e exactly that is the entry scaled by 2^e — one integer instruction instead of a shift,
// a mask and an add.  (Non-FP64 instructions are not free next to DFMAs here: tools/dfma_mix.cu.)
__device__ __forceinline__ void exp_table_init() {
  for (int j = threadIdx.x; j < kExpTableSize; j += blockDim.x) {
    const double v = exp2((double)j * (1.0 / 256.0));
    s_exp_tab[j] = __hiloint2double(__double2hiint(v) - (j << 12), __double2loint(v));
  }
}

// exp(x) for |x| < 708, no range check, no branch: 8 FP64-pipe instructions + LOP3/SHL, LDS, IMAD.
//   k = round(256 x / ln 2),  r = x - k ln2/256 (|r| <= ln2/512),  exp(x) = 2^(k>>8) * T[k&255] * (1 + r + r^2/2 + r^3/6 + r^4/24)
// truncation error r^5/120 < 4e-17.
__device__ __forceinline__ double so_exp_fast(double x) {
  double kd = fma(x, kExp.inv, kExp.shift);
  const int ki = __double2loint(kd);
  kd -= kExp.shift;
  const double r = fma(kd, kExp.nln2hi, x);
  double q = fma(r, kExp.c4, SO_EXP_C3);
  q = fma(q, r, 0.5);
  q = fma(q, r, 1.0);
  const double pm1 = q * r;
  const double Tb = s_exp_tab[ki & (kExpTableSize - 1)];
  const double T = __hiloint2double(__double2hiint(Tb) + (ki << 12), __double2loint(Tb));   // 2^(k/256), |x| < 708: normal
  return fma(T, pm1, T);
}

// rare path (|x| >= 708, inf, NaN): a real call keeps the ~60 libdevice instructions out of the hot loop body
static __device__ __noinline__ double exp_slow(double x) { return exp(x); }

// ---- exp of a PRODUCT, argument given in units of ln2/256: so_exps(p, q) = exp(p q ln2/256) ----------------------
// The kernels' exp arguments are products (rate * t, -(zz * zz)); with the 256/ln2 folded into one factor on the host
// the product itself never has to be formed: k = round(p q) and r = p q - k both come out of one FMA each, so an
// exponential costs 8 FP64-pipe instructions INCLUDING its argument (so_exp_fast: 8 + the multiplication that forms
// x).  r is exact up to one rounding (no ln2 constant in the reduction any more); the polynomial is the same degree-4
// series in c r, c = ln2/256, with the powers of c folded into the coefficients (c3 rounded to a 32-bit immediate:
// 1.5e-7 relative on a term <= 4.1e-10, i.e. 6e-17 on the result).  Error vs libm of the exact product: <= 1.5 ulp; the
// rounding of the scaled factor perturbs the ARGUMENT by <= 1 ulp, as forming x = rate * t did before.
struct ExpSConsts { double shift, c1, c2, c4, unscale; };
static __constant__ ExpSConsts kExpS = {
    6755399441055744.0,        // 1.5 * 2^52
    0x1.62e42fefa39efp-9,      // c      = ln 2 / 256
    0x1.ebfbdff82c58fp-19,     // c^2 / 2
    0x1.3b2ab6fba4e77p-39,     // c^4 / 24
    0x1.62e42fefa39efp-9,      // scaled argument -> natural argument (slow path)
};
#define SO_EXPS_C3 __longlong_as_double(0x3E2C6B0900000000LL)      // c^3 / 6 = 0x1.c6b08d704a0c0p-29 to 32 bits
constexpr double kExpScale = 0x1.71547652b82fep+8;                 // 256 / ln 2: host-side factor of the scaled arguments
constexpr double kExpScaleSqrt = 0x1.337cc2183b050p+4;             // its square root (for arguments of the form -(zz)^2)
constexpr double kExpScaledLimit = 700.0 * 0x1.71547652b82fep+8;   // |p q| below this: so_exps is valid (708 less 1 % margin)
// valid for |p q| < kExpScaledLimit; the checked kernels test that on the factors (so_scalar.cu: hi_below)
__device__ __forceinline__ double so_exps(double p, double q) {
  double kd = fma(p, q, kExpS.shift);
  const int ki = __double2loint(kd);
  kd -= kExpS.shift;
  const double r = fma(p, q, -kd);                                 // |r| <= 1/2 (units of ln2/256)
  double s = fma(r, kExpS.c4, SO_EXPS_C3);
  s = fma(s, r, kExpS.c2);
  s = fma(s, r, kExpS.c1);
  const double pm1 = s * r;
  const double Tb = s_exp_tab[ki & (kExpTableSize - 1)];
  const double T = __hiloint2double(__double2hiint(Tb) + (ki << 12), __double2loint(Tb));
  return fma(T, pm1, T);
}
static __device__ __noinline__ double exps_slow(double p, double q) { return exp(p * q * kExpS.unscale); }

// exp(x) with the range check folded in (one branch); used where the check is not hoisted by the caller.
__device__ __forceinline__ double so_exp(double x) {
  if (!exp_in_range(x)) return exp_slow(x);
  return so_exp_fast(x);
}

// ---- cross-entropy link helpers: log1p(e) for e = exp(-|z|) in [0, 1] and 1/(1 + e) ----------------------------
// log(1 + e) with a 129-entry table of (1/c_j, log c_j), c_j = 1 + j/128: j = round(128 e), r = (e - j/128)/c_j,
// |r| <= 2^-8, log1p(r) to degree 6 (truncation r^7/7 < 2e-18): 11 FP64-pipe instructions + 1 LDS.128 where libdevice
// log1p takes ~40 with branches.  For e < 2^-8 (j = 0) it is the plain series in e: full relative accuracy near 0.
constexpr int kLogTableSize = 129;
__shared__ double2 s_log_tab[kLogTableSize];
struct LogConsts { double shift, n128, c6, c5, c4, c3; };
static __constant__ LogConsts kLog = { 6755399441055744.0, -1.0 / 128.0, -1.0 / 6.0, 1.0 / 5.0, -1.0 / 4.0, 1.0 / 3.0 };
__device__ __forceinline__ void log_table_init() {
  for (int j = threadIdx.x; j < kLogTableSize; j += blockDim.x) {
    const double c = 1.0 + (double)j * (1.0 / 128.0);
    s_log_tab[j] = make_double2(1.0 / c, log(c));
  }
}
__device__ __forceinline__ double so_log1p_unit(double e) {
  double kd = fma(e, 128.0, kLog.shift);
  const int j = __double2loint(kd);
  kd -= kLog.shift;
  const double2 t = s_log_tab[j];
  const double r = fma(kd, kLog.n128, e) * t.x;
  double q = fma(r, kLog.c6, kLog.c5);
  q = fma(q, r, kLog.c4);
  q = fma(q, r, kLog.c3);
  q = fma(q, r, -0.5);
  return fma(r * r, q, r) + t.y;
}
// 1/x for x in the normal range (here 1 + e in [1, 2]): MUFU seed (>= 23 bits) + two Newton steps, no branches
__device__ __forceinline__ double so_rcp_fast(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  r = fma(r, fma(-x, r, 1.0), r);
  return fma(r, fma(-x, r, 1.0), r);
}

// ---- mbarrier + TMA bulk copy (cp.async.bulk, SASS UBLKCP): global -> shared without touching registers ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
// bytes: multiple of 16; dst and src 16-byte aligned.  Completion is signalled on `bar` (complete_tx).
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// streaming 128-bit / 64-bit loads: read once, do not keep in L1
__device__ __forceinline__ double2 ld_stream2(const double2* p) { return __ldcs(p); }
__device__ __forceinline__ double ld_stream(const double* p) { return __ldcs(p); }

// lanes per output in the last block's walk over the per-block partials: the largest power of two <= 32 with
// outputs * lanes <= threads (1 when there are more outputs than threads / 2)
__device__ __forceinline__ int partials_lanes(int outputs, int threads) {
  int G = 32;
  while (G > 1 && outputs * G > threads) G >>= 1;
  return G;
}

// ---- so_finish: cross-GPU sum over NVLink peer memory + delivery to the host, by the block that folded the partials ----
// (protocol: Ctl in so_internal.h).  Called by ALL threads of that one block after it wrote out[0, Q).
__device__ __forceinline__ unsigned long long so_globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void st_relaxed_sys(double* p, double v) {
  asm volatile("st.relaxed.sys.global.f64 [%0], %1;" :: "l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ double ld_relaxed_sys(const double* p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ void so_finish(unsigned int* ticket, const double* out, int Q) {
  const Ctl* c = reinterpret_cast<const Ctl*>(ticket);
  if (c->mode == 0) return;
  const unsigned long long t0 = so_globaltimer();
  __syncthreads();                                      // out[] as written by this block's threads
  const unsigned long long seq = *reinterpret_cast<volatile unsigned long long*>(c->seq) + 1ull;
  const int W = c->world, me = c->rank, cap = c->cap;
  const int par = (int)(seq & 1ull);
  __shared__ int s_timed_out;
  if (threadIdx.x == 0) s_timed_out = 0;
  if (W > 1) {
    const size_t mine = (size_t)(par * kMaxRanks + me) * cap;
#pragma unroll 1
    for (int q = threadIdx.x; q < Q; q += blockDim.x) {
      const double v = __ldcg(out + q);
#pragma unroll 1
      for (int r = 0; r < W; ++r) st_relaxed_sys(c->slots[r] + mine + q, v);
    }
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < W) {
      st_release_sys(c->flags[threadIdx.x] + par * kMaxRanks + me, seq);           // my sums have landed at rank threadIdx.x
      const unsigned long long* f = c->flags[me] + par * kMaxRanks + threadIdx.x;  // ... and rank threadIdx.x's at mine?
      const unsigned long long give_up = so_globaltimer() + c->timeout_ns;
      while (ld_acquire_sys(f) < seq) {
        if (so_globaltimer() > give_up) { s_timed_out = 1; break; }
      }
    }
    __syncthreads();
#pragma unroll 1
    for (int q = threadIdx.x; q < Q; q += blockDim.x) {
      double s = 0.0;
#pragma unroll 1
      for (int r = 0; r < W; ++r) s += ld_relaxed_sys(c->slots[me] + (size_t)(par * kMaxRanks + r) * cap + q);
      c->host_out[q] = s;
    }
  } else {
#pragma unroll 1
    for (int q = threadIdx.x; q < Q; q += blockDim.x) c->host_out[q] = __ldcg(out + q);
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    *c->seq = seq;
    c->host_out[cap] = (double)t0;                      // device timestamps (ns): entry into so_finish, delivery
    c->host_out[cap + 1] = (double)so_globaltimer();
    if (s_timed_out) *reinterpret_cast<volatile unsigned long long*>(c->host_err) = seq;
    __threadfence_system();
    st_release_sys(c->host_flag, seq);
  }
}

// Reduce Q per-thread accumulators over the grid, deterministically.
//   acc[Q]      per-thread partial sums
//   smem        >= (blockDim.x/32) * Q doubles of shared memory
//   partials    gridDim.x * Q doubles of global workspace
//   out[q]      = scale[q] * sum_b partials[b][q], written by the last block to finish
//   ticket      global counter, zero before the launch, reset to zero by the last block
template <int Q>
__device__ __forceinline__ void grid_reduce(double (&acc)[Q], double* smem, double* __restrict__ partials,
                                            double* __restrict__ out, const double* __restrict__ scale,
                                            unsigned int* ticket, const signed char* __restrict__ src = nullptr,
                                            int qout = Q, const double* __restrict__ scale2 = nullptr) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
#pragma unroll
  for (int q = 0; q < Q; ++q) {
    const double s = warp_sum(acc[q]);
    if (lane == 0) smem[warp * Q + q] = s;
  }
  __syncthreads();
  for (int q = threadIdx.x; q < Q; q += blockDim.x) {
    double s = 0.0;
    for (int w = 0; w < nwarps; ++w) s += smem[w * Q + q];
    partials[(size_t)blockIdx.x * Q + q] = s;
  }
  __threadfence();
  __syncthreads();
  __shared__ unsigned int s_last;
  if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
  __syncthreads();
  if (s_last) {
    __threadfence();
    // out[q] = scale[q] * S[src[q]]  (+ scale2[q] * S[qout + q] when the result is the sum of two accumulator sets)
    // G lanes per output add the blocks b = sub, sub + G, ... in order and are then folded by a fixed butterfly: the order is
    // a function of (gridDim, Q) only, and a 5-output kernel does not walk 1184 partials with one thread per output
    // (that walk was 20 of the 25 us of a 1e6-row evaluation).
    const int G = partials_lanes(qout, blockDim.x);
    const int per_round = blockDim.x / G, sub = threadIdx.x % G;
    for (int q0 = 0; q0 < qout; q0 += per_round) {
      const int q = q0 + threadIdx.x / G;
      const bool act = q < qout;
      const int qs = act ? (src ? (int)src[q] : q) : 0;        // identical sums are kept once
      double s = 0.0, s2 = 0.0;
      if (act) {
        for (unsigned b = sub; b < gridDim.x; b += G) s += __ldcg(&partials[(size_t)b * Q + qs]);
        if (scale2) for (unsigned b = sub; b < gridDim.x; b += G) s2 += __ldcg(&partials[(size_t)b * Q + qout + q]);
      }
      for (int o = G >> 1; o >= 1; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
      if (act && sub == 0) {
        const double v = scale ? s * scale[q] : s;
        out[q] = scale2 ? fma(s2, scale2[q], v) : v;
      }
    }
    if (threadIdx.x == 0) *ticket = 0u;
    so_finish(ticket, out, qout);
  }
}

// same, run-time Q (accumulators already reduced to one value per (block, q) in `block_vals` shared memory)
__device__ __forceinline__ void grid_reduce_block_vals(const double* block_vals, int Q, double* __restrict__ partials,
                                                       double* __restrict__ out, const double* __restrict__ scale,
                                                       unsigned int* ticket) {
  for (int q = threadIdx.x; q < Q; q += blockDim.x) partials[(size_t)blockIdx.x * Q + q] = block_vals[q];
  __threadfence();
  __syncthreads();
  __shared__ unsigned int s_last2;
  if (threadIdx.x == 0) s_last2 = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
  __syncthreads();
  if (s_last2) {
    __threadfence();
    const int G = partials_lanes(Q, blockDim.x);
    const int per_round = blockDim.x / G, sub = threadIdx.x % G;
    for (int q0 = 0; q0 < Q; q0 += per_round) {
      const int q = q0 + threadIdx.x / G;
      double s = 0.0;
      if (q < Q) for (unsigned b = sub; b < gridDim.x; b += G) s += __ldcg(&partials[(size_t)b * Q + q]);
      for (int o = G >> 1; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (q < Q && sub == 0) out[q] = scale ? s * scale[q] : s;
    }
    if (threadIdx.x == 0) *ticket = 0u;
    so_finish(ticket, out, Q);
  }
}

}  // namespace so
