#!/usr/bin/env python
"""bench.py — headline benchmark of the data-set evaluation hot path (BASELINE.json).

Metric: fp64 observations/s for the fused loss + J^T r + J^T J evaluation (K2, `so_eval_normal_eq`) on
BASELINE config 3: Levenberg-Marquardt Gaussian-peak + exponential-background fit, n = 5 parameters,
10^9 observations PER GPU (16 GB of fp64 (t, y) pairs resident in HBM; weak scaling over 1/2/4/8 B200).

  python bench.py --gpus 1 --steps K --warmup W              own arm (this repo's CUDA path through the C ABI)
  python bench.py --impl reference ...                       the reference's algorithm on the host cores
  torchrun --nproc-per-node N bench.py --gpus N ...          one rank per GPU, cross-GPU sum inside the evaluation kernel

One "step" = one evaluation of (loss, J^T r, J^T J) over the whole resident data set at a fresh parameter
vector.  `value` is timed with the observations already in HBM; `e2e` re-ingests the observations from
pinned host memory (H2D) inside every timed step and reads the result back, through the same C ABI.
Inputs are larger than L2 (16 GB vs 126 MB), so no L2 flush is needed between iterations.

Scaling.  The default line is WEAK scaling (10^9 observations per GPU).  BASELINE configs[2] read literally is STRONG
scaling — 10^9 observations in total on 1/2/4/8 GPUs — and every run measures that too, in the same process, on a second
resident data set of 10^9/N observations per GPU: `roofline.strong_scaling` (observations/s, per-call overhead) and
`e2e.lm` (Levenberg-Marquardt iterations/s, weak and strong).  `--scaling strong` makes the strong numbers the headline.
Before anything is timed, `config.parity` holds a rank-mode parity check of K1-K4 (+TSQR) against the CPU oracle on
10^6 observations per rank (tests/rank_parity.py), so every N of a scaling run carries its own correctness statement.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TRUTH = np.array([2.0, 0.5, 0.05, 1.0, 3.0])     # (A, mu, sigma, B, lambda), SURVEY.md §8d cfg3
NOISE = 0.05
SEED = 12345
METRIC = "fp64 observations/s, loss + J^T r + J^T J evaluation (LM normal equations, n=5)"
UNIT = "observations/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--rows", type=int, default=1_000_000_000, help="observations per GPU")
    ap.add_argument("--e2e-steps", type=int, default=None, help="timed steps of the e2e leg (default: min(steps, 5))")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-lm", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-strong", action="store_true")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --rows observations per GPU; strong: --rows-total observations over all GPUs")
    ap.add_argument("--rows-total", type=int, default=1_000_000_000, help="observations of the strong-scaling data set")
    ap.add_argument("--strong-steps", type=int, default=None, help="timed steps of the strong-scaling leg (default: max(steps, 200))")
    ap.add_argument("--parity-rows", type=int, default=1_000_000, help="observations per rank of the parity check")
    return ap.parse_args()


def params_for_step(i: int) -> np.ndarray:
    # a different, nearby parameter vector every step (what successive LM iterates look like)
    return TRUTH * (1.1 - 0.002 * (i % 50))


# ------------------------------------------------------------------------------------------------------
# clocks during the timed region (B200_PROFILING.md)
# ------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                 "-i", str(self.gpu)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        try:
            self.proc.terminate()
            self.proc.wait(timeout=5)
        except Exception:
            pass
        sm, mx, reasons, power = [], [], set(), []
        try:
            for line in open(self.path):
                c = [x.strip() for x in line.split(",")]
                if len(c) < 9:
                    continue
                try:
                    sm.append(float(c[1])); mx.append(float(c[2])); power.append(float(c[3]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                     c[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons),
                       samples=len(sm), power_w_max=float(max(power)))
        return out


# ------------------------------------------------------------------------------------------------------
# CPU legs (the oracle is the checker / the reported baseline, never the product path)
# ------------------------------------------------------------------------------------------------------
def host_threads() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return os.cpu_count() or 1


def cpu_model_name() -> str:
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def cpu_single_pass(threads: int, budget_s: float = 12.0):
    """oracle port, single-pass analytic normal equations, par[threads] — an optimistic stand-in for the
    reference (no JVM boxing, no finite differences, no 2n+1-pass QR)."""
    import oracle
    fam = oracle.GAUSS_EXPBKG
    m0 = 2_000_000
    X, y = oracle.generate(fam, m0, 1, TRUTH, SEED, NOISE)
    t0 = time.perf_counter()
    oracle.normal_eq(fam, X, y, params_for_step(0), parts=threads, threads=threads)
    rate0 = m0 / (time.perf_counter() - t0)
    m = int(min(max(rate0 * budget_s, m0), 200_000_000))
    if m > m0:
        X, y = oracle.generate(fam, m, 1, TRUTH, SEED, NOISE)
    t0 = time.perf_counter()
    oracle.normal_eq(fam, X, y, params_for_step(1), parts=threads, threads=threads)
    dt = time.perf_counter() - t0
    return m / dt, m


def reference_sample_rows(threads: int, target_s: float = 3.0) -> int:
    import oracle
    fam = oracle.GAUSS_EXPBKG
    m0 = 1_000_000
    X, y = oracle.generate(fam, m0, 1, TRUTH, SEED, NOISE)
    t0 = time.perf_counter()
    oracle.lm_reference_passes(fam, X, y, params_for_step(0), threads=threads)
    rate = m0 / (time.perf_counter() - t0)
    return int(min(max(rate * target_s, 200_000), 50_000_000))


def run_reference(args):
    """--impl reference: the reference's OWN algorithm for this evaluation — finite-difference Jacobian rows
    (MSEFunction.scala:117-136), the 2n+1-pass Householder QR over the data set (QR.scala:131-250) and the loss
    pass (LevenbergMarquardt.scala:164) — restated in C (oracle/oracle.c; the Scala original cannot run: no JVM
    here or on the GPU box), partitions folded on all host threads like Spark local[N]."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    oracle.build()
    threads = host_threads()
    fam = oracle.GAUSS_EXPBKG
    # bounded sample: the whole --steps/--warmup run is sized to ~2 minutes of CPU work
    per_step_s = min(3.0, max(0.05, 120.0 / (args.steps + args.warmup)))
    m = reference_sample_rows(threads, target_s=per_step_s)
    X, y = oracle.generate(fam, m, 1, TRUTH, SEED, NOISE)
    for i in range(args.warmup):
        oracle.lm_reference_passes(fam, X, y, params_for_step(i), threads=threads)
    t0 = time.perf_counter()
    for i in range(args.steps):
        oracle.lm_reference_passes(fam, X, y, params_for_step(i), threads=threads)
    dt = time.perf_counter() - t0
    value = m * args.steps / dt
    sample = (f"{m} observations per step (bounded sample of the 1e9-per-GPU workload), reference algorithm: "
              f"FD Jacobian + 11-pass QR + loss pass, C restatement, {threads} threads, {cpu_model_name()}")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": dict(workload_config(args, args.gpus), reference_sample_rows_per_step=m,
                       note=f"CPU arm: rate measured on a bounded sample of {m} observations per step, not on 1e9"),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, world, rows_per_gpu=None, scaling="weak"):
    rows = args.rows if rows_per_gpu is None else rows_per_gpu
    return {
        "workload": "BASELINE configs[2]: Levenberg-Marquardt Gaussian-peak + exponential-background fit, "
                    "n=5 params, fp64 (t,y) observations, normal-equation evaluation (loss + J^T r + J^T J)",
        "family": "gauss_expbkg", "n_params": 5, "rows_per_gpu": rows, "rows_total": rows * world,
        "scaling": scaling, "bytes_per_row": 16,
        "parallelism": f"dp{world} (row shards; the 21-double cross-GPU sum is done by the evaluation kernel's folding "
                       f"block over NVLink peer memory, results delivered to mapped host memory)",
        "l2_policy": "inputs larger than L2 (>= 2 GB per GPU vs 126 MB), no flush needed",
    }


# ------------------------------------------------------------------------------------------------------
# own arm
# ------------------------------------------------------------------------------------------------------
class K2Loop:
    """K blocking so_eval_normal_eq calls through the C ABI with every argument marshalled once (what a JNI caller with
    pinned arrays does): the per-call cost of the harness is one ctypes foreign call."""

    def __init__(self, native, ctx, ds, fam):
        import ctypes as C
        self.fn = native.load().so_eval_normal_eq
        self.ctx, self.h, self.fam = ctx, ds._h, fam
        dp = C.POINTER(C.c_double)
        self.params = [np.ascontiguousarray(params_for_step(i)) for i in range(50)]
        self.pp = [a.ctypes.data_as(dp) for a in self.params]
        self.JtJ, self.Jtr, self.rtr = np.empty((5, 5)), np.empty(5), C.c_double()
        self.pJ, self.pr, self.prt = self.JtJ.ctypes.data_as(dp), self.Jtr.ctypes.data_as(dp), C.byref(self.rtr)

    def run(self, steps, first=0):
        fn, h, fam, pp, pJ, pr, prt = self.fn, self.h, self.fam, self.pp, self.pJ, self.pr, self.prt
        for i in range(first, first + steps):
            rc = fn(h, fam, pp[i % 50], 5, pJ, pr, prt)
            if rc != 0:
                self.ctx.check(rc)


def time_k2(native, ctx, ds, fam, steps, warmup, barrier, max_over_ranks, sampler=None):
    """-> dict: wall seconds (max over ranks) of `steps` evaluations, CUDA-event kernel/device time, time inside the
    in-kernel exchange, launches."""
    loop = K2Loop(native, ctx, ds, fam)
    loop.run(warmup)
    ctx.stats_reset()
    launches0 = ctx.stats().total_launches
    if sampler is not None:
        sampler.start()
    barrier()
    t0 = time.perf_counter()
    loop.run(steps, first=warmup)
    barrier()
    wall_s = time.perf_counter() - t0
    clocks = sampler.stop() if sampler is not None else {}
    st = ctx.stats()
    return {"wall_s": max_over_ranks(wall_s), "kernel_ms": max_over_ranks(st.sum_kernel_ms),
            "device_ms": max_over_ranks(st.sum_device_ms), "finish_ms": max_over_ranks(st.sum_finish_ms),
            "kernel_ms_best": st.min_kernel_ms, "launches": int(st.total_launches - launches0), "clocks": clocks,
            "finished_on_device": int(st.finished_on_device)}


def run_own(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    import torch.distributed as dist

    torch.cuda.set_device(local_rank)
    uid = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from scalaopt_b200 import native
    native.load()
    if world > 1:
        box = [native.Context.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        uid = box[0]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    fam = native.GAUSS_EXPBKG
    ctx = native.Context(rank=rank, world=world, device_id=local_rank, nccl_uid=uid)

    # ---- parity in rank mode, before anything is timed (the oracle is the checker here, nothing else) ----
    parity = None
    if not args.no_parity:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import rank_parity
        allsum, allgather = rank_parity.torch_collectives(dist, world)
        parity = rank_parity.check(native, ctx, rank, world, args.parity_rows, allsum, allgather)
        parity["what"] = ("K1-K4 + TSQR through the C ABI in rank mode vs the CPU oracle summed over ranks, relative errors; "
                          "tolerances 1e-12 (loss, gradient, phi), 1e-11 (J^T J, J^T r, dphi, Hd, R^T R)")
        if not parity["ok"]:
            raise SystemExit(f"rank-mode parity FAILED: {parity}")

    strong_rows = args.rows_total // world
    strong_rows -= strong_rows & 1
    headline_strong = args.scaling == "strong"
    m = strong_rows if headline_strong else args.rows
    ds = native.Dataset(ctx, m, 1).generate(fam, TRUTH, seed=SEED, noise_sigma=NOISE, row_offset=rank * m)
    total_rows = ds.size()
    assert total_rows == m * world

    # ---- value: observations resident in HBM ----
    sampler = ClockSampler(local_rank) if rank == 0 else None
    head = time_k2(native, ctx, ds, fam, args.steps, args.warmup, barrier, max_over_ranks, sampler)
    wall_s, launches, clocks = head["wall_s"], head["launches"], head["clocks"]
    kernel_ms_max, device_ms, kernel_ms_best = head["kernel_ms"], head["device_ms"], head["kernel_ms_best"]
    value = total_rows * args.steps / wall_s

    # ---- strong scaling: the SAME 10^9-observation data set on 1/2/4/8 GPUs ----
    strong = None
    ds_s = None
    if not args.no_strong:
        if headline_strong or (world == 1 and m == strong_rows):
            ds_s = ds
        else:
            ds_s = native.Dataset(ctx, strong_rows, 1).generate(fam, TRUTH, seed=SEED, noise_sigma=NOISE,
                                                                row_offset=rank * strong_rows)
        ssteps = args.strong_steps if args.strong_steps is not None else max(args.steps, 200)
        r = time_k2(native, ctx, ds_s, fam, ssteps, max(args.warmup, 5), barrier, max_over_ranks)
        per_call_ms = 1e3 * r["wall_s"] / ssteps
        strong = {
            "rows_total": strong_rows * world, "rows_per_gpu": strong_rows, "n_gpus": world, "steps": ssteps,
            "observations_per_s": strong_rows * world * ssteps / r["wall_s"], "ms_per_call": per_call_ms,
            "kernel_ms_per_call": r["kernel_ms"] / ssteps, "kernel_ms_fastest": r["kernel_ms_best"],
            # kernel_ms includes the in-kernel exchange (and the wait for the slowest GPU) when finished_on_device
            "exchange_ms_per_call": r["finish_ms"] / ssteps,
            "host_overhead_ms_per_call": per_call_ms - r["kernel_ms"] / ssteps,
            "finished_on_device": r["finished_on_device"],
        }

    # ---- LM iterations/s through the host-side mirror of LevenbergMarquardt.minimize ----
    lm = None
    if not args.no_lm:
        try:
            from scalaopt_b200 import host as sohost
            lm = sohost.bench_lm(ctx, ds, fam, TRUTH * 1.1, barrier, max_over_ranks)
            # the same fit with one K1 loss pass per trial point + one K2 pass per outer iteration (no speculation)
            plain = sohost.bench_lm(ctx, ds, fam, TRUTH * 1.1, barrier, max_over_ranks, speculate=False)
            lm["without_speculation"] = {k: plain[k] for k in ("seconds", "iterations", "iterations_per_s", "k2_passes", "k1_passes")}
            lm["rows_total"] = total_rows
            if ds_s is not None:
                if ds_s is ds:
                    lm["strong"] = {k: lm[k] for k in ("seconds", "iterations", "iterations_per_s", "k2_passes", "k1_passes")}
                else:
                    ls = sohost.bench_lm(ctx, ds_s, fam, TRUTH * 1.1, barrier, max_over_ranks)
                    lm["strong"] = {k: ls[k] for k in ("seconds", "iterations", "iterations_per_s", "k2_passes", "k1_passes", "minimiser")}
                lm["strong"]["rows_total"] = strong_rows * world
        except Exception as e:  # the evaluation numbers stand on their own
            lm = {"unavailable": f"{type(e).__name__}: {e}"}
    if ds_s is not None and ds_s is not ds:
        ds_s.free()

    # ---- e2e: host buffers, H2D of the observations + evaluation + result read-back, every step ----
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, native, ctx, ds, fam, m, world, barrier, max_over_ranks)

    # ---- CPU baseline beside it (rank 0, N = 1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        import oracle
        oracle.build()
        threads = host_threads()
        v, ms = cpu_single_pass(threads, budget_s=8.0)
        v1, ms1 = cpu_single_pass(1, budget_s=4.0)
        mref = reference_sample_rows(threads, target_s=2.0)
        Xr, yr = oracle.generate(oracle.GAUSS_EXPBKG, mref, 1, TRUTH, SEED, NOISE)
        t0 = time.perf_counter()
        oracle.lm_reference_passes(oracle.GAUSS_EXPBKG, Xr, yr, params_for_step(0), threads=threads)
        vref = mref / (time.perf_counter() - t0)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{ms} observations, single-pass analytic loss+J^T r+J^T J, oracle par[{threads}] "
                         f"({cpu_model_name()}); the JVM/Spark reference cannot run here (no JVM)",
               "seq": {"value": v1, "unit": UNIT, "cores": 1,
                       "sample": f"{ms1} observations, same single-pass evaluation, oracle seq (models SeqDataSet)"},
               "reference_algorithm": {"value": vref, "unit": UNIT, "cores": threads,
                                       "sample": f"{mref} observations, FD Jacobian + 11-pass QR + loss pass "
                                                 f"(what LevenbergMarquardt.iterate does per outer iteration)"}}

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md 6650 GB/s)"
        alg_bytes = 16.0 * m                                   # per launch: one GPU's shard, 16 B per observation
        ach = alg_bytes / (kernel_ms_max / args.steps * 1e-3) / 1e9
        # DRAM bytes of one launch from the committed `ncu --set full` capture (only if it was taken at this launch size)
        traffic, traffic_src, traffic_at = None, None, None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "k2_traffic.json")))
            if int(tr["rows_per_launch"]) == int(m):
                traffic = float(tr["dram_read_bytes"]) + float(tr["dram_write_bytes"])
                traffic_src = tr["source"]
                traffic_at = tr.get("measured_at")
        except Exception:
            pass
        # north_star's roofline is the SLOWER of bytes at HBM bandwidth and flops at FP64 peak; this kernel executes 42 FP64
        # instructions per observation (SASS of k_scalar<OpNormalEq<FamGaussExpBkg>>), which at the DFMA
        # rate measured on this pool (profiles/r1_fp64_peak.json) takes longer than its 16 bytes at the HBM rate
        fp64 = None
        try:
            pk = json.load(open(os.path.join(ROOT, "profiles", "r1_fp64_peak.json")))
            dfma_per_s = float(pk["dfma_tflops_burst"]) * 1e12 / 2.0
            instr = 42.0 * m
            bound_ms = instr / dfma_per_s * 1e3
            fp64 = {"fp64_instr_per_observation": 42, "dfma_peak_instr_per_s": dfma_per_s, "bound_ms_per_launch": bound_ms,
                    "hbm_bound_ms_per_launch": alg_bytes / (hbm_peak * 1e9) * 1e3,
                    "frac_of_fp64_bound": bound_ms / (kernel_ms_max / args.steps),
                    "note": "explanatory only; roofline.frac above is against HBM"}
        except Exception:
            pass
        if strong is not None:
            sb = 16.0 * strong["rows_per_gpu"]
            strong["achieved_gbs_per_gpu"] = sb / (strong["kernel_ms_per_call"] * 1e-3) / 1e9
            strong["frac_of_hbm"] = strong["achieved_gbs_per_gpu"] / hbm_peak
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * wall_s / args.steps, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, world, m, args.scaling),
            "device_ms_per_step": device_ms / args.steps, "kernel_ms_per_step": kernel_ms_max / args.steps,
            "roofline": {"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                         "traffic": traffic, "traffic_source": traffic_src, "traffic_measured_at": traffic_at,
                         "peak_source": peak_src,
                         "kernel": "k_scalar<OpNormalEq<FamGaussExpBkg>>", "algorithmic_bytes_per_launch": alg_bytes,
                         # the timed region runs under the board's power cap (see "clocks"); the fastest single launch of the
                         # region is reported beside the average the roofline fraction is computed from
                         "fastest_launch": {"kernel_ms": kernel_ms_best, "achieved": alg_bytes / (kernel_ms_best * 1e-3) / 1e9,
                                            "frac": alg_bytes / (kernel_ms_best * 1e-3) / 1e9 / hbm_peak},
                         "exchange_ms_per_step": head["finish_ms"] / args.steps,
                         "host_overhead_ms_per_step": 1e3 * wall_s / args.steps - kernel_ms_max / args.steps,
                         "fp64": fp64},
            "clocks": clocks, "gpu_launches": int(launches),
        }
        if strong is not None:
            line["roofline"]["strong_scaling"] = strong
            line["strong"] = strong
        if parity is not None:
            line["config"]["parity"] = parity
        if e2e is not None:
            line["e2e"] = e2e
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if lm is not None:
            line["lm"] = lm
            if e2e is not None:
                # the driver keeps the sub-keys of `e2e`: LM iterations/s (the metric's second half) travels there too
                line["e2e"]["lm"] = {k: lm[k] for k in ("iterations", "iterations_per_s", "seconds", "rows_total", "strong") if k in lm}
        print(json.dumps(line), flush=True)
    ds.free()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def run_e2e(args, native, ctx, ds, fam, m, world, barrier, max_over_ranks):
    """Every timed step: so_dataset_clear + so_dataset_append from PINNED HOST buffers (H2D of all of this rank's
    observations) + so_eval_normal_eq with the result landing in host memory."""
    steps = args.e2e_steps if args.e2e_steps is not None else min(args.steps, 5)
    # bound the pinned allocation by what the host can spare
    avail = 0
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable"):
                avail = int(line.split()[1]) * 1024
    except Exception:
        pass
    rows = m
    if avail:
        rows = int(min(m, max(1_000_000, (avail // (4 * world)) // 16)))
    rows -= rows & 1
    hx, hy = native.PinnedBuffer(rows), native.PinnedBuffer(rows)
    # fill the host buffers with this rank's own observations (device -> host, outside the timed region)
    CH = 1 << 26
    for b in range(0, rows, CH):
        e = min(rows, b + CH)
        X, y = ds.read(b, e - b)
        hx.array[b:e] = X[:, 0]
        hy.array[b:e] = y
    ds2 = ds if rows == m else native.Dataset(ctx, rows, 1)
    xa, ya = hx.array.ctypes.data, hy.array.ctypes.data

    def step(i):
        ds2.clear()
        ds2.append_raw(xa, ya, rows)
        return ds2.normal_eq(fam, params_for_step(i))

    step(0)
    barrier()
    t0 = time.perf_counter()
    for i in range(steps):
        step(i)
    barrier()
    dt = max_over_ranks(time.perf_counter() - t0)
    numa = None
    try:
        import torch
        pr = torch.cuda.get_device_properties(torch.cuda.current_device())
        bus = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        numa = {"gpu_pci": bus, "gpu_numa_node": int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())}
    except Exception:
        pass
    out = {"value": rows * world * steps / dt, "unit": UNIT, "h2d_bytes_per_step": 16 * rows * world,
           "h2d_gbs_per_gpu": 16 * rows * steps / dt / 1e9, "h2d_gbs_aggregate": 16 * rows * world * steps / dt / 1e9,
           "host_numa": numa,
           "d2h_bytes_per_step": 21 * 8, "steps": steps, "ms_per_step": 1e3 * dt / steps,
           "rows_per_gpu": rows,
           "what": "so_dataset_clear + so_dataset_append(pinned host -> HBM) + so_eval_normal_eq per step"}
    if ds2 is not ds:
        ds2.free()
    hx.free(); hy.free()
    return out


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_own(args)


if __name__ == "__main__":
    main()
