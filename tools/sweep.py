#!/usr/bin/env python
"""tools/sweep.py — kernel-level sweep over the BASELINE configs on one B200: every evaluation kernel at the
sizes of configs 2-5, CUDA-event kernel time (so_stats), achieved GB/s against the measured HBM peak and, for the
compute-bound Gram, TFLOP/s against the measured FP64 peak; plus optimizer iterations/s through the host mirror.
Prints one JSON object per line; `python tools/sweep.py > profiles/r1_sweep.jsonl`."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from scalaopt_b200 import host, native  # noqa: E402

HBM = 6554.2
try:
    HBM = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
FP64_TF = 37.16   # profiles/r1_fp64_peak.json (DMMA); DFMA 33.78


def timed(ctx, fn, reps=10, warm=2):
    for _ in range(warm):
        fn()
    ks = []
    for _ in range(reps):
        fn()
        ks.append(ctx.stats().kernel_ms)
    return float(np.median(ks)), float(np.min(ks))


def report(**kw):
    print(json.dumps(kw), flush=True)


def run_family(ctx, name, fam, m, f, truth, sigma, kinds, flops_gram=None):
    ds = native.Dataset(ctx, m, f).generate(fam, truth, noise_sigma=sigma)
    n = len(truth)
    p = np.asarray(truth) * 0.95 + 0.01
    d = np.cos(np.arange(n) + 1.0) * 0.01
    gb = m * 8.0 * (f + 1) / 1e9
    for kind in kinds:
        if kind == "K1":
            fn = lambda: ds.loss_grad(fam, p)
        elif kind == "K1loss":
            fn = lambda: ds.loss_grad(fam, p, want_grad=False)
        elif kind == "K2":
            fn = lambda: ds.normal_eq(fam, p)
        elif kind == "TSQR":
            fn = lambda: ds.qr(fam, p)
        elif kind.startswith("K3"):
            k = int(kind.split("k=")[1])
            al = [0.0, 1.0, 2.0, 4.0, 8.0, 0.5, 0.25, 0.125][:k]
            fn = lambda: ds.line(fam, p, d, al)
        else:
            fn = lambda: ds.hess_vec(fam, p, d)
        med, best = timed(ctx, fn)
        rec = dict(config=name, kernel=kind, rows=m, f=f, n=n, kernel_ms=med, kernel_ms_best=best,
                   rows_per_s=m / (med * 1e-3), gbs=gb / (med * 1e-3), hbm_frac=gb / (med * 1e-3) / HBM)
        if kind == "K2" and flops_gram:
            tf = m * flops_gram / (med * 1e-3) / 1e12
            rec.update(tflops=tf, fp64_frac=tf / FP64_TF)
        report(**rec)
    return ds


def main():
    which = sys.argv[1:] or ["cfg3", "cfg2", "cfg4", "cfg5"]
    with native.Context(1) as ctx:
        if "cfg3" in which:
            truth = [2.0, 0.5, 0.05, 1.0, 3.0]
            ds = run_family(ctx, "cfg3 gauss_expbkg 1e9", native.GAUSS_EXPBKG, 1_000_000_000, 1, truth, 0.05,
                            ["K2", "TSQR", "K1", "K1loss", "K3 k=1", "K3 k=2", "K3 k=4", "K4"])
            for spec in (True, False):
                lm = host.bench_lm(ctx, ds, native.GAUSS_EXPBKG, np.array(truth) * 1.1, speculate=spec)
                report(config="cfg3 LevenbergMarquardt gauss_expbkg 1e9", **lm)
            ds.free()
            ds = run_family(ctx, "cfg3b expdecay 1e9", native.EXPDECAY, 1_000_000_000, 1, [2.0, -1.3], 0.1, ["K2", "K1", "K4"])
            ds.free()
            ds = run_family(ctx, "cfg3c gauss 1e9", native.GAUSS, 1_000_000_000, 1, [2.0, 0.5, 0.05], 0.05, ["K2", "K1"])
            ds.free()
            ds = run_family(ctx, "poly n=4 1e9", native.POLY, 1_000_000_000, 1, [1.0, -2.0, 0.5, 3.0], 0.1, ["K2", "K1"])
            ds.free()
        if "cfg2" in which:
            f = 32
            truth = [(((j % 7) - 3) / 4.0) for j in range(f + 1)]
            m = 100_000_000
            ds = run_family(ctx, "cfg2 logistic 1e8x32", native.LOGISTIC, m, f, truth, 1.0,
                            ["K1", "K1loss", "K3 k=1", "K3 k=4", "K3 k=8", "K4", "K2"], flops_gram=34 * 35 + 0.0)
            # BFGS through the host mirror, BFGSConfig defaults, start 0
            fobj = host.Objective.gpu(ctx, ds, native.LOGISTIC, f + 1)
            t0 = time.perf_counter()
            x, res = host.minimize(host.BFGS, fobj, np.zeros(f + 1))
            dt = time.perf_counter() - t0
            c = fobj.counters()
            report(config="cfg2 BFGS logistic 1e8x32", iterations=int(res.iterations), seconds=dt,
                   iterations_per_s=res.iterations / dt, max_abs_err_vs_truth=float(np.abs(x - truth).max()), counters=c)
            fobj.free(); ds.free()
        if "cfg4" in which:
            f = 256
            truth = [(j + 1) / 256.0 for j in range(f)]
            m = 10_000_000
            ds = run_family(ctx, "cfg4 linear 1e7x256", native.LINEAR_NOINT, m, f, truth, 1.0,
                            ["K1", "K4", "K2"], flops_gram=257 * 258 + 0.0)
            fobj = host.Objective.gpu(ctx, ds, native.LINEAR_NOINT, f)
            t0 = time.perf_counter()
            x, res = host.minimize(host.NEWTON_CG, fobj, np.zeros(f))
            dt = time.perf_counter() - t0
            report(config="cfg4 Newton-CG linear 1e7x256", iterations=int(res.iterations), seconds=dt,
                   iterations_per_s=res.iterations / dt, max_abs_err_vs_truth=float(np.abs(x - truth).max()),
                   counters=fobj.counters())
            fobj.free()
            # Steihaug trust region (gradient/SteihaugCG.scala:51-148), delta0 = 1 as SURVEY.md 8d cfg4 asks
            fobj = host.Objective.gpu(ctx, ds, native.LINEAR_NOINT, f)
            cfg = host.default_config(host.STEIHAUG_CG)
            cfg.delta0 = 1.0
            t0 = time.perf_counter()
            try:
                x, res = host.minimize(host.STEIHAUG_CG, fobj, np.zeros(f), cfg)
                outcome = "converged"
            except host.MaxIterException as e:      # thrown, not returned, like the reference
                x, res, outcome = None, None, f"MaxIterException: {e}"
            dt = time.perf_counter() - t0
            c = fobj.counters()
            report(config="cfg4 Steihaug-CG linear 1e7x256 (delta0=1)", outcome=outcome, seconds=dt,
                   iterations=int(res.iterations) if res is not None else None,
                   iterations_per_s=(res.iterations / dt) if res is not None else None,
                   max_abs_err_vs_truth=float(np.abs(x - truth).max()) if x is not None else None,
                   passes_per_s=(c["k1_passes"] + c["k4_passes"]) / dt, counters=c)
            fobj.free()
            lm = host.bench_lm(ctx, ds, native.LINEAR_NOINT, np.zeros(f), repeats=1, runs=1)
            report(config="cfg4 LevenbergMarquardt linear 1e7x256 (compute-bound Gram)", **{k: v for k, v in lm.items() if k != "minimiser"})
            ds.free()
        if "cfg5" in which:
            for n, m in ((4, 1_000_000_000), (4, 100_000_000), (4, 1_000_000), (32, 100_000_000), (32, 1_000_000),
                         (128, 100_000_000), (128, 1_000_000)):
                truth = [(j + 1) / n for j in range(n)]
                ds = run_family(ctx, f"cfg5 linear n={n} m={m:.0e}", native.LINEAR_NOINT, m, n, truth, 1.0,
                                ["K1", "K2"] + (["TSQR"] if n <= 8 else []), flops_gram=(n + 1) * (n + 2) + 0.0)
                ds.free()


if __name__ == "__main__":
    main()
