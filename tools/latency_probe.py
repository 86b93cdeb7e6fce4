"""Fixed cost of one blocking evaluation: wall time and kernel time of K1 / K2 at m = 1e3 .. 1e7 rows (linear n = 4 and the
Gaussian-peak + exponential-background family).  python tools/latency_probe.py"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scalaopt_b200 import native
with native.Context(1) as ctx:
    for name, fam, f, truth in (("linear n=4", native.LINEAR_NOINT, 4, np.array([0.25, 0.5, 0.75, 1.0])),
                                ("gauss_expbkg n=5", native.GAUSS_EXPBKG, 1, np.array([2.0, 0.5, 0.05, 1.0, 3.0]))):
        for m in (1_000, 10_000, 100_000, 1_000_000, 10_000_000):
            ds = native.Dataset(ctx, m, f).generate(fam, truth, noise_sigma=0.05)
            p = truth * 1.05
            out = {"family": name, "rows": m}
            for what, fn in (("k1", lambda: ds.loss_grad(fam, p)), ("k2", lambda: ds.normal_eq(fam, p))):
                for _ in range(20): fn()
                wall, kern = [], []
                for _ in range(200):
                    t0 = time.perf_counter(); fn(); wall.append((time.perf_counter() - t0) * 1e6); kern.append(ctx.stats().kernel_ms * 1e3)
                out[what + "_wall_us"] = round(float(np.median(wall)), 2); out[what + "_kernel_us"] = round(float(np.median(kern)), 2)
            print(json.dumps(out), flush=True)
            ds.free()
