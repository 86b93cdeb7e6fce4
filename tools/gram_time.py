"""K2 timing over the mid-width Gram shapes (8 < n <= ~128): python tools/gram_time.py [tag]
Run with SCALAOPT_GRAM_LEGACY=1 for round 1's k_gram_tri / k_gram_ws on the same shapes.  One JSON line per shape."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scalaopt_b200 import native
HBM = 6554.2
try:
    HBM = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
tag = sys.argv[1] if len(sys.argv) > 1 else ("legacy" if os.environ.get("SCALAOPT_GRAM_LEGACY") else "ph")
shapes = [("linear_noint", 32, 100_000_000), ("logistic", 32, 100_000_000), ("linear", 31, 100_000_000),
          ("linear_noint", 16, 200_000_000), ("linear_noint", 24, 100_000_000), ("linear_noint", 40, 80_000_000),
          ("linear_noint", 48, 60_000_000), ("linear_noint", 64, 40_000_000), ("linear_noint", 80, 30_000_000),
          ("linear_noint", 96, 30_000_000), ("linear_noint", 104, 20_000_000), ("linear", 127, 20_000_000),
          ("linear_noint", 128, 20_000_000)]
if len(sys.argv) > 2:
    shapes = [s for s in shapes if str(s[1]) in sys.argv[2].split(",")]
with native.Context(1) as ctx:
    for name, f, m in shapes:
        fam = {"linear_noint": native.LINEAR_NOINT, "linear": native.LINEAR, "logistic": native.LOGISTIC}[name]
        n = f + (0 if name == "linear_noint" else 1)
        truth = np.arange(1, n + 1) / n
        ds = native.Dataset(ctx, m, f).generate(fam, truth)
        ks = []
        for i in range(8):
            ds.normal_eq(fam, truth * 0.9)
            ks.append(ctx.stats().kernel_ms)
        ks = ks[2:]
        ms = float(np.median(ks))
        gb = m * 8.0 * (f + 1) / 1e9
        T = (f + 7) // 8
        dmma_tf = m / 4 * (T * (T + 1) / 2) * 512 / (ms * 1e-3) / 1e12       # executed DMMA flops (tiles over the inputs)
        print(json.dumps({"tag": tag, "family": name, "f": f, "n": n, "rows": m, "kernel_ms": ms, "kernel_ms_best": float(min(ks)),
                          "hbm_frac": gb / (ms * 1e-3) / HBM, "useful_tflops": m * (n + 1.0) * (n + 2.0) / (ms * 1e-3) / 1e12,
                          "dmma_tflops_executed": dmma_tf, "dmma_frac_of_37.16": dmma_tf / 37.16}), flush=True)
        ds.free()
