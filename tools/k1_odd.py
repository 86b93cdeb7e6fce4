"""K1 / K3 / K4 on odd and even row lengths: python tools/k1_odd.py [f,f,...]   (SCALAOPT_CUDA_LIB selects another build)"""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scalaopt_b200 import native
HBM = 6554.2
with native.Context(1) as ctx:
    only = [int(v) for v in sys.argv[1].split(",")] if len(sys.argv) > 1 else None
    for name, f, m in (("linear", 31, 100_000_000), ("linear_noint", 32, 100_000_000), ("linear", 127, 20_000_000), ("linear_noint", 128, 20_000_000), ("logistic", 33, 80_000_000), ("linear", 35, 80_000_000), ("linear_noint", 40, 80_000_000), ("logistic", 32, 100_000_000), ("linear_noint", 256, 10_000_000), ("linear", 47, 60_000_000), ("linear_noint", 48, 60_000_000), ("linear", 63, 50_000_000), ("linear_noint", 64, 50_000_000), ("linear_noint", 72, 40_000_000), ("linear", 150, 20_000_000), ("linear_noint", 300, 10_000_000), ("linear", 9, 200_000_000), ("linear", 10, 200_000_000), ("logistic", 12, 200_000_000), ("linear_noint", 16, 200_000_000), ("linear", 15, 200_000_000)):
        if only and f not in only or (only and name == 'logistic' and f == 33 and 33 not in only):
            continue
        fam = {"linear_noint": native.LINEAR_NOINT, "linear": native.LINEAR, "logistic": native.LOGISTIC}[name]
        n = f + (0 if name == "linear_noint" else 1)
        truth = np.arange(1, n + 1) / n
        ds = native.Dataset(ctx, m, f).generate(fam, truth)
        out = {"lib": os.path.basename(os.environ.get("SCALAOPT_CUDA_LIB", "default")) + ("+small" + os.environ["SCALAOPT_SMALL_MAX_N"] if os.environ.get("SCALAOPT_SMALL_MAX_N") else ""), "family": name, "f": f, "rows": m}
        for what, fn in (("k1_ms", lambda: ds.loss_grad(fam, truth * 0.9)), ("k4_ms", lambda: ds.hess_vec(fam, truth * 0.9, truth)),
                         ("k3_1_ms", lambda: ds.line(fam, truth * 0.9, truth * 0.01, [0.5])),
                         ("k3_4_ms", lambda: ds.line(fam, truth * 0.9, truth * 0.01, [0.0, 0.5, 1.0, 2.0]))):
            ks = []
            for _ in range(6):
                fn(); ks.append(ctx.stats().kernel_ms)
            out[what] = float(np.median(ks[1:])); out[what.replace("_ms", "_hbm_frac")] = m * 8.0 * (f + 1) / 1e9 / (out[what] * 1e-3) / HBM
        print(json.dumps(out), flush=True)
        ds.free()
