"""TSQR timing: python tools/run_qr.py  -> one JSON line per shape (so_eval_qr next to so_eval_normal_eq on the same rows)"""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scalaopt_b200 import native
shapes = [("linear_noint", 4, 400_000_000), ("linear", 10, 100_000_000), ("linear_noint", 16, 100_000_000),
          ("linear_noint", 32, 50_000_000), ("logistic", 32, 50_000_000), ("linear_noint", 47, 20_000_000)]
with native.Context(1) as ctx:
    for name, f, m in shapes:
        fam = {"linear_noint": native.LINEAR_NOINT, "linear": native.LINEAR, "logistic": native.LOGISTIC}[name]
        n = f + (0 if name == "linear_noint" else 1)
        truth = np.arange(1, n + 1) / n
        ds = native.Dataset(ctx, m, f).generate(fam, truth)
        out = {"family": name, "f": f, "n": n, "rows": m}
        for what, fn in (("tsqr_ms", lambda: ds.qr(fam, truth * 0.9)), ("k2_ms", lambda: ds.normal_eq(fam, truth * 0.9))):
            ks = []
            for _ in range(5):
                fn()
                ks.append(ctx.stats().kernel_ms)
            out[what] = float(np.median(ks[1:]))
        out["tsqr_over_k2"] = out["tsqr_ms"] / out["k2_ms"]
        print(json.dumps(out), flush=True)
        ds.free()
