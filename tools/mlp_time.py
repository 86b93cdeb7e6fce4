"""kernel times of the MLP family: python tools/mlp_time.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scalaopt_b200 import native
rng = np.random.default_rng(0)
with native.Context(1) as ctx:
    for f, h, m in ((2, 2, 400_000_000), (4, 4, 200_000_000)):
        n = h * (f + 1) + h + 1
        ds = native.Dataset(ctx, m, f)
        blk = 4_000_000
        X = rng.standard_normal((blk, f)); y = (rng.random(blk) < 0.5).astype(float)
        for _ in range(m // blk): ds.append(X, y)
        p = rng.uniform(-1, 1, n)
        for fam in (native.MLP_MSE, native.MLP_CE):
            for _ in range(3): ds.loss_grad(fam, p)
            k1 = ctx.stats().kernel_ms
            k2 = None
            if n <= 9:
                for _ in range(3): ds.normal_eq(fam, p)
                k2 = ctx.stats().kernel_ms
            gb = 8e-6 * (f + 1) * m
            print(f"f={f} h={h} fam={fam} rows={m}: K1 {k1:.3f} ms ({gb / k1:.0f} GB/s)" + (f", K2 {k2:.3f} ms ({gb / k2:.0f} GB/s)" if k2 else ""))
        ds.free()
