"""timing of so_eval_qr against so_eval_normal_eq: python tools/run_qr.py <family> <rows>"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scalaopt_b200 import native
name, m = sys.argv[1], int(sys.argv[2])
fam = getattr(native, name)
truth = {"GAUSS_EXPBKG": [2.0, 0.5, 0.05, 1.0, 3.0], "GAUSS": [2.0, 0.5, 0.05], "EXPDECAY": [2.0, -1.3],
         "POLY": [1.0, -2.0, 0.5, 3.0]}[name]
truth = np.array(truth)
with native.Context(1) as ctx:
    ds = native.Dataset(ctx, m, 1).generate(fam, truth)
    p = truth * 1.1 + 0.01
    for _ in range(3):
        ds.normal_eq(fam, p)
    k2 = ctx.stats().kernel_ms
    for _ in range(3):
        R = ds.qr(fam, p)
    qr = ctx.stats().kernel_ms
    JtJ, Jtr, rtr = ds.normal_eq(fam, p)
    n = truth.size
    G = np.block([[JtJ, Jtr[:, None]], [Jtr[None, :], np.array([[rtr]])]])
    err = np.max(np.abs(R.T @ R - G) / np.sqrt(np.outer(np.diag(G), np.diag(G))))
    print(f"{name} rows={m}: K2 {k2:.3f} ms, TSQR {qr:.3f} ms ({16e-6 * m / qr:.0f} GB/s), |R^T R - G| rel {err:.2e}")
    ds.free()
