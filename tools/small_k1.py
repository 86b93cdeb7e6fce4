import os, sys
import numpy as np
sys.path.insert(0, "/root/repo")
from scalaopt_b200 import native
n, m = 128, 1_000_000
truth = [(j + 1) / n for j in range(n)]
with native.Context(1) as ctx:
    ds = native.Dataset(ctx, m, n).generate(native.LINEAR_NOINT, truth, noise_sigma=1.0)
    p = np.asarray(truth) * 0.95 + 0.01
    ts = []
    for i in range(30):
        ds.loss_grad(native.LINEAR_NOINT, p); ts.append(round(ctx.stats().kernel_ms, 4))
    print("K1 n=128 m=1e6:", ts)
    ts = []
    for i in range(10):
        ds.loss_grad(native.LINEAR_NOINT, p, want_grad=False); ts.append(round(ctx.stats().kernel_ms, 4))
    print("K1loss:", ts)
    ds.free()
