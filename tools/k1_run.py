import os, sys
import numpy as np
sys.path.insert(0, "/root/repo")
from scalaopt_b200 import native
f, m = 31, 40_000_000
truth = np.arange(1, f + 2) / (f + 1)
with native.Context(1) as ctx:
    ds = native.Dataset(ctx, m, f).generate(native.LINEAR, truth)
    for _ in range(3):
        ds.loss_grad(native.LINEAR, truth * 0.9)
    print(ctx.stats().kernel_ms)
