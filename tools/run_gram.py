"""tiny driver for profiling the wide Gram kernels: python tools/run_gram.py <f> <rows> [logistic]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scalaopt_b200 import native
f, m = int(sys.argv[1]), int(sys.argv[2])
fam = native.LOGISTIC if len(sys.argv) > 3 else native.LINEAR_NOINT
n = f + (1 if fam == native.LOGISTIC else 0)
truth = np.arange(1, n + 1) / n
with native.Context(1) as ctx:
    ds = native.Dataset(ctx, m, f).generate(fam, truth)
    for _ in range(4):
        ds.normal_eq(fam, truth * 0.9)
    print(ctx.stats().kernel_ms)
