import sys, numpy as np
sys.path.insert(0, "/root/repo")
from scalaopt_b200 import native
with native.Context(1) as ctx:
    m = 1_000_000_000
    ds = native.Dataset(ctx, m, 1)
    rng = np.random.default_rng(0)
    blk = 10_000_000
    x = np.where(rng.random(blk) < 0.3, rng.normal(125.0, 10.0, blk), 50.0 - np.log(rng.random(blk)) / 0.05)
    for _ in range(m // blk): ds.append(x.reshape(-1, 1), np.zeros(blk))
    for _ in range(4): v = ds.nll(native.PDF_GAUSS_EXPBKG, [0.3, 125.0, 10.0, 0.05], [50.0])
    print("nll", v, "ms", ctx.stats().kernel_ms, "GB/s", 8e-6 * m / ctx.stats().kernel_ms)
