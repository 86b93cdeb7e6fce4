import sys, numpy as np
sys.path.insert(0, "/root/repo")
from scalaopt_b200 import native
truth = np.array([2.0, 0.5, 0.05, 1.0, 3.0]); fam = native.GAUSS_EXPBKG
with native.Context(1) as ctx:
    ds = native.Dataset(ctx, 1_000_000_000, 1).generate(fam, truth, noise_sigma=0.05)
    p = truth * 0.95 + 0.01; d = np.cos(np.arange(5) + 1.0) * 0.01
    for k in (1, 2, 3, 4, 5, 6, 8):
        al = [0.0, 1.0, 2.0, 4.0, 8.0, 0.5, 0.25, 0.125][:k]
        ts = []
        for _ in range(5): ds.line(fam, p, d, al); ts.append(ctx.stats().kernel_ms)
        print("K3 k=%d: %.3f ms" % (k, sorted(ts)[2]))
    ds.free()
