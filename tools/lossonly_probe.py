import sys, os, numpy as np
sys.path.insert(0, os.getcwd())
from scalaopt_b200 import native
with native.Context(1) as ctx:
    f = 32; fam = native.LOGISTIC
    truth = np.array([(((j % 7) - 3) / 4.0) for j in range(f + 1)])
    ds = native.Dataset(ctx, 100_000_000, f).generate(fam, truth)
    p = truth * 0.95 + 0.01
    for _ in range(3): ds.loss_grad(fam, p)
    a, b = [], []
    for i in range(12):
        ds.loss_grad(fam, p); a.append(ctx.stats().kernel_ms)
        ds.loss_grad(fam, p, want_grad=False); b.append(ctx.stats().kernel_ms)
    print("alternating: grad", np.round(a, 3), "loss-only", np.round(b, 3))
    a = []
    for i in range(8): ds.loss_grad(fam, p, want_grad=False); a.append(ctx.stats().kernel_ms)
    print("loss-only x8", np.round(a, 3))
    a = []
    for i in range(8): ds.loss_grad(fam, p); a.append(ctx.stats().kernel_ms)
    print("grad x8", np.round(a, 3))
    st = ctx.stats(); print(st)
