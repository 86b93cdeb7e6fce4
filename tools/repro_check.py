"""bitwise run-to-run reproducibility at full size: python tools/repro_check.py [rows] [reps]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scalaopt_b200 import native
m = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
truth = np.array([2.0, 0.5, 0.05, 1.0, 3.0])
fam = native.GAUSS_EXPBKG
with native.Context(1) as ctx:
    ds = native.Dataset(ctx, m, 1).generate(fam, truth)
    p = truth * 1.1 + 0.01
    for name, fn in (("qr", lambda: ds.qr(fam, p)), ("normal_eq", lambda: np.concatenate([np.ravel(a) for a in ds.normal_eq(fam, p)])),
                     ("loss_grad", lambda: np.concatenate([np.ravel(a) for a in ds.loss_grad(fam, p)]))):
        outs = [np.array(fn(), dtype=np.float64).copy() for _ in range(reps)]
        same = [bool(np.array_equal(outs[0].view(np.uint64), o.view(np.uint64))) for o in outs]
        dev = max(float(np.max(np.abs(o - outs[0]) / (np.abs(outs[0]) + 1e-300))) for o in outs)
        print(name, "rows", m, "bitwise_equal", same, "max_rel_dev", dev, flush=True)
    ds.free()
