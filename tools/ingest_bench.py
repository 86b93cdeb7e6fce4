"""tools/ingest_bench.py — ingest throughput (GB/s) into HBM: pinned host memory, pageable host memory (staged),
raw fp64 files (page cache).  One JSON line."""
import json, os, sys, tempfile, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scalaopt_b200 import native

m, f = 100_000_000, 1            # 1.6 GB
with native.Context(1) as ctx:
    ds = native.Dataset(ctx, m, f)
    X = np.random.default_rng(1).random((m, f)); y = np.random.default_rng(2).random(m)
    gb = (X.nbytes + y.nbytes) / 1e9
    out = {"rows": m, "f": f, "gigabytes": gb}
    hx, hy = native.PinnedBuffer(m * f), native.PinnedBuffer(m)
    hx.array[:] = X.reshape(-1); hy.array[:] = y
    for name, fn in (("pinned", lambda: ds.append_raw(hx.array.ctypes.data, hy.array.ctypes.data, m)),
                     ("pageable_staged", lambda: ds.append(X, y))):
        best = 1e9
        for _ in range(3):
            ds.clear(); t0 = time.perf_counter(); fn(); best = min(best, time.perf_counter() - t0)
        out[name + "_gbs"] = gb / best
    with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as d:
        px, py = os.path.join(d, "X.f64"), os.path.join(d, "y.f64")
        X.tofile(px); y.tofile(py)
        best = 1e9
        for _ in range(3):
            ds.clear(); t0 = time.perf_counter(); ds.load_raw(px, py, m); best = min(best, time.perf_counter() - t0)
        out["raw_file_gbs"] = gb / best
    print(json.dumps(out))
