#!/usr/bin/env python
"""tools/cfg5.py — BASELINE configs[4]: data-set evaluation sweep m in {1e6..4e9} x n in {4, 32, 128} on G GPUs in rank
mode (one process per GPU; run under torchrun for G > 1), K1 (loss + gradient) and K2 (loss + J^T r + J^T J).

Per point: the wall time of one BLOCKING C-ABI call (what the optimizer waits for), the CUDA-event kernel time, the time
the folding block spent in the cross-GPU exchange, and what is left (launch + host).  m is the TOTAL number of
observations; each rank holds m/G of them.  Points whose shard does not fit 150 GB of HBM are skipped and say so.
Family per SURVEY.md 8d cfg5: linear with intercept, f = n - 1; the no-intercept f = n variant (16-byte aligned rows)
is measured beside it.  With --cpu (G = 1) the oracle's single-pass evaluation is timed on the host cores beside it
(seq = 1 core, par[N] = all cores) on a bounded sample of the same rows.

  python tools/cfg5.py --cpu > profiles/r2_cfg5_g1.jsonl
  torchrun --nproc-per-node 8 --master-addr 127.0.0.1 tools/cfg5.py > profiles/r2_cfg5_g8.jsonl
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

HBM = 6554.2
try:
    HBM = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
FP64_TF = 37.16      # profiles/r1_fp64_peak.json (DMMA m8n8k4)
_dp = C.POINTER(C.c_double)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cpu", action="store_true")
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--ms", default="1e6,1e7,1e8,1e9,4e9")
    ap.add_argument("--ns", default="4,32,128")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from scalaopt_b200 import native
    L = native.load()
    uid = None
    if world > 1:
        box = [native.Context.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        uid = box[0]
    ctx = native.Context(rank=rank, world=world, device_id=local_rank, nccl_uid=uid)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxr(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def emit(**kw):
        if rank == 0:
            print(json.dumps(kw), flush=True)

    for n in [int(x) for x in args.ns.split(",")]:
        for m_total in [int(float(x)) for x in args.ms.split(",")]:
            for icpt in (True, False):
                fam = native.LINEAR if icpt else native.LINEAR_NOINT
                f = n - 1 if icpt else n
                rows = m_total // world
                name = f"cfg5 linear{'+icpt' if icpt else ''} n={n} f={f} m={m_total:.0e} G={world}"
                gbytes = rows * 8.0 * (f + 1) / 1e9
                if gbytes > 150.0:
                    emit(config=name, skipped=f"{gbytes:.0f} GB per GPU exceeds HBM")
                    continue
                truth = np.array([(j + 1) / n for j in range(n)])
                try:
                    ds = native.Dataset(ctx, rows, f).generate(fam, truth, noise_sigma=1.0, row_offset=rank * rows)
                except native.NativeError as e:
                    emit(config=name, skipped=f"{e}")
                    continue
                p = np.ascontiguousarray(truth * 0.95 + 0.01)
                loss, grad = C.c_double(), np.empty(n)
                JtJ, Jtr, rtr = np.empty((n, n)), np.empty(n), C.c_double()
                calls = {
                    "K1": lambda: L.so_eval_loss_grad(ds._h, fam, p.ctypes.data_as(_dp), n, C.byref(loss), grad.ctypes.data_as(_dp)),
                    "K2": lambda: L.so_eval_normal_eq(ds._h, fam, p.ctypes.data_as(_dp), n, JtJ.ctypes.data_as(_dp),
                                                      Jtr.ctypes.data_as(_dp), C.byref(rtr)),
                }
                for kind, fn in calls.items():
                    for _ in range(3):
                        ctx.check(fn())
                    ctx.stats_reset()
                    barrier()
                    t0 = time.perf_counter()
                    for _ in range(args.reps):
                        rc = fn()
                    barrier()
                    wall = maxr(time.perf_counter() - t0) / args.reps * 1e3
                    ctx.check(rc)
                    st = ctx.stats()
                    kms, fms = maxr(st.sum_kernel_ms) / args.reps, maxr(st.sum_finish_ms) / args.reps
                    rec = dict(config=name, kernel=kind, m=m_total, n=n, f=f, G=world, rows_per_gpu=rows,
                               wall_ms_per_call=wall, kernel_ms=kms, kernel_ms_best=st.min_kernel_ms, exchange_ms=fms,
                               host_and_launch_ms=wall - kms, rows_per_s=m_total / (wall * 1e-3),
                               gbs_per_gpu=gbytes / (kms * 1e-3), hbm_frac=gbytes / (kms * 1e-3) / HBM,
                               finished_on_device=int(st.finished_on_device),
                               regime="latency-bound (per-call fixed cost comparable to the kernel)" if wall > 1.5 * (gbytes / HBM * 1e3) and kms < 0.2 else "throughput")
                    if kind == "K2":
                        tf = rows * (n + 1.0) * (n + 2.0) / (kms * 1e-3) / 1e12
                        rec.update(tflops_per_gpu=tf, fp64_frac=tf / FP64_TF)
                    emit(**rec)
                if args.cpu and rank == 0 and world == 1 and icpt and m_total <= 10_000_000:
                    import oracle
                    oracle.build()
                    threads = max(1, len(os.sched_getaffinity(0)))
                    mc = int(min(m_total, max(20_000, 4.0e9 / (n * n + 16 * n))))      # bounded sample, a few seconds per leg
                    X, y = ds.read(0, mc)
                    for kind, fn in (("K1", lambda t: oracle.gradient(fam, X, y, p, parts=t, threads=t)),
                                     ("K2", lambda t: oracle.normal_eq(fam, X, y, p, parts=t, threads=t))):
                        for label, t in (("seq", 1), (f"par[{threads}]", threads)):
                            fn(t)
                            t0 = time.perf_counter()
                            fn(t)
                            dt = time.perf_counter() - t0
                            emit(config=name, kernel=kind, cpu=label, cores=t, sample_rows=mc, rows_per_s=mc / dt,
                                 note="oracle C port, single-pass analytic evaluation; the JVM/Spark reference cannot run here")
                ds.free()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
