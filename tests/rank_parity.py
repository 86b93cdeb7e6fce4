"""Rank-mode parity: one process per GPU (``so_init_rank``, what ``bench.py`` runs under torchrun) against the CPU oracle.

Every rank generates its shard of ONE global synthetic data set on its GPU, evaluates K1-K4 (+ TSQR) through the C ABI
(the cross-GPU sum happens inside the library), reads its own shard back, computes the oracle's sums over that shard on
the host, and the per-rank oracle sums are added with ``torch.distributed`` (fp64).  Checked: the tolerances of
BASELINE.json's north_star (loss / gradient 1e-12, J^T J 1e-11) and that all ranks hold bit-identical results.

Used by ``tests/test_gpu_rank.py`` (spawns torchrun on >= 2 GPUs) and by ``bench.py`` before its timed region.
Run by hand:  torchrun --nproc-per-node 2 --master-addr 127.0.0.1 tests/rank_parity.py
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

TRUTH = np.array([2.0, 0.5, 0.05, 1.0, 3.0])


def _rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def _gram_rel(G, Gref):
    dg = np.sqrt(np.abs(np.diag(Gref)))
    scale = np.outer(dg, dg)
    scale[scale == 0] = 1.0
    return float(max((np.abs(G - Gref) / scale).max(), np.linalg.norm(G - Gref) / np.linalg.norm(Gref)))


def check(native, ctx, rank: int, world: int, rows_per_rank: int, allsum, allgather, family=None, f: int = 1,
          truth=None, noise=0.05, seed=12345):
    """-> dict of relative errors (and 'bitwise_equal_across_ranks').  allsum(np.ndarray) -> element-wise sum over ranks;
    allgather(np.ndarray) -> list of every rank's array."""
    import oracle
    oracle.build()
    fam = oracle.GAUSS_EXPBKG if family is None else family
    truth = TRUTH if truth is None else np.asarray(truth, float)
    n = truth.size
    w = rows_per_rank
    ds = native.Dataset(ctx, w, f).generate(fam, truth, seed=seed, noise_sigma=noise, row_offset=rank * w)
    m = ds.size()
    assert m == w * world, (m, w, world)
    X, y = ds.read(0, w)
    p = truth * 1.1
    out = {}
    # --- GPU, through the C ABI (all ranks call collectively) ---
    loss, grad = ds.loss_grad(fam, p)
    fin1 = int(ctx.stats().finished_on_device)
    JtJ, Jtr, rtr = ds.normal_eq(fam, p)
    d = -grad
    alphas = [0.0, 0.5, 1.0, 2.0]
    phi, dphi = ds.line(fam, p, d, alphas)
    Hd = ds.hess_vec(fam, p, d)
    R = None
    try:
        R = ds.qr(fam, p)
    except native.NativeError as e:
        if e.code != -5:      # SO_EUNSUPPORTED: outside the TSQR catalogue
            raise
    # --- oracle on this rank's shard, summed over ranks ---
    wgt = w / m
    lo = allsum(np.array([oracle.apply(fam, X, y, p) * wgt]))[0]
    go = allsum(oracle.gradient(fam, X, y, p) * wgt)
    J0, r0, t0 = oracle.normal_eq(fam, X, y, p)
    flat = allsum(np.concatenate([J0.ravel(), r0, [t0]]))
    JtJo, Jtro, rtro = flat[: n * n].reshape(n, n), flat[n * n: n * n + n], flat[-1]
    po, dpo = oracle.line(fam, X, y, p, d, alphas)
    pd = allsum(np.concatenate([po, dpo]) * wgt)
    Hdo = allsum(oracle.hess_vec(fam, X, y, p, d) * wgt)
    out["loss"] = _rel(loss, lo)
    out["grad"] = _rel(grad, go)
    out["JtJ"] = _gram_rel(JtJ, JtJo)
    out["Jtr"] = _rel(Jtr, Jtro)
    out["rtr"] = _rel(rtr, rtro)
    out["phi"] = _rel(phi, pd[: len(alphas)])
    out["dphi"] = _rel(dphi, pd[len(alphas):])
    out["Hd"] = _rel(Hd, Hdo)
    if R is not None:
        Gm = np.block([[JtJo, Jtro[:, None]], [Jtro[None, :], np.array([[rtro]])]])
        out["tsqr_RtR"] = _gram_rel(R.T @ R, Gm)
    mine = np.concatenate([[loss], grad, JtJ.ravel(), Jtr, [rtr], phi, dphi, Hd])
    everyone = allgather(mine)
    out["bitwise_equal_across_ranks"] = bool(all(np.array_equal(everyone[0], e) for e in everyone))
    out["finished_on_device"] = fin1
    out["rows_per_rank"] = int(w)
    out["world"] = int(world)
    ds.free()
    ok = (out["loss"] <= 1e-12 and out["grad"] <= 1e-12 and out["JtJ"] <= 1e-11 and out["Jtr"] <= 1e-11 and
          out["rtr"] <= 1e-12 and out["phi"] <= 1e-12 and out["dphi"] <= 1e-11 and out["Hd"] <= 1e-11 and
          out.get("tsqr_RtR", 0.0) <= 1e-11 and out["bitwise_equal_across_ranks"])
    out["ok"] = bool(ok)
    return out


def torch_collectives(dist, world):
    import torch

    def allsum(a):
        if world == 1:
            return np.asarray(a, float)
        t = torch.tensor(np.asarray(a, float), dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        return t.cpu().numpy()

    def allgather(a):
        if world == 1:
            return [np.asarray(a, float)]
        t = torch.tensor(np.asarray(a, float), dtype=torch.float64, device="cuda")
        outs = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(outs, t)
        return [o.cpu().numpy() for o in outs]

    return allsum, allgather


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    rows = int(os.environ.get("RANK_PARITY_ROWS", "100003"))
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from scalaopt_b200 import native
    native.load()
    uid = None
    if world > 1:
        box = [native.Context.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        uid = box[0]
    ctx = native.Context(rank=rank, world=world, device_id=local_rank, nccl_uid=uid)
    allsum, allgather = torch_collectives(dist, world)
    res = {"gauss_expbkg": check(native, ctx, rank, world, rows, allsum, allgather)}
    import oracle
    wt = np.array([(((j % 7) - 3) / 4.0) for j in range(9)])
    res["logistic_f8"] = check(native, ctx, rank, world, rows // 4 + 1, allsum, allgather, family=oracle.LOGISTIC, f=8,
                               truth=wt, noise=1.0)
    wl = np.array([0.5 + 0.1 * j for j in range(5)])
    res["linear_f4"] = check(native, ctx, rank, world, rows // 2, allsum, allgather, family=oracle.LINEAR, f=4,
                             truth=wl, noise=1.0)
    ctx.close()
    if rank == 0:
        print("RANK_PARITY " + json.dumps(res), flush=True)
    if world > 1:
        dist.destroy_process_group()
    if not all(v["ok"] for v in res.values()):
        sys.exit(3)


if __name__ == "__main__":
    main()
