"""CPU tests of the N > 1 path's host logic (world_size 2/3, gloo): the product's shard plan (`so_shard_plan`, called
through the C ABI — the same function so_dataset_create uses) splits the rows, every rank evaluates its shard, and the
SUM of the per-rank buffers equals the evaluation of the whole data set.  The GPU side of the same thing — so_init_rank,
the in-kernel sum over peer memory or NCCL — is tests/test_gpu_rank.py, which needs GPUs."""
import os
import socket

import numpy as np
import pytest


def test_shard_plan_covers_rows_once(native):
    for m in (0, 1, 2, 7, 1000, 1001, 10**9 + 7):
        for parts in (1, 2, 3, 4, 8):
            for f in (1, 32):
                b = [native.shard_plan(m, f, parts, g) for g in range(parts)]
                assert b[0][0] == 0 and b[-1][0] + b[-1][1] == m
                assert all(b[i][0] + b[i][1] == b[i + 1][0] for i in range(parts - 1))
                if f == 1:
                    assert all(lo % 2 == 0 for lo, cnt in b if cnt > 0)     # 16-byte aligned shard starts
    assert [native.shard_plan(10, 32, 4, g) for g in range(4)] == [(0, 3), (3, 3), (6, 3), (9, 1)]
    assert [native.shard_plan(10, 1, 4, g) for g in range(4)] == [(0, 4), (4, 4), (8, 2), (10, 0)]
    with pytest.raises(native.NativeError):
        native.shard_plan(10, 1, 4, 4)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, m, out_dir):
    import torch
    import torch.distributed as dist

    import oracle
    from scalaopt_b200 import native

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    def allsum(a):
        t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64).copy())
        dist.all_reduce(t)
        return t.numpy()

    fam = oracle.GAUSS_EXPBKG
    truth = np.array([2.0, 0.5, 0.05, 1.0, 3.0])
    p = truth * 1.1
    lo, cnt = native.shard_plan(m, 1, world, rank)
    # every rank generates ITS shard of the one global data set (counter-based generator, row_offset = lo)
    X, y = oracle.generate(fam, cnt, 1, truth, noise_sigma=0.05, row_offset=lo, total_rows=m)
    n = 5
    if cnt > 0:
        JtJ, Jtr, rtr = oracle.normal_eq(fam, X, y, p)
        g = oracle.gradient(fam, X, y, p) * cnt
        loss = oracle.apply(fam, X, y, p) * cnt
    else:
        JtJ, Jtr, rtr, g, loss = np.zeros((n, n)), np.zeros(n), 0.0, np.zeros(n), 0.0
    iu = np.triu_indices(n)
    tot = allsum(np.concatenate([JtJ[iu], Jtr, [rtr]]))
    lg = allsum(np.concatenate([[loss], g]))
    rows = allsum(np.array([float(cnt)]))
    assert rows[0] == m
    if rank == 0:
        np.save(os.path.join(out_dir, "normal_eq.npy"), tot)
        np.save(os.path.join(out_dir, "loss_grad.npy"), lg / m)
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_evaluation_with_gloo_allreduce_matches_whole(oracle, tmp_path, world):
    import torch.multiprocessing as mp
    m = 20_001
    port = _free_port()
    mp.spawn(_worker, args=(world, port, m, str(tmp_path)), nprocs=world, join=True)
    fam = oracle.GAUSS_EXPBKG
    truth = np.array([2.0, 0.5, 0.05, 1.0, 3.0])
    p = truth * 1.1
    X, y = oracle.generate(fam, m, 1, truth, noise_sigma=0.05)
    JtJ, Jtr, rtr = oracle.normal_eq(fam, X, y, p)
    tot = np.load(tmp_path / "normal_eq.npy")
    ref = np.concatenate([JtJ[np.triu_indices(5)], Jtr, [rtr]])
    np.testing.assert_allclose(tot, ref, rtol=1e-12, atol=1e-12 * np.abs(ref).max())   # sums with cancellation
    lg = np.load(tmp_path / "loss_grad.npy")
    assert lg[0] == pytest.approx(oracle.apply(fam, X, y, p), rel=1e-12)
    gref = oracle.gradient(fam, X, y, p)
    np.testing.assert_allclose(lg[1:], gref, rtol=1e-11, atol=1e-12 * np.abs(gref).max())
