"""CPU tests of the N > 1 path's host logic with world_size-2/3 gloo: shard plan + SUM all-reduce of the partial
buffers reproduces the single-process evaluation (the GPU side does the same all-reduce with NCCL)."""
import os
import socket

import numpy as np
import pytest


def test_shard_bounds_cover_rows_once():
    from scalaopt_b200.sharding import shard_bounds
    for m in (0, 1, 2, 7, 1000, 1001, 10**9 + 7):
        for parts in (1, 2, 3, 4, 8):
            for f in (1, 32):
                b = shard_bounds(m, parts, f)
                assert b[0][0] == 0 and b[-1][1] == m
                assert all(b[i][1] == b[i + 1][0] for i in range(parts - 1))
                if f == 1:
                    assert all(lo % 2 == 0 for lo, hi in b if hi > lo)
    assert shard_bounds(10, 4, 32) == [(0, 3), (3, 6), (6, 9), (9, 10)]
    assert shard_bounds(10, 4, 1) == [(0, 4), (4, 8), (8, 10), (10, 10)]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, m, out_dir):
    import torch.distributed as dist

    import oracle
    from scalaopt_b200.sharding import allreduce_partials, partial_buffer_doubles, shard_bounds

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    fam = oracle.GAUSS_EXPBKG
    truth = np.array([2.0, 0.5, 0.05, 1.0, 3.0])
    p = truth * 1.1
    lo, hi = shard_bounds(m, world, 1)[rank]
    # every rank generates ITS shard of the one global data set (counter-based generator, row_offset = lo)
    X, y = oracle.generate(fam, hi - lo, 1, truth, noise_sigma=0.05, row_offset=lo, total_rows=m)
    n = 5
    if hi > lo:
        JtJ, Jtr, rtr = oracle.normal_eq(fam, X, y, p)
        g = oracle.gradient(fam, X, y, p) * (hi - lo)
        loss = oracle.apply(fam, X, y, p) * (hi - lo)
    else:
        JtJ, Jtr, rtr, g, loss = np.zeros((n, n)), np.zeros(n), 0.0, np.zeros(n), 0.0
    iu = np.triu_indices(n)
    buf = np.concatenate([JtJ[iu], Jtr, [rtr]])
    assert buf.size == partial_buffer_doubles("normal_eq", n)
    tot = allreduce_partials(buf)
    lg = allreduce_partials(np.concatenate([[loss], g]))
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
