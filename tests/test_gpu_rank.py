"""Rank mode (one process per GPU, ``so_init_rank`` — what bench.py and the 1/2/4/8-GPU scaling runs use) against the
oracle: tests/rank_parity.py under torchrun with as many ranks as there are GPUs (2, 4, 8), and as a single rank."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    import torch
    return torch.cuda.device_count()


def _run(world, extra_env=None):
    env = dict(os.environ, RANK_PARITY_ROWS="60001")
    env.update(extra_env or {})
    script = os.path.join(ROOT, "tests", "rank_parity.py")
    if world == 1:
        cmd = [sys.executable, script]
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
               "--master-addr", "127.0.0.1", "--master-port", str(29600 + world), script]
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    lines = [l for l in proc.stdout.splitlines() if l.startswith("RANK_PARITY ")]
    assert proc.returncode == 0 and lines, f"rc={proc.returncode}\n{proc.stdout[-3000:]}\n{proc.stderr[-3000:]}"
    return json.loads(lines[-1][len("RANK_PARITY "):])


def test_single_rank_matches_oracle():
    res = _run(1)
    for fam, r in res.items():
        assert r["ok"], (fam, r)
        assert r["finished_on_device"] == 1, (fam, r)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_ranks_match_oracle(world):
    if _ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    res = _run(world)
    for fam, r in res.items():
        assert r["ok"], (fam, r)
        assert r["bitwise_equal_across_ranks"], (fam, r)
        assert r["finished_on_device"] == 1, (fam, r)       # the sum was done by the kernels over peer memory, not by NCCL


def test_ranks_match_oracle_through_nccl():
    """the NCCL all-reduce route (results too large for the mailboxes take it; forced here for everything)"""
    if _ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    res = _run(2, {"SCALAOPT_NO_PEER_FINISH": "1"})
    for fam, r in res.items():
        assert r["ok"], (fam, r)
        assert r["finished_on_device"] == 0, (fam, r)
