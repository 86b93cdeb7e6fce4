"""Row sharding of a data set over GPUs / ranks — the host-side plan that ``so_dataset_create`` applies
(``scalaopt_b200/csrc/so_api.cu``) and that a multi-process launch (one rank per GPU) follows.

Every quantity of the hot path is a sum over independent rows (which is why the reference can express it as
``DataSet.aggregate``, DataSet.scala:49), so the exchange step is one SUM all-reduce of the small partial
buffer: (1+n) doubles for loss+gradient, n(n+1)/2+n+1 for the normal equations, 2k for a line search, n for
a Hessian-vector product.
"""
from __future__ import annotations


def shard_bounds(m: int, parts: int, f: int = 0) -> list[tuple[int, int]]:
    """Contiguous row blocks [begin, end) of `parts` shards over m rows: ceil(m/parts) rows each, rounded up to an
    even count for single-input data sets so that every shard starts 16-byte aligned."""
    if parts < 1 or m < 0:
        raise ValueError("parts >= 1 and m >= 0 required")
    mg = (m + parts - 1) // parts
    if f == 1 and (mg & 1) and parts > 1:
        mg += 1
    return [(min(g * mg, m), min((g + 1) * mg, m)) for g in range(parts)]


def partial_buffer_doubles(kind: str, n: int, k: int = 1) -> int:
    """Size of the buffer one evaluation all-reduces."""
    if kind == "loss_grad":
        return 1 + n
    if kind == "normal_eq":
        return n * (n + 1) // 2 + n + 1
    if kind == "line":
        return 2 * k
    if kind == "hess_vec":
        return n
    raise ValueError(kind)


def allreduce_partials(local, group=None):
    """SUM all-reduce of a rank's partial buffer with torch.distributed (gloo on CPU in the tests, NCCL inside
    libscalaopt_cuda.so on the GPUs).  Returns a new array."""
    import numpy as np
    import torch
    import torch.distributed as dist

    t = torch.from_numpy(np.ascontiguousarray(local, dtype=np.float64).copy())
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.numpy()
