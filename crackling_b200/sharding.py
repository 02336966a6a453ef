"""Guide partitioning across GPUs / ranks.

Guides are independent (ref isslScoreOfftargets.cpp:316-317), so the multi-GPU path is: replicate
the index, give every rank a contiguous range of the guide file, write results back in input order.
There is no collective on the data path; torch.distributed is only used to time (max over ranks)
and, in tests, to gather the per-rank score ranges to rank 0.
"""
from __future__ import annotations

import numpy as np


def shard_bounds(n: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous range of rank `rank` out of `world` (same split as the host program's threads)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad world/rank")
    return n * rank // world, n * (rank + 1) // world


def max_over_ranks(value: float, dist=None, device=None) -> float:
    """Step time of the job = the slowest rank's."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_in_order(local: np.ndarray, n_total: int, dist=None) -> np.ndarray | None:
    """Concatenates every rank's contiguous score range on rank 0 (None elsewhere)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [shard_bounds(n_total, world, r)[1] - shard_bounds(n_total, world, r)[0] for r in range(world)]
    pad = max(sizes)
    buf = torch.zeros(pad, dtype=torch.float64)
    buf[:local.size] = torch.from_numpy(np.ascontiguousarray(local, dtype=np.float64))
    out = [torch.zeros(pad, dtype=torch.float64) for _ in range(world)] if rank == 0 else None
    dist.gather(buf, out, dst=0)
    if rank != 0:
        return None
    return np.concatenate([o.numpy()[:s] for o, s in zip(out, sizes)])
