"""Cross-rank bookkeeping of the benchmark / multi-process runs (torch.distributed is plumbing only: the DP path
needs no data-path collective -- every rank scores its own shard of the problem list)."""
from __future__ import annotations

import numpy as np


def shard_problems(cells: np.ndarray, world: int, rank: int) -> np.ndarray:
    """Indices of the problems `rank` handles: problems sorted by size (largest first, stable) and dealt round-robin, so
    every rank gets the same mix of sizes.  (The executable, csrc/host/main.cpp, deals its alignment passes round-robin within
    1/4-octave size classes instead of a full sort (my_shard: same balance, O(n)) and hands phase 1 out dynamically in runs of
    the same largest-first order, because its GPU processes come up one after the other.)"""
    order = np.argsort(-cells.astype(np.int64), kind="stable")
    return order[rank::world]


def reduce_over_ranks(times_ms, work, device="cpu"):
    """max over ranks of every time, sum over ranks of every work counter; returns two python lists."""
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(times_ms), dtype=torch.float64, device=device)
    w = torch.tensor(list(work), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(w, op=dist.ReduceOp.SUM)
    return t.tolist(), w.tolist()
