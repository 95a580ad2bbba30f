"""Scenario sharding across GPUs: scenarios are independent, each rank owns a
contiguous block, there is no data-path collective; torch.distributed is used only for the
barrier, the max-over-ranks of device times and gathering small per-rank metrics."""
from __future__ import annotations


def shard_range(total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block [lo, hi) of `total` scenarios owned by `rank` (sizes differ by <= 1)."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def max_over_ranks(values, device=None):
    """Element-wise max of a small list of floats over all ranks (no-op without a process group)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t]


def sum_over_ranks(values, device=None):
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(v) for v in t]


def balanced_shards(costs, world: int):
    """Load-balanced assignment for MIXED scenario sizes (n = 5..15, N = 4..10):
    sort by the cost proxy (7 n N binaries per scenario) and deal round-robin.  Returns one index list per rank."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    return [sorted(order[r::world]) for r in range(world)]
