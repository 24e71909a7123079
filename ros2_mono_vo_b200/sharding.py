"""Stream-group sharding across GPUs: replicas only, no collective on the data path.

A VO stream is sequential, streams are independent: stream s lives on rank `s mod world` for its whole life.
torch.distributed is used only for the start/stop barrier and the max-over-ranks of the timed region.
"""
from __future__ import annotations

from typing import List


def streams_for_rank(n_streams: int, rank: int, world: int) -> List[int]:
    """Round-robin ownership: global stream ids served by `rank`."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank / world")
    return list(range(rank, n_streams, world))


def weak_scaling_streams(streams_per_gpu: int, rank: int, world: int) -> List[int]:
    """Weak scaling (bench.py): every rank gets `streams_per_gpu` streams; ids are disjoint across ranks and
    equal the round-robin layout of `streams_per_gpu * world` streams."""
    return streams_for_rank(streams_per_gpu * world, rank, world)


def max_over_ranks(value_ms: float, device=None) -> float:
    """Timing reduction: the job is as slow as its slowest rank."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value_ms)
    t = torch.tensor([value_ms], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def total_over_ranks(count: int, device=None) -> int:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return int(count)
    t = torch.tensor([count], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return int(t.item())
