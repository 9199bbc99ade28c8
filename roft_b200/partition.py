"""Multi-GPU host logic: independent tracks are statically partitioned across ranks.

A frame does not shard; there is no data-path collective.  Track t of T goes to rank floor(t*G/T)
(contiguous blocks), per-track synthetic data is seeded by the global track id so it does not
depend on the partition, and the only cross-rank operations are a barrier and a MAX over the
per-rank elapsed times (both via torch.distributed, NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Tuple


def track_range(rank: int, world: int, total_tracks: int) -> Tuple[int, int]:
    """[first, last) global track ids owned by `rank` (block partition, sizes differ by at most one)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    lo = (rank * total_tracks) // world
    hi = ((rank + 1) * total_tracks) // world
    return lo, hi


def owner_of(track: int, world: int, total_tracks: int) -> int:
    """Rank owning a global track id (inverse of track_range)."""
    if not (0 <= track < total_tracks):
        raise ValueError("track out of range")
    r = (track * world) // total_tracks
    while track_range(r, world, total_tracks)[1] <= track:
        r += 1
    while track_range(r, world, total_tracks)[0] > track:
        r -= 1
    return r


def max_over_ranks_ms(local_ms: float, device=None) -> float:
    """Elapsed time of a step region = MAX over ranks (never wall clock of one rank)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(local_ms)
    t = torch.tensor([float(local_ms)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
