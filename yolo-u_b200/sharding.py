"""Batch sharding for the multi-GPU configuration (BASELINE cfg 3): slices are independent, volumes are
kept whole, each rank takes a contiguous run of volumes.  No data-path collective -- only SegMetrics.reduce()."""
from __future__ import annotations

from typing import List, Tuple


def shard_volumes(n_volumes: int, world_size: int, rank: int) -> Tuple[int, int]:
    """[lo, hi) volume range of `rank`; remainders go to the lowest ranks so sizes differ by at most 1."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of {world_size}")
    base, rem = divmod(n_volumes, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_slices(n_volumes: int, slices_per_volume: int, world_size: int, rank: int) -> Tuple[int, int]:
    lo, hi = shard_volumes(n_volumes, world_size, rank)
    return lo * slices_per_volume, hi * slices_per_volume


def batches(lo: int, hi: int, batch: int) -> List[Tuple[int, int]]:
    return [(s, min(s + batch, hi)) for s in range(lo, hi, batch)]
