"""Host-side partitioning for multi-GPU runs (one process per GPU).

* many flights (configs 3, 5): contiguous blocks of flights per rank, no communication;
* one very large grid (config 4): each rank OWNS a band of rows and replays, in log order, every frame
  whose rays can reach it -- cells have exactly one owner, so the bands are exact and are only gathered
  (one all-gather over NCCL), never summed.
"""
from __future__ import annotations

from typing import Tuple


def flight_shard(n_flights: int, rank: int, world: int) -> Tuple[int, int]:
    """[first, count) of the flights rank ``rank`` replays; blocks differ by at most one flight."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(n_flights, world)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def row_band(H: int, rank: int, world: int, align: int = 4) -> Tuple[int, int]:
    """[row0, rows) of the grid rows rank ``rank`` owns; band edges are multiples of ``align``."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    units = (H + align - 1) // align
    base, extra = divmod(units, world)
    u0 = rank * base + min(rank, extra)
    u1 = u0 + base + (1 if rank < extra else 0)
    r0, r1 = min(u0 * align, H), min(u1 * align, H)
    return r0, r1 - r0
