"""Host-side partitioning for multi-GPU runs: thin wrappers over the C functions
``uqs_flight_shard`` / ``uqs_row_band`` of libuqs_mapping.so (csrc/uqs_multi.cu), kept so that Python harnesses
and the C library can never disagree on who owns what.

* many flights (configs 3, 5): contiguous blocks of flights per rank, no communication;
* one very large grid (config 4): each rank OWNS a band of rows and replays, in log order, every frame
  whose rays can reach it -- cells have exactly one owner, so the bands are exact and are only gathered
  (one all-gather over NCCL), never summed.
"""
from __future__ import annotations

from typing import Tuple

from . import flight_shard as _flight_shard
from . import row_band as _row_band


def flight_shard(n_flights: int, rank: int, world: int) -> Tuple[int, int]:
    """[first, count) of the flights rank ``rank`` replays; blocks differ by at most one flight."""
    return _flight_shard(n_flights, rank, world)


def row_band(H: int, rank: int, world: int, align: int = 4) -> Tuple[int, int]:
    """[row0, rows) of the grid rows rank ``rank`` owns; band edges are multiples of ``align``."""
    return _row_band(H, rank, world, align)
