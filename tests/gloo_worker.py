"""Two-rank gloo worker: exercises the multi-GPU host logic on CPU.

The device replay is replaced by the CPU oracle (test infrastructure) -- what is under test
is the partitioning and the collectives: (a) flights sharded with no data-path collective,
results gathered for comparison; (b) one grid split into owned row bands, each rank replaying
the whole log into ITS rows only, bands all-gathered -- must equal the single-rank grid."""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import orc  # noqa: E402

synth = importlib.import_module("micro-quad-slam_b200.synth")
sharding = importlib.import_module("micro-quad-slam_b200.sharding")


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    o = orc.Oracle()

    # (a) flight shards
    w = synth.scaled(synth.CONFIGS["c3"], n_flights=5, n_samples=120)
    p = w.params()
    first, cnt = sharding.flight_shard(w.n_flights, rank, world)
    d = synth.generate(w, flight_id0=first, n_flights=cnt, n_threads=1)
    mine, _ = o.replay_flights(p, d["x_true"], d["y_true"], d["frame_yaw_deg"], d["ranges"])
    sums = torch.zeros(w.n_flights, dtype=torch.int64)
    sums[first:first + cnt] = torch.from_numpy(mine.reshape(cnt, -1).astype(np.int64).sum(1))
    dist.all_reduce(sums)      # verification only; the data path has no collective
    if rank == 0:
        dall = synth.generate(w, n_threads=1)
        full, _ = o.replay_flights(p, dall["x_true"], dall["y_true"], dall["frame_yaw_deg"], dall["ranges"])
        assert np.array_equal(sums.numpy(), full.reshape(w.n_flights, -1).astype(np.int64).sum(1))

    # (b) owned row bands of one grid + all-gather
    w1 = synth.scaled(synth.CONFIGS["c1"], n_samples=150)
    p1 = w1.params()
    d1 = synth.generate(w1, n_threads=1)
    full, _ = o.replay(p1, d1["x_true"][0], d1["y_true"][0], d1["yaw_deg"][0], d1["ranges"][0])
    r0, rows = sharding.row_band(p1.H, rank, world)
    band = torch.from_numpy(full[r0:r0 + rows].copy())     # owner-computes: rank keeps only its rows
    sizes = [sharding.row_band(p1.H, r, world)[1] for r in range(world)]
    parts = [torch.empty((s, p1.W), dtype=torch.int8) for s in sizes]
    dist.all_gather(parts, band)
    assert np.array_equal(torch.cat(parts).numpy(), full)
    dist.barrier()
    if rank == 0:
        print("GLOO_WORKER_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
