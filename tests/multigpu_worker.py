"""torchrun worker (one process per GPU, NCCL): the two multi-GPU partitionings of DESIGN.md section 8.

  (a) config 3/5: flights sharded over ranks, no data-path collective; per-flight checksums are all-reduced
      only to be compared with the single-GPU result on rank 0;
  (b) config 4: one large grid split into row bands OWNED by ranks (uqs_replay_dev row0/rows); every rank
      replays the whole log into its band; one NCCL all-gather assembles the grid, which must equal the
      oracle's byte for byte.
Run:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tests/multigpu_worker.py
"""
import importlib
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
m = importlib.import_module("micro-quad-slam_b200")
synth = importlib.import_module("micro-quad-slam_b200.synth")
sharding = importlib.import_module("micro-quad-slam_b200.sharding")
from oracle import orc  # noqa: E402  (checker only)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    m.init(local)
    m.set_stream(torch.cuda.current_stream().cuda_stream)
    o = orc.Oracle()

    # (a) flight shards -------------------------------------------------------------------------
    w = synth.scaled(synth.CONFIGS["c3"], n_flights=13, n_samples=400)
    p = w.params()
    first, cnt = sharding.flight_shard(w.n_flights, rank, world)
    d = synth.generate(w, flight_id0=first, n_flights=cnt)
    grids, px, py, st = m.replay_flow(p, d["t_ms"], d["of_rate_x"], d["of_rate_y"], d["h_m"], d["yaw_deg"], d["of_q"], d["ranges"])
    sums = torch.zeros(w.n_flights, dtype=torch.int64, device=dev)
    sums[first:first + cnt] = torch.from_numpy(grids.reshape(cnt, -1).astype(np.int64).sum(1)).to(dev)
    dist.all_reduce(sums)
    if rank == 0:
        dall = synth.generate(w)
        ox, oy = o.pose_integrate(dall["t_ms"], dall["of_rate_x"], dall["of_rate_y"], dall["h_m"], dall["yaw_deg"], dall["of_q"])
        want, _ = o.replay_flights(p, ox, oy, dall["frame_yaw_deg"], dall["ranges"])
        assert np.array_equal(sums.cpu().numpy(), want.reshape(w.n_flights, -1).astype(np.int64).sum(1)), "flight shards differ"
        print(f"[multigpu] flight shards over {world} GPUs: checksums equal the oracle's", flush=True)

    # (b) owned row bands of one 16384^2 grid + one NCCL all-gather, all inside the C library -------------------
    uid = torch.zeros(m.COMM_ID_BYTES, dtype=torch.uint8, device=dev)
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(m.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, 0)                       # the launcher's own plumbing carries the 128 bytes
    m.comm_init_rank(bytes(uid.cpu().numpy().tobytes()), world, rank)
    assert m.comm_nranks() == world
    n_samples = int(os.environ.get("UQS_C4_SAMPLES", "4000"))
    w4 = synth.scaled(synth.CONFIGS["c4"], n_samples=n_samples)
    p4 = w4.params()
    d4 = synth.generate(w4)                       # every rank sees the whole log
    x, y = synth.frame_poses(d4, d4["x_true"], d4["y_true"])
    tx, ty, tyaw, tr = (torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (x, y, d4["frame_yaw_deg"], d4["ranges"]))
    full = torch.full((p4.H, p4.W), 5, dtype=torch.int8, device=dev)      # stale contents must all be overwritten
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    m.replay_banded_dev(p4, w4.n_frames, tx.data_ptr(), ty.data_ptr(), tyaw.data_ptr(), tr.data_ptr(), full.data_ptr(), gather=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    # the host-buffer form: each rank uploads 1/world of the log, NCCL all-gathers it
    hgrid, st = m.replay_banded(p4, x[0], y[0], d4["frame_yaw_deg"][0], d4["ranges"][0], want_grid=(rank == 0))
    want, U = o.replay(p4, x[0], y[0], d4["frame_yaw_deg"][0], d4["ranges"][0])
    got = full.cpu().numpy()
    assert np.array_equal(got, want), f"rank {rank}: {int((got != want).sum())} cells differ after the all-gather"
    assert st["ray_cell_updates"] == U
    if rank == 0:
        assert np.array_equal(hgrid, want), "host-buffer banded replay differs"
        print(f"[multigpu] 16384^2 grid in {world} owned row bands + ncclAllGather inside libuqs_mapping (NCCL {m.nccl_version()}): "
              f"byte-identical to the oracle on every rank ({U} updates, {dt*1e3:.1f} ms incl. gather)", flush=True)
        print("MULTIGPU_OK", flush=True)
    m.comm_destroy()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
