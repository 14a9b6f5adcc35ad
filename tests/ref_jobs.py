"""Expected results of whole BASELINE configurations from the REFERENCE'S OWN CODE (oracle/_ref, built from
/root/reference by oracle/build_ref.sh), as 64-bit grid digests -- test infrastructure for
test_gpu_zz_complete.py.  P0 has no reference counterpart: poses come from the CPU statement of the P0 spec
(oracle/uqs_oracle.c), the mapping from the reference library.

  python tests/ref_jobs.py c2|c4 out.json      one long log on one core (run in the background by conftest.py)
  flights_digests(...)                          many flights, one forked worker per host core
"""
import importlib
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _mods():
    from oracle import orc
    synth = importlib.import_module("micro-quad-slam_b200.synth")
    return orc, synth


def _ref_grid_digest(o, ref):
    cells = ref.W * ref.H
    return int(o.L.orc_grid_hash64(ref.L.ref_grid(), cells))


def single_log(name: str) -> dict:
    """Digest of config 2 (P0 + mapping) or config 4 (true poses, two frames per sample) from the reference's code."""
    orc, synth = _mods()
    o = orc.Oracle()
    w = synth.CONFIGS[name]
    d = synth.generate(w)
    t0 = time.perf_counter()
    if name == "c2":
        x, y = o.pose_integrate(d["t_ms"][0], d["of_rate_x"][0], d["of_rate_y"][0], d["h_m"][0], d["yaw_deg"][0], d["of_q"][0])
    else:
        xs, ys = synth.frame_poses(d, d["x_true"], d["y_true"])
        x, y = xs[0], ys[0]
    ref = orc.Reference(w.W, w.W, w.res)
    ref.reset(0.0, 0.0)
    x, y = np.ascontiguousarray(x, np.float32), np.ascontiguousarray(y, np.float32)
    yaw = np.ascontiguousarray(d["frame_yaw_deg"][0], np.float32)
    rng = np.ascontiguousarray(d["ranges"][0], np.float32)
    ref.L.ref_replay(x.size, x.ctypes.data, y.ctypes.data, yaw.ctypes.data, rng.ctypes.data, 0)
    return {"config": name, "digest": _ref_grid_digest(o, ref), "frames": int(x.size), "seconds": time.perf_counter() - t0}


def _flight_worker(c, cores, job, out):
    orc, synth = _mods()
    o = orc.Oracle()
    ref = orc.Reference(job["W"], job["W"], job["res"])
    for i in range(c, len(job["flights"]), cores):
        w, fid = job["flights"][i]
        d = synth.generate(w, flight_id0=fid, n_flights=1, n_threads=1)
        if job["flow"]:
            x, y = o.pose_integrate(d["t_ms"][0], d["of_rate_x"][0], d["of_rate_y"][0], d["h_m"][0], d["yaw_deg"][0], d["of_q"][0])
        else:
            x, y = d["x_true"][0], d["y_true"][0]
        x, y = np.ascontiguousarray(x, np.float32), np.ascontiguousarray(y, np.float32)
        ref.reset(0.0, 0.0)
        ref.L.ref_replay(x.size, x.ctypes.data, y.ctypes.data, d["frame_yaw_deg"][0].ctypes.data, d["ranges"][0].ctypes.data, 0)
        out[i] = _ref_grid_digest(o, ref)


def flights_digests(flights, W, res, flow, cores=None):
    """flights = [(workload, flight_id), ...] sharing one geometry.  One forked process per core (the reference's
    static grid forbids threads); returns the digests as np.uint64 [len(flights)]."""
    cores = min(cores or os.cpu_count() or 1, len(flights))
    ctx = mp.get_context("fork")
    out = ctx.Array("Q", len(flights), lock=False)
    job = {"flights": flights, "W": W, "res": res, "flow": flow}
    procs = [ctx.Process(target=_flight_worker, args=(c, cores, job, out)) for c in range(cores)]
    for pr in procs:
        pr.start()
    for pr in procs:
        pr.join()
    if any(pr.exitcode != 0 for pr in procs):
        raise RuntimeError("reference worker failed")
    return np.frombuffer(out, dtype=np.uint64).copy()


if __name__ == "__main__":
    res = single_log(sys.argv[1])
    with open(sys.argv[2] + ".tmp", "w") as f:
        json.dump(res, f)
    os.replace(sys.argv[2] + ".tmp", sys.argv[2])
