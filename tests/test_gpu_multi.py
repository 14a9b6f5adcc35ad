"""Multi-GPU entry points of the C library (include/uqs_mapping.h, "Multi-GPU").  The single-rank forms run on
any box; the NCCL paths need >= 2 GPUs (skipped otherwise; the 2-rank CPU version of the partitioning logic is in
test_synth_and_sharding.py): flight shards, owned row bands + ncclAllGather inside the library
(multigpu_worker.py, one process per GPU) and the one-host-thread N-device mode (ncclCommInitAll)."""
import importlib
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from conftest import first_diff

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _c4_slice(synth, n_samples):
    w = synth.scaled(synth.CONFIGS["c4"], n_samples=n_samples)
    d = synth.generate(w)
    x, y = synth.frame_poses(d, d["x_true"], d["y_true"])
    return w, d, x[0], y[0]


def test_banded_replay_without_a_communicator_is_the_whole_grid(gpu, oracle, synth):
    w, d, x, y = _c4_slice(synth, 1500)
    p = w.params()
    assert gpu.comm_nranks() == 1
    grid, st = gpu.replay_banded(p, x, y, d["frame_yaw_deg"][0], d["ranges"][0])
    want, U = oracle.replay(p, x, y, d["frame_yaw_deg"][0], d["ranges"][0])
    assert np.array_equal(grid, want), first_diff(grid, want)
    assert st["ray_cell_updates"] == U and st["frames"] == w.n_frames


def test_one_thread_device_contexts_on_one_gpu(oracle, synth, pkg):
    """uqs_multi_init(1): the context machinery of the one-host-thread mode (select, banded replay, per-context
    pipeline) with a single device; in a subprocess so that the session's own context stays untouched."""
    code = f"""
import importlib, sys
sys.path.insert(0, {ROOT!r})
import numpy as np
m = importlib.import_module("micro-quad-slam_b200"); synth = importlib.import_module("micro-quad-slam_b200.synth")
from oracle import orc
o = orc.Oracle()
w = synth.scaled(synth.CONFIGS["c4"], n_samples=1200); d = synth.generate(w); p = w.params()
x, y = synth.frame_poses(d, d["x_true"], d["y_true"])
m.multi_init(1)
g, st = m.multi_replay_banded(p, x[0], y[0], d["frame_yaw_deg"][0], d["ranges"][0])
want, U = o.replay(p, x[0], y[0], d["frame_yaw_deg"][0], d["ranges"][0])
assert np.array_equal(g, want) and st["ray_cell_updates"] == U
m.multi_select(0)
g2, _ = m.replay(p, x, y, d["frame_yaw_deg"], d["ranges"])          # the single-device calls address the selected context
assert np.array_equal(g2[0], want)
m.multi_shutdown()
m.init(0)                                                            # and the ordinary context still works afterwards
g3, _ = m.replay(p, x, y, d["frame_yaw_deg"], d["ranges"])
assert np.array_equal(g3[0], want)
print("ONE_THREAD_OK")
"""
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ONE_THREAD_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]


def _need_gpus(n):
    import torch
    if torch.cuda.device_count() < n:
        pytest.skip(f"needs >= {n} GPUs")


def test_two_gpu_shards_and_row_bands():
    _need_gpus(2)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "multigpu_worker.py")],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "MULTIGPU_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_one_host_thread_drives_every_gpu(tmp_path):
    """examples/replay_building_multi.c: plain C, ncclCommInitAll, every GPU of the box; multi-GPU grid == 1-GPU grid."""
    _need_gpus(2)
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_next_rows_cpu import write_scanlog
    synth = importlib.import_module("micro-quad-slam_b200.synth")
    n_gpus = min(torch.cuda.device_count(), 8)
    exe = str(tmp_path / "replay_building_multi")
    lib_dir = os.path.join(ROOT, "micro-quad-slam_b200")
    r = subprocess.run(["gcc", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "replay_building_multi.c"),
                        "-L", lib_dir, "-luqs_mapping", "-lm", "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    w = synth.scaled(synth.CONFIGS["c4"], n_samples=3000)
    d = synth.generate(w)
    x, y = synth.frame_poses(d, d["x_true"], d["y_true"])
    recs = []
    for i in range(w.n_frames):
        mm = np.repeat(np.clip(np.nan_to_num(d["ranges"][0, i], nan=65.535) * 1000.0, 0, 65535), 8).reshape(4, 8, 8).transpose(0, 2, 1)
        recs.append({"host_ms": 10 * i, "scan_ms": 10 * i, "x": float(x[0, i]), "y": float(y[0, i]), "yaw": float(d["frame_yaw_deg"][0, i]),
                     "alt": 0.5, "rf": 0.5, "ofx": 0.0, "ofy": 0.0, "q": 200, "kf": 0, "raw": mm.astype("<u2").reshape(-1).view(np.uint8)})
    path = str(tmp_path / "scanlog.bin")
    write_scanlog(path, recs)
    out = subprocess.run([exe, path, "16384", "0.01", str(n_gpus)], capture_output=True, text=True,
                         env=dict(os.environ, LD_LIBRARY_PATH=lib_dir), timeout=600)
    assert out.returncode == 0 and "MATCH" in out.stdout and f"gpus={n_gpus}" in out.stdout, out.stdout + out.stderr
