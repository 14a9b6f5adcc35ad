"""Generator determinism/bounds and the host-side multi-GPU partitioning (incl. a 2-rank gloo run)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_generator_is_deterministic_and_flight_addressable(synth):
    w = synth.scaled(synth.CONFIGS["c3"], n_flights=6, n_samples=200)
    a = synth.generate(w, n_threads=1)
    b = synth.generate(w, n_threads=4)
    for k in a:
        assert np.array_equal(a[k], b[k], equal_nan=True), k
    tail = synth.generate(w, flight_id0=4, n_flights=2)
    assert np.array_equal(tail["ranges"], a["ranges"][4:], equal_nan=True)
    assert np.array_equal(tail["of_rate_x"], a["of_rate_x"][4:])
    assert not np.array_equal(a["ranges"][0], a["ranges"][1], equal_nan=True)


@pytest.mark.parametrize("name", ["c1", "c2", "c3", "c4"])
def test_trajectories_stay_inside_60_percent_of_half_extent(synth, name):
    """so that map_recentre_if_needed (uav_local_nav.c:328-332) could never fire during a replay"""
    w = synth.CONFIGS[name]
    ws = synth.scaled(w, n_flights=min(w.n_flights, 2), n_samples=min(w.n_samples, 30000))
    d = synth.generate(ws)
    lim = 0.6 * 0.5 * w.size_m
    assert np.abs(d["x_true"]).max() < lim and np.abs(d["y_true"]).max() < lim
    r = d["ranges"]
    assert np.nanmin(r) >= 0.02 and np.nanmax(r) <= 4.0
    assert 0.005 < np.isnan(r).mean() < 0.04
    assert d["ranges"].shape == (ws.n_flights, ws.n_frames, 32)
    assert (np.diff(d["t_ms"][0].astype(np.int64)) > 0).all()


def test_c4_has_two_reference_frames_per_sample(synth):
    w = synth.scaled(synth.CONFIGS["c4"], n_samples=100)
    d = synth.generate(w)
    assert d["frame_yaw_deg"].shape == (1, 200)
    assert np.array_equal(d["frame_yaw_deg"][0, 1::2], d["yaw_deg"][0] + np.float32(45.0))
    assert np.array_equal(d["frame_sample"], np.repeat(np.arange(100), 2))


def test_c5_geometries_match_the_oracle_build_list(synth):
    lines = [l.split() for l in open(os.path.join(ROOT, "oracle", "c5_geometries.txt")) if l.strip()]
    assert [(int(a), b) for a, b in lines] == [(synth.c5_width(r), r) for r in synth.C5_RES]
    assert len(synth.C5_SIGMA_R) == 16 and synth.C5_SIGMA_R[0] == 0.0 and abs(synth.C5_SIGMA_R[-1] - 0.10) < 1e-9


def test_flight_shards_partition_exactly(pkg):
    sh = __import__("importlib").import_module("micro-quad-slam_b200.sharding")
    for n in [1, 7, 64, 4096, 16384]:
        for world in [1, 2, 3, 4, 8]:
            spans = [sh.flight_shard(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == n
            for (a, c), (b, _) in zip(spans, spans[1:]):
                assert a + c == b
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


def test_row_bands_partition_exactly(pkg):
    sh = __import__("importlib").import_module("micro-quad-slam_b200.sharding")
    for H in [400, 2000, 16384, 334, 7]:
        for world in [1, 2, 4, 8]:
            bands = [sh.row_band(H, r, world) for r in range(world)]
            assert bands[0][0] == 0 and sum(c for _, c in bands) == H
            for (a, c), (b, _) in zip(bands, bands[1:]):
                assert a + c == b


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_rank_gloo_sharded_replay_equals_single_rank():
    """world_size-2 gloo run of the N>1 host logic (flight shards gathered; row bands all-gathered) with
    the CPU oracle standing in for the device replay -- see tests/gloo_worker.py."""
    port = _free_port()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "gloo_worker.py")], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=300)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
    assert "GLOO_WORKER_OK" in outs[0]
