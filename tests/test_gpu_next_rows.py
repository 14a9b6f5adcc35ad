"""SURVEY section 8(f) rows on the device, through the C ABI, against the oracle (pinned to the reference in
tests/test_next_rows_cpu.py) and the reference's own code where oracle/_ref is present."""
import numpy as np
import pytest

from conftest import first_diff, have_ref
from test_next_rows_cpu import random_scans, wandering_log, write_scanlog

pytestmark = pytest.mark.gpu


def test_n1_scans_to_beams(gpu, oracle):
    raw = random_scans(np.random.default_rng(12), 3000)
    beams, dmin = gpu.beams_from_scans(raw)
    wb, wd = oracle.beams_from_scans(raw)
    assert np.array_equal(beams.view(np.uint32), wb.view(np.uint32))
    assert np.array_equal(dmin.view(np.uint32), wd.view(np.uint32))


def test_n2_replay_with_recentering(gpu, oracle, orc_mod, synth):
    w, d, x, y = wandering_log(synth)
    p = w.params()
    grid, origin, events, st = gpu.replay_recentering(p, x, y, d["yaw_deg"][0], d["ranges"][0])
    want, worigin, n_ev, U = oracle.replay_recentering(p, x, y, d["yaw_deg"][0], d["ranges"][0])
    assert len(events) == n_ev >= 3
    assert np.array_equal(grid, want), first_diff(grid, want)
    assert np.float32(origin[0]) == np.float32(worigin[0]) and np.float32(origin[1]) == np.float32(worigin[1])
    assert st["ray_cell_updates"] == U
    if have_ref(orc_mod, 400, 400, "0.05"):
        ref = orc_mod.Reference(400, 400, "0.05")
        rg = ref.replay(x, y, d["yaw_deg"][0], d["ranges"][0], allow_recenter=True)
        assert np.array_equal(grid, rg)


def test_n2_n3_dropin_symbols(gpu, oracle, orc_mod, synth):
    """map_recentre_if_needed + map_update_from_beams + frontier_score_dir called like log_tick / control_tick do."""
    w, d, x, y = wandering_log(synth, n=900, speed=9.0)
    p = w.params()
    di = gpu.DropIn(p)
    di.hover_init(float(x[0]), float(y[0]))
    q = p.copy()
    q.origin_x, q.origin_y = x[0], y[0]
    want, worigin, n_ev, _ = oracle.replay_recentering(q, x, y, d["yaw_deg"][0], d["ranges"][0])
    scores = []
    for i in range(900):
        di.set_beams(d["ranges"][0, i])
        di.map_recentre_if_needed(x[i], y[i])
        di.map_update_from_beams(x[i], y[i], d["yaw_deg"][0, i])
        if i % 150 == 149:
            scores.append((i, di.frontier_score_dir(x[i], y[i], d["yaw_deg"][0, i], 90.0)))
    got = di.grid()
    assert n_ev >= 2 and np.array_equal(got, want), first_diff(got, want)
    assert np.float32(di.origin()[0]) == np.float32(worigin[0]) and di.kf_flags() & (1 << 5)
    # frontier scores at those instants against the oracle replayed up to the same frame
    for i, s in scores:
        g, o, _, _ = oracle.replay_recentering(q, x[:i + 1], y[:i + 1], d["yaw_deg"][0, :i + 1], d["ranges"][0, :i + 1])
        r = q.copy()
        r.origin_x, r.origin_y = np.float32(o[0]), np.float32(o[1])
        assert s == oracle.frontier_score(r, g, x[i], y[i], d["yaw_deg"][0, i], np.float32(90.0)), i


def test_n3_frontier_scores_batch(gpu, oracle, synth):
    w = synth.scaled(synth.CONFIGS["c1"], n_samples=1500)
    d = synth.generate(w)
    p = w.params()
    grid, _ = gpu.replay(p, d["x_true"], d["y_true"], d["frame_yaw_deg"], d["ranges"])
    rng = np.random.default_rng(6)
    n = 3000
    x, y = rng.uniform(-10.2, 10.2, n).astype(np.float32), rng.uniform(-10.2, 10.2, n).astype(np.float32)
    yaw, off = rng.uniform(-180, 180, n).astype(np.float32), rng.choice([0.0, 90.0, -90.0, 180.0], n).astype(np.float32)
    got = gpu.frontier_scores(p, grid[0], x, y, yaw, off)
    want = np.array([oracle.frontier_score(p, grid[0], x[i], y[i], yaw[i], off[i]) for i in range(n)], np.int32)
    assert np.array_equal(got, want), int((got != want).sum())
    assert len(set(got.tolist())) > 20


def test_scanlog_to_grid_end_to_end(gpu, oracle, synth, tmp_path):
    """N4 -> N1 -> hot path: a scanlog.bin written in the reference's record format is read back, its raw 8x8
    scans reduced to beams on the device and replayed; equals the oracle fed the same records."""
    rng = np.random.default_rng(21)
    w = synth.scaled(synth.CONFIGS["c1"], n_samples=400)
    d = synth.generate(w)
    p = w.params()
    recs = []
    for i in range(400):
        mm = np.clip(np.repeat(np.nan_to_num(d["ranges"][0, i], nan=65.535) * 1000.0, 8).reshape(4, 8, 8).transpose(0, 2, 1), 0, 65535)
        mm = (mm + rng.integers(-15, 15, mm.shape)).clip(0, 65535).astype("<u2")       # 8 rows per column, jittered
        recs.append({"host_ms": 20 * i, "scan_ms": 20 * i, "x": float(d["x_true"][0, i]), "y": float(d["y_true"][0, i]),
                     "yaw": float(d["yaw_deg"][0, i]), "alt": 0.5, "rf": 0.5, "ofx": 0.0, "ofy": 0.0, "q": 200, "kf": 0,
                     "raw": mm.reshape(-1).view(np.uint8)})
    path = str(tmp_path / "scanlog.bin")
    write_scanlog(path, recs)
    log = gpu.scanlog_read(path)
    beams, _ = gpu.beams_from_scans(log["grid_raw"])
    grid, st = gpu.replay(p, log["x_m"], log["y_m"], log["yaw_deg"], beams)
    ob, _ = oracle.beams_from_scans(log["grid_raw"])
    want, U = oracle.replay(p, log["x_m"], log["y_m"], log["yaw_deg"], ob)
    assert np.array_equal(grid[0], want), first_diff(grid[0], want)
    assert st["ray_cell_updates"] == U > 100000


def test_plain_c_harness_links_and_matches(gpu, oracle, synth, tmp_path):
    """examples/replay_scanlog.c: host code in plain C against include/uqs_mapping.h; batch route and drop-in
    route must agree with each other and with the oracle."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "replay_scanlog")
    lib_dir = os.path.join(root, "micro-quad-slam_b200")
    r = subprocess.run(["gcc", "-O2", "-Wall", "-I", os.path.join(root, "include"), os.path.join(root, "examples", "replay_scanlog.c"),
                        "-L", lib_dir, "-luqs_mapping", "-lm", "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    w = synth.scaled(synth.CONFIGS["c1"], n_samples=300)
    d = synth.generate(w)
    recs = []
    for i in range(300):
        mm = np.repeat(np.clip(np.nan_to_num(d["ranges"][0, i], nan=65.535) * 1000.0, 0, 65535), 8).reshape(4, 8, 8).transpose(0, 2, 1)
        recs.append({"host_ms": 20 * i, "scan_ms": 20 * i, "x": float(d["x_true"][0, i]), "y": float(d["y_true"][0, i]),
                     "yaw": float(d["yaw_deg"][0, i]), "alt": 0.5, "rf": 0.5, "ofx": 0.0, "ofy": 0.0, "q": 200, "kf": 0,
                     "raw": mm.astype("<u2").reshape(-1).view(np.uint8)})
    path = str(tmp_path / "scanlog.bin")
    write_scanlog(path, recs)
    out = subprocess.run([exe, path, "400", "0.05"], capture_output=True, text=True, env=dict(os.environ, LD_LIBRARY_PATH=lib_dir))
    assert out.returncode == 0 and "MATCH" in out.stdout, out.stdout + out.stderr
    log = gpu.scanlog_read(path)
    p = w.params()
    p.origin_x, p.origin_y = log["x_m"][0], log["y_m"][0]
    ob, _ = oracle.beams_from_scans(log["grid_raw"])
    want, U = oracle.replay(p, log["x_m"], log["y_m"], log["yaw_deg"], ob)
    assert f"fnv_batch={oracle.fnv1a32(want):08x}" in out.stdout and f"updates={U}" in out.stdout, out.stdout
