"""SURVEY section 8(f) rows on the CPU: the oracle's restatements of N1-N3 pinned against the reference's
own functions (oracle/_ref), and the N4 scan-log reader (host C, no GPU needed)."""
import os
import struct

import numpy as np
import pytest

from conftest import first_diff, have_ref


def random_scans(rng, n):
    mm = rng.integers(0, 4500, (n, 256)).astype(np.uint16)
    mm[rng.random((n, 256)) < 0.15] = 0xFFFF
    mm[rng.random((n, 256)) < 0.10] = 0
    mm[rng.random((n, 256)) < 0.05] = rng.integers(1, 25)          # <= 0.02 m region
    mm[0] = 0xFFFF                                                 # a frame with no return at all
    mm[1, :64] = 1234                                              # all rows equal
    return mm.astype("<u2").view(np.uint8).reshape(n, 512)


def test_n1_beams_oracle_equals_reference(orc_mod, oracle):
    if not have_ref(orc_mod, 400, 400, "0.05"):
        pytest.skip("oracle/_ref not built")
    ref = orc_mod.Reference(400, 400, "0.05")
    raw = random_scans(np.random.default_rng(2), 400)
    beams, _ = oracle.beams_from_scans(raw)
    for i in range(raw.shape[0]):
        want = ref.beams_from_frame(raw[i])
        assert np.array_equal(beams[i].view(np.uint32), want.view(np.uint32)), i
    assert np.isnan(beams[0]).all() and (beams[1, :8] == np.float32(1234) * np.float32(0.001)).all()


def wandering_log(synth, n=2500, res="0.05", W=400, size=20.0, speed=4.0):
    """a log that leaves the 60 % box several times (forces recentering)"""
    w = synth.Workload("wander", 99, 1, n, W, res, size, 50.0)
    d = synth.generate(w)
    t = np.arange(n, dtype=np.float32) / np.float32(50.0)
    x = (d["x_true"][0] + np.float32(speed * 0.35) * t).astype(np.float32)      # drifts out of the map
    y = (d["y_true"][0] - np.float32(speed * 0.22) * t).astype(np.float32)
    return w, d, x, y


def test_n2_recentering_oracle_equals_reference(orc_mod, oracle, synth):
    if not have_ref(orc_mod, 400, 400, "0.05"):
        pytest.skip("oracle/_ref not built")
    ref = orc_mod.Reference(400, 400, "0.05")
    w, d, x, y = wandering_log(synth)
    p = w.params()
    want = ref.replay(x, y, d["yaw_deg"][0], d["ranges"][0], allow_recenter=True)
    got, origin, n_ev, U = oracle.replay_recentering(p, x, y, d["yaw_deg"][0], d["ranges"][0])
    assert ref.recentered() and n_ev >= 3
    assert np.array_equal(got, want), first_diff(got, want)
    assert np.float32(origin[0]) == np.float32(ref.origin()[0]) and np.float32(origin[1]) == np.float32(ref.origin()[1])


def test_n3_frontier_oracle_equals_reference(orc_mod, oracle, synth):
    if not have_ref(orc_mod, 400, 400, "0.05"):
        pytest.skip("oracle/_ref not built")
    ref = orc_mod.Reference(400, 400, "0.05")
    w = synth.scaled(synth.CONFIGS["c1"], n_samples=1200)
    d = synth.generate(w)
    p = w.params()
    grid = ref.replay(d["x_true"][0], d["y_true"][0], d["yaw_deg"][0], d["ranges"][0])
    rng = np.random.default_rng(4)
    for _ in range(400):
        x, y = np.float32(rng.uniform(-9.9, 9.9)), np.float32(rng.uniform(-9.9, 9.9))
        yaw, off = np.float32(rng.uniform(-180, 180)), np.float32(rng.choice([0.0, 90.0, -90.0, 180.0, 37.5]))
        assert oracle.frontier_score(p, grid, x, y, yaw, off) == ref.frontier_score(x, y, yaw, off)


def write_scanlog(path, recs, extra_header_at=None):
    """scanrec_t exactly as uav_local_nav.c:1522-1547 packs it (569 bytes, little endian)."""
    with open(path, "wb") as f:
        f.write(b"SCLOG2\n")
        for i, r in enumerate(recs):
            if extra_header_at is not None and i == extra_header_at:
                f.write(b"SCLOG2\n")
            f.write(struct.pack("<III4f2f3fBBBHI", 0x324E4353, r["host_ms"], r["scan_ms"], r["x"], r["y"], r["yaw"], r["alt"],
                                0.01, -0.02, r["rf"], r["ofx"], r["ofy"], r["q"], 3, r["kf"], 0, 0x1234))
            f.write(bytes(r["raw"]))


def test_n4_scanlog_reader(pkg, tmp_path):
    rng = np.random.default_rng(8)
    recs = []
    for i in range(57):
        recs.append({"host_ms": 1000 + 100 * i, "scan_ms": 7 + 100 * i, "x": float(np.float32(0.1 * i)), "y": float(np.float32(-0.05 * i)),
                     "yaw": float(np.float32(3.0 * i - 90)), "alt": 0.5, "rf": 0.48, "ofx": float(np.float32(0.01 * i)),
                     "ofy": -0.25, "q": 200 - i, "kf": (1 << 5) if i == 20 else 0, "raw": rng.integers(0, 256, 512, dtype=np.uint8)})
    recs[3]["x"] = float("nan")           # no position yet (:1559): skipped
    recs[9]["yaw"] = float("nan")         # no attitude yet (:1561): skipped
    path = str(tmp_path / "scanlog.bin")
    write_scanlog(path, recs, extra_header_at=30)
    assert os.path.getsize(path) == 7 * 2 + 57 * 569
    d = pkg.scanlog_read(path)
    keep = [r for i, r in enumerate(recs) if i not in (3, 9)]
    assert d["x_m"].size == 55
    assert np.array_equal(d["host_ms"], np.array([r["host_ms"] for r in keep], np.uint32))
    assert np.array_equal(d["x_m"], np.array([r["x"] for r in keep], np.float32))
    assert np.array_equal(d["yaw_deg"], np.array([r["yaw"] for r in keep], np.float32))
    assert np.array_equal(d["of_rate_x"], np.array([r["ofx"] for r in keep], np.float32))
    assert np.array_equal(d["of_q"], np.array([r["q"] for r in keep], np.uint8))
    assert d["kf_flags"][18] == (1 << 5)
    assert np.array_equal(d["grid_raw"], np.stack([r["raw"] for r in keep]))
    assert pkg.scanlog_read(path, keep_nan_pose=True)["x_m"].size == 57
    # truncated tail (power cut mid-record) and bad files
    with open(path, "ab") as f:
        f.write(b"\x53\x43\x4e\x32" + b"\x00" * 100)
    assert pkg.scanlog_read(path)["x_m"].size == 55
    bad = str(tmp_path / "bad.bin")
    open(bad, "wb").write(b"NOTALOG" + b"\x00" * 600)
    with pytest.raises(pkg.UqsError):
        pkg.scanlog_read(bad)
    with pytest.raises(pkg.UqsError):
        pkg.scanlog_read(str(tmp_path / "missing.bin"))


NAVLOG_HEADER = ("t_ms,state,want_arm,armed,mode,yaw_deg,alt_m,alt_src,x_m,y_m,vx_mps,vy_mps,"
                 "rf_m,of_q,of_rate_x,of_rate_y,tof_f,tof_r,tof_b,tof_l,batt_v,batt_cells\n")


def navlog_row(r):
    """One row exactly as log_tick formats it (uav_local_nav.c:1586-1625): %.3f / %.4f fields, "nan" when missing."""
    f3 = lambda v: "nan" if v is None else "%.3f" % v
    s = "%d,%s,%d,%d,%d," % (r["t"], r["state"], 1, 1, 4)
    s += f3(r["yaw"]) + "," + f3(r["alt"]) + ",RF,"
    s += ("%.3f,%.3f,%.3f,%.3f," % (r["x"], r["y"], r["vx"], r["vy"])) if r["x"] is not None else "nan,nan,nan,nan,"
    s += f3(r["rf"]) + "," + "%d," % r["q"]
    s += ("%.4f,%.4f," % (r["ofx"], r["ofy"])) if r["ofx"] is not None else "nan,nan,"
    s += "%.3f,%.3f,%.3f,%.3f," % tuple(r["tof"])
    s += "%.3f,%d\n" % (11.1, 3)
    return s


def test_n4_navlog_reader(pkg, tmp_path):
    """navlog.csv -> the SoA columns P0 consumes; 'nan' fields, appended flights, a second header, a truncated tail."""
    rng = np.random.default_rng(12)
    rows = []
    for i in range(73):
        rows.append({"t": 5_000_000_000 + 100 * i, "state": "HOVER" if i % 3 else "TURN", "yaw": float(rng.uniform(-180, 180)),
                     "alt": 0.5 + 0.001 * i, "x": float(rng.uniform(-5, 5)), "y": float(rng.uniform(-5, 5)),
                     "vx": float(rng.uniform(-1, 1)), "vy": float(rng.uniform(-1, 1)), "rf": 0.48, "q": int(rng.integers(0, 256)),
                     "ofx": float(rng.uniform(-2, 2)), "ofy": float(rng.uniform(-2, 2)), "tof": rng.uniform(0.1, 4.0, 4).tolist()})
    rows[4]["yaw"] = None                       # no attitude yet (:1596-1597)
    rows[5]["x"] = None                         # no local position (:1603-1605)
    rows[6]["ofx"] = None; rows[6]["q"] = 0     # stale optical flow (:1611-1616)
    rows[7]["rf"] = None                        # stale rangefinder (:1607-1609)
    path = str(tmp_path / "navlog.csv")
    with open(path, "w") as f:
        f.write(NAVLOG_HEADER)
        for i, r in enumerate(rows):
            if i == 40:
                f.write(NAVLOG_HEADER)          # a log started in a fresh file and concatenated by hand
            f.write(navlog_row(r))
        f.write(navlog_row(rows[0])[:37])       # power cut mid-row: no newline, too few columns
    d = pkg.navlog_read(path)
    assert d["t_ms"].size == 73
    f32 = lambda key: np.array([np.float32("nan") if r[key] is None else np.float32(float("%.3f" % r[key])) for r in rows], np.float32)
    assert np.array_equal(d["t_ms"], np.array([r["t"] & 0xFFFFFFFF for r in rows], np.uint32))
    for key, col in (("yaw", "yaw_deg"), ("alt", "alt_m"), ("x", "x_m"), ("rf", "rf_m")):
        assert np.array_equal(d[col], f32(key), equal_nan=True), col
    assert np.isnan(d["y_m"][5]) and np.isnan(d["vx_mps"][5]) and np.isnan(d["of_rate_y"][6])
    want_ofx = np.array([np.float32("nan") if r["ofx"] is None else np.float32(float("%.4f" % r["ofx"])) for r in rows], np.float32)
    assert np.array_equal(d["of_rate_x"], want_ofx, equal_nan=True)
    assert np.array_equal(d["of_q"], np.array([r["q"] for r in rows], np.uint8))
    assert np.array_equal(d["tof4"], np.array([[np.float32(float("%.3f" % v)) for v in r["tof"]] for r in rows], np.float32))
    # the columns feed P0 as they are: t differences survive the 32-bit truncation
    assert np.all(np.diff(d["t_ms"].astype(np.int64)) % (1 << 32) == 100)
    with pytest.raises(pkg.UqsError):
        pkg.navlog_read(str(tmp_path / "missing.csv"))
