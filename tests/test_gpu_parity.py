"""Device parity, all through the C ABI (libuqs_mapping.so): bit-exact cell indices and int8 log-odds
against the CPU oracle (and the reference's own code where oracle/_ref is present), on seeded inputs."""
import ctypes
import os

import numpy as np
import pytest

from conftest import first_diff, have_ref

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["tiles", "resident4", "resident8", "resident16", "resident32", "resident4fan", "resident8fan", "resident16fan",
                        "resident32fan", "resident4dec", "resident8dec", "resident16dec"])
def engine(request, gpu):
    """Both replay engines (the resident engine's warp counts, both of its lane layouts, shared or dedicated decode
    warp) must give identical bytes."""
    name = request.param
    fan, dec = name.endswith("fan"), name.endswith("dec")
    e, nw = {"tiles": (1, 0), "resident4": (2, 4), "resident8": (2, 8), "resident16": (2, 16), "resident32": (2, 32)}[name[:-3] if fan or dec else name]
    gpu.set_engine(e, nw)
    gpu.set_fan_layout(1 if fan else 0)
    gpu.set_decode_warp(1 if dec else 0)
    yield name
    gpu.set_engine(0, 0)
    gpu.set_fan_layout(-1)
    gpu.set_decode_warp(-1)


def oracle_grids(oracle, p, d, x=None, y=None):
    x = d["x_true"] if x is None else x
    y = d["y_true"] if y is None else y
    return oracle.replay_flights(p, x, y, d["frame_yaw_deg"], d["ranges"])


# ----------------------------------------------------------------------------------------------
# A6 third-party arithmetic: glibc sincosf restated on the device
# ----------------------------------------------------------------------------------------------
def test_device_sincosf_equals_host_libm(gpu, oracle):
    """strided sweep over every binade of |y| < 120 plus dense neighbourhoods of the branch points;
    the exhaustive sweep is tests/test_gpu_sweeps.py::test_device_sincosf_exhaustive."""
    bits = np.arange(0, 0x42F00000, 193, dtype=np.uint32)
    edges = []
    for b in (0x39800000, 0x3F490FDB, 0x3FC90FDB, 0x40490FDB, 0x40C90FDB, 0x42F00000 - 70000):
        edges.append(np.arange(b - 65536, b + 65536, dtype=np.uint32))
    bits = np.concatenate([bits] + edges)
    bits = np.concatenate([bits, bits | np.uint32(0x80000000)])
    a = bits.view(np.float32)
    ds, dc = gpu.sincosf_batch(a)
    hs, hc = oracle.libm_sincosf(a)
    bad = np.flatnonzero((ds.view(np.uint32) != hs.view(np.uint32)) | (dc.view(np.uint32) != hc.view(np.uint32)))
    assert bad.size == 0, f"{bad.size} mismatches, first y={a[bad[0]]!r}"


# ----------------------------------------------------------------------------------------------
# A3 + A6: cell indices, one by one
# ----------------------------------------------------------------------------------------------
def test_beam_end_cells_equal_oracle(gpu, oracle, synth):
    rng = np.random.default_rng(7)
    n = 6000
    p = gpu.make_params(400, 400, 0.05, 20.0)
    p.origin_x, p.origin_y = np.float32(0.31), np.float32(-1.07)
    x = rng.uniform(-11, 11, n).astype(np.float32)            # some poses off the 20 m grid
    y = rng.uniform(-11, 11, n).astype(np.float32)
    yaw = rng.uniform(-400, 400, n).astype(np.float32)
    r = rng.uniform(0, 4.6, (n, 32)).astype(np.float32)
    r[rng.random((n, 32)) < 0.05] = np.nan
    r[rng.random((n, 32)) < 0.05] = np.float32(0.05)
    r[rng.random((n, 32)) < 0.03] = np.float32(3.95)
    # poses exactly on half-cell ties
    x[:200] = np.float32(0.025) + np.float32(0.05) * np.arange(200, dtype=np.float32) - np.float32(5.0) + p.origin_x
    cells, origin = gpu.beam_cells(p, x, y, yaw, r)
    wc, wo = oracle.beam_cells(p, x, y, yaw, r)
    assert np.array_equal(origin, wo)
    assert np.array_equal(cells, wc), first_diff(cells, wc)
    assert (cells[..., 0] >= 0).mean() > 0.3


def test_nonfinite_and_huge_poses_follow_the_reference(gpu, oracle):
    p = gpu.make_params(400, 400, 0.05, 20.0)
    x = np.array([np.nan, np.inf, -np.inf, 1e30, 3e9, 2.0 ** 32 * 0.05, 0.0, -0.0], np.float32)
    y = np.zeros_like(x)
    yaw = np.zeros_like(x)
    r = np.full((x.size, 32), 1.0, np.float32)
    cells, origin = gpu.beam_cells(p, x, y, yaw, r)
    wc, wo = oracle.beam_cells(p, x, y, yaw, r)
    assert np.array_equal(origin, wo) and np.array_equal(cells, wc)


# ----------------------------------------------------------------------------------------------
# A1-A6: whole replays
# ----------------------------------------------------------------------------------------------
def test_c1_single_flight_bit_exact(gpu, oracle, orc_mod, synth, engine):
    w = synth.CONFIGS["c1"]
    d = synth.generate(w)
    p = w.params()
    got, st = gpu.replay(p, d["x_true"], d["y_true"], d["frame_yaw_deg"], d["ranges"])
    want, U = oracle_grids(oracle, p, d)
    assert np.array_equal(got, want), first_diff(got, want)
    assert st["ray_cell_updates"] == U and st["frames"] == 3000 and st["domain_errors"] == 0
    assert st["rays_accepted"] + st["rays_skipped"] == 3000 * 32
    if have_ref(orc_mod, 400, 400, "0.05"):
        ref = orc_mod.Reference(400, 400, "0.05")
        rg = ref.replay(d["x_true"][0], d["y_true"][0], d["frame_yaw_deg"][0], d["ranges"][0])
        assert np.array_equal(got[0], rg), first_diff(got[0], rg)
    assert (got == -80).sum() > 10000 and got.max() > 40


def test_c1_from_flow_samples_poses_and_grid(gpu, oracle, synth):
    """P0 + mapping in one call: poses bit-identical to the CPU statement of the P0 spec, grid bit-exact."""
    w = synth.CONFIGS["c1"]
    d = synth.generate(w)
    p = w.params()
    grids, px, py, st = gpu.replay_flow(p, d["t_ms"], d["of_rate_x"], d["of_rate_y"], d["h_m"], d["yaw_deg"], d["of_q"], d["ranges"])
    ox, oy = oracle.pose_integrate(d["t_ms"], d["of_rate_x"], d["of_rate_y"], d["h_m"], d["yaw_deg"], d["of_q"])
    assert np.array_equal(px.view(np.uint32), ox.view(np.uint32)) and np.array_equal(py.view(np.uint32), oy.view(np.uint32))
    want, U = oracle_grids(oracle, p, d, ox, oy)
    assert np.array_equal(grids, want), first_diff(grids, want)
    assert st["ray_cell_updates"] == U


def test_c3_drift_ensemble_scaled(gpu, oracle, synth):
    w = synth.scaled(synth.CONFIGS["c3"], n_flights=48, n_samples=1000)
    d = synth.generate(w)
    p = w.params()
    grids, px, py, st = gpu.replay_flow(p, d["t_ms"], d["of_rate_x"], d["of_rate_y"], d["h_m"], d["yaw_deg"], d["of_q"], d["ranges"])
    ox, oy = oracle.pose_integrate(d["t_ms"], d["of_rate_x"], d["of_rate_y"], d["h_m"], d["yaw_deg"], d["of_q"])
    assert np.array_equal(px.view(np.uint32), ox.view(np.uint32)) and np.array_equal(py.view(np.uint32), oy.view(np.uint32))
    want, U = oracle_grids(oracle, p, d, ox, oy)
    assert np.array_equal(grids, want), first_diff(grids, want)
    assert st["ray_cell_updates"] == U
    assert not np.array_equal(grids[0], grids[1])      # drift makes the members differ


def test_c2_long_log_fine_grid_scaled(gpu, oracle, orc_mod, synth):
    w = synth.scaled(synth.CONFIGS["c2"], n_samples=12000)
    d = synth.generate(w)
    p = w.params()
    got, st = gpu.replay(p, d["x_true"], d["y_true"], d["frame_yaw_deg"], d["ranges"])
    want, U = oracle_grids(oracle, p, d)
    assert np.array_equal(got, want), first_diff(got, want)
    assert st["ray_cell_updates"] == U
    if have_ref(orc_mod, 2000, 2000, "0.01"):
        ref = orc_mod.Reference(2000, 2000, "0.01")
        rg = ref.replay(d["x_true"][0, :3000], d["y_true"][0, :3000], d["frame_yaw_deg"][0, :3000], d["ranges"][0, :3000])
        g3, _ = gpu.replay(p, d["x_true"][:, :3000], d["y_true"][:, :3000], d["frame_yaw_deg"][:, :3000], d["ranges"][:, :3000])
        assert np.array_equal(g3[0], rg), first_diff(g3[0], rg)


def test_c4_multizone_sweep_scaled(gpu, oracle, synth):
    """64 beams per sample = two reference frames (yaw, yaw+45) at one pose, 16384^2 grid."""
    w = synth.scaled(synth.CONFIGS["c4"], n_samples=3000)
    d = synth.generate(w)
    p = w.params()
    x, y = synth.frame_poses(d, d["x_true"], d["y_true"])
    got, st = gpu.replay(p, x, y, d["frame_yaw_deg"], d["ranges"])
    want, U = oracle.replay_flights(p, x, y, d["frame_yaw_deg"], d["ranges"])
    assert np.array_equal(got, want), first_diff(got, want)
    assert st["ray_cell_updates"] == U and st["frames"] == 6000


@pytest.mark.parametrize("i_res,i_sigma", [(0, 3), (3, 0), (5, 5), (6, 15), (9, 8), (15, 15), (12, 1)])
def test_c5_resolution_noise_sweep_samples(gpu, oracle, orc_mod, synth, i_res, i_sigma):
    w = synth.c5_workload(i_res, i_sigma, n_flights=3, n_samples=700)
    d = synth.generate(w)
    p = w.params()
    got, st = gpu.replay(p, d["x_true"], d["y_true"], d["frame_yaw_deg"], d["ranges"])
    want, U = oracle_grids(oracle, p, d)
    assert np.array_equal(got, want), first_diff(got, want)
    assert st["ray_cell_updates"] == U
    if have_ref(orc_mod, w.W, w.W, w.res):
        ref = orc_mod.Reference(w.W, w.W, w.res)
        assert np.float32(ref.res) == np.float32(p.res_m)
        rg = ref.replay(d["x_true"][1], d["y_true"][1], d["frame_yaw_deg"][1], d["ranges"][1])
        assert np.array_equal(got[1], rg), first_diff(got[1], rg)


# ----------------------------------------------------------------------------------------------
# edge cases
# ----------------------------------------------------------------------------------------------
def test_ragged_and_degenerate_logs(gpu, oracle, synth, engine):
    """frame counts not a multiple of 32, a single frame, all-skipped frames, poses off the grid."""
    rng = np.random.default_rng(11)
    p = gpu.make_params(236, 236, 0.085, 20.0)
    for n in (1, 31, 33, 95, 1025):
        x = rng.uniform(-10.5, 10.5, (2, n)).astype(np.float32)
        y = rng.uniform(-10.5, 10.5, (2, n)).astype(np.float32)
        yaw = rng.uniform(-180, 180, (2, n)).astype(np.float32)
        r = rng.uniform(0, 4.4, (2, n, 32)).astype(np.float32)
        r[rng.random(r.shape) < 0.1] = np.nan
        r[1, ::3] = np.nan                                   # whole frames without a single return
        got, st = gpu.replay(p, x, y, yaw, r)
        want, U = oracle.replay_flights(p, x, y, yaw, r)
        assert np.array_equal(got, want), (n, first_diff(got, want))
        assert st["ray_cell_updates"] == U


def test_flight_without_any_return_stays_untouched(gpu, oracle, synth, engine):
    """a member whose every range is NaN touches nothing: zero grid in a fresh replay, unchanged when accumulating"""
    import torch
    w = synth.scaled(synth.CONFIGS["c3"], n_flights=3, n_samples=300)
    d = synth.generate(w)
    p = w.params()
    d["ranges"][1] = np.nan
    got, st = gpu.replay(p, d["x_true"], d["y_true"], d["frame_yaw_deg"], d["ranges"])
    want, U = oracle_grids(oracle, p, d)
    assert np.array_equal(got, want) and not got[1].any() and st["ray_cell_updates"] == U
    dev = torch.device("cuda:0")
    t = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (d["x_true"], d["y_true"], d["frame_yaw_deg"], d["ranges"])]
    g = torch.full((3, p.H, p.W), 17, dtype=torch.int8, device=dev)
    gpu.set_stream(torch.cuda.current_stream().cuda_stream)
    try:
        gpu.replay_dev(p, 3, 300, *(a.data_ptr() for a in t), g.data_ptr(), accumulate=True)
        torch.cuda.synchronize()
    finally:
        gpu.set_stream(None)
    out = g.cpu().numpy()
    assert (out[1] == 17).all() and (out[0] != 17).any()
    start = np.full((p.H, p.W), 17, np.int8)
    ref0, _ = oracle.replay(p, d["x_true"][0], d["y_true"][0], d["frame_yaw_deg"][0], d["ranges"][0], grid=start)
    assert np.array_equal(out[0], ref0)


def test_saturation_hazards_hover(gpu, oracle, engine):
    """a hovering drone next to a wall drives cells into both clamps with interleaved +6 / -1:
    the order-sensitive case of SURVEY 0.4 (accumulate-then-clamp gets these cells wrong)."""
    rng = np.random.default_rng(5)
    n = 4000
    p = gpu.make_params(400, 400, 0.05, 20.0)
    x = (0.02 * rng.standard_normal(n)).astype(np.float32)
    y = (0.02 * rng.standard_normal(n)).astype(np.float32)
    yaw = (rng.uniform(-6, 6, n)).astype(np.float32)
    r = (1.0 + 0.04 * rng.standard_normal((n, 32))).astype(np.float32)   # hits scatter over 2-3 cells
    got, st = gpu.replay(p, x[None], y[None], yaw[None], r[None])
    want, U = oracle.replay(p, x, y, yaw, r)
    assert np.array_equal(got[0], want), first_diff(got[0], want)
    # the order-free result (sum every cell's deltas, clamp once) differs on this input, so the test has teeth
    cells, origin = gpu.beam_cells(p, x, y, yaw, r)
    naive = _accumulate_then_clamp(p, origin, cells, r)
    assert (naive != want).sum() > 0, "accumulate-then-clamp happens to agree: the input no longer exercises update order"
    assert (want == 80).sum() > 0 and (want == -80).sum() > 0


def _accumulate_then_clamp(p, origin, cells, ranges):
    """What shared-memory atomics merged and clamped once would give (the north star's sketch): per cell the plain
    sum of all deltas, clamped at the end.  Cells from the closed form of the reference's walk."""
    n = origin.shape[0]
    ok = (cells[..., 0] >= 0) & (origin[:, None, 0] >= 0)
    x0 = np.broadcast_to(origin[:, None, 0], (n, 32))[ok].astype(np.int64)
    y0 = np.broadcast_to(origin[:, None, 1], (n, 32))[ok].astype(np.int64)
    dx, dy = cells[..., 0][ok] - x0, cells[..., 1][ok] - y0
    hit = (ranges < np.float32(p.max_range_m) - np.float32(p.hit_margin_m))[ok]
    adx, ady = np.abs(dx), np.abs(dy)
    xmaj = adx >= ady
    m, nn = np.where(xmaj, adx, ady), np.where(xmaj, ady, adx)
    acc = np.zeros(p.W * p.H, np.int64)
    for k in range(int(m.max()) + 1):
        live = k <= m
        q = np.where(m > 0, (k * nn + m // 2) // np.maximum(m, 1), 0)
        cx = x0 + np.sign(dx) * np.where(xmaj, k, q)
        cy = y0 + np.sign(dy) * np.where(xmaj, q, k)
        delta = np.where(k == m, np.where(hit, p.lo_occ, -(p.lo_free // 2)), -p.lo_free)
        np.add.at(acc, (cy * p.W + cx)[live], delta[live])
    return np.clip(acc, p.lo_min, p.lo_max).astype(np.int8).reshape(p.H, p.W)


def test_short_ranges_share_cells_inside_one_frame(gpu, oracle, engine):
    """ranges of a few cells: several beams of ONE frame end in / pass through the same cells."""
    rng = np.random.default_rng(9)
    n = 3000
    p = gpu.make_params(200, 200, 0.10, 20.0)
    x = rng.uniform(-1, 1, n).astype(np.float32)
    y = rng.uniform(-1, 1, n).astype(np.float32)
    yaw = rng.uniform(-180, 180, n).astype(np.float32)
    r = rng.uniform(0.051, 0.9, (n, 32)).astype(np.float32)
    got, _ = gpu.replay(p, x[None], y[None], yaw[None], r[None])
    want, _ = oracle.replay(p, x, y, yaw, r)
    assert np.array_equal(got[0], want), first_diff(got[0], want)


def test_other_log_odds_constants(gpu, oracle, synth, engine):
    """LO_FREE_DEC=3 makes the max-range end cell rule -(3/2) = -1 (integer division, uav_local_nav.c:266)."""
    w = synth.scaled(synth.CONFIGS["c1"], n_samples=800)
    d = synth.generate(w)
    p = w.params()
    p.lo_free, p.lo_occ, p.lo_min, p.lo_max = 3, 11, -100, 127
    got, _ = gpu.replay(p, d["x_true"], d["y_true"], d["frame_yaw_deg"], d["ranges"])
    want, _ = oracle_grids(oracle, p, d)
    assert np.array_equal(got, want), first_diff(got, want)


def test_bad_parameters_are_rejected(gpu):
    z = np.zeros((1, 4), np.float32)
    with pytest.raises(gpu.UqsError) as e:
        gpu.replay(gpu.make_params(400, 400, 0.0), z, z, z, np.ones((1, 4, 32), np.float32))
    assert e.value.code == gpu.ERR_BAD_ARG


def test_every_yaw_is_in_the_domain(gpu, oracle, engine):
    """|angle| >= 120 rad (glibc's reduce_large branch), Inf and NaN yaw: the reference maps them like any other
    frame (NaN end points land on the centre cell through lrintf's integer-indefinite result); so does the device."""
    rng = np.random.default_rng(21)
    n = 600
    p = gpu.make_params(400, 400, 0.05, 20.0)
    x = rng.uniform(-3, 3, n).astype(np.float32)
    y = rng.uniform(-3, 3, n).astype(np.float32)
    yaw = (rng.uniform(-1, 1, n) * 10.0 ** rng.uniform(3.5, 38, n)).astype(np.float32)     # 7e3 .. 3e38 degrees
    yaw[::50] = np.inf
    yaw[7::50] = -np.inf
    yaw[13::50] = np.nan
    r = rng.uniform(0.2, 4.4, (n, 32)).astype(np.float32)
    got, st = gpu.replay(p, x[None], y[None], yaw[None], r[None])
    want, U = oracle.replay(p, x, y, yaw, r)
    assert np.array_equal(got[0], want), first_diff(got[0], want)
    assert st["ray_cell_updates"] == U and st["domain_errors"] == 0
    cells, origin = gpu.beam_cells(p, x, y, yaw, r)
    wc, wo = oracle.beam_cells(p, x, y, yaw, r)
    assert np.array_equal(cells, wc) and np.array_equal(origin, wo)


def test_device_sincosf_large_and_nonfinite_arguments(gpu, oracle):
    """strided sweep of every binade from 120 to FLT_MAX, Inf and NaN payloads (exhaustive: test_gpu_sweeps.py)."""
    bits = np.concatenate([np.arange(0x42F00000 - 4096, 0x7F800000, 257, dtype=np.uint32),
                           np.arange(0x7F800000 - 70000, 0x7F800000 + 70000, dtype=np.uint32),
                           np.arange(0x7F800000, 0x80000000, 4099, dtype=np.uint32)])
    bits = np.concatenate([bits, bits | np.uint32(0x80000000)])
    a = bits.view(np.float32)
    ds, dc = gpu.sincosf_batch(a)
    hs, hc = oracle.libm_sincosf(a)
    bad = np.flatnonzero((ds.view(np.uint32) != hs.view(np.uint32)) | (dc.view(np.uint32) != hc.view(np.uint32)))
    assert bad.size == 0, f"{bad.size} mismatches, first bits={bits[bad[0]]:#x}"


def test_pose_integration_nonfinite_yaw_follows_the_cpu_statement(gpu, oracle, synth):
    w = synth.scaled(synth.CONFIGS["c1"], n_samples=500)
    d = synth.generate(w)
    yaw = d["yaw_deg"].copy()
    yaw[0, 100] = np.float32(1e9)
    yaw[0, 300] = np.inf                      # libm gives NaN: every later pose is NaN
    args = (d["t_ms"], d["of_rate_x"], d["of_rate_y"], d["h_m"], yaw, d["of_q"])
    px, py = gpu.pose_integrate(*args, mode=0)
    ox, oy = oracle.pose_integrate(*args)
    assert np.array_equal(np.isnan(px), np.isnan(ox)) and np.isnan(px[0, 300:]).all()
    assert np.array_equal(px[0, :300].view(np.uint32), ox[0, :300].view(np.uint32))
    assert np.array_equal(py[0, :300].view(np.uint32), oy[0, :300].view(np.uint32))


# ----------------------------------------------------------------------------------------------
# other input / output forms of the host-buffer call: u16 millimetre ranges, boxed grids
# ----------------------------------------------------------------------------------------------
def test_millimetre_ranges_and_boxed_output(gpu, oracle, synth):
    """uqs_replay_flow_mm (ranges as the sensor's u16 millimetres, converted on the device as uav_local_nav.c:1328
    does) and uqs_replay_flow_boxed (touched boxes instead of dense grids) give the bytes of uqs_replay_flow."""
    w = synth.on_mm_lattice(synth.scaled(synth.CONFIGS["c3"], n_flights=40, n_samples=700))
    d = synth.generate(w)
    p = w.params()
    d["ranges"][3] = np.nan                                       # a flight without any return: an empty box
    mm = synth.ranges_to_mm(d["ranges"])
    assert np.array_equal(synth.mm_to_ranges(mm).view(np.uint32), d["ranges"].view(np.uint32))
    args = (d["t_ms"], d["of_rate_x"], d["of_rate_y"], d["h_m"], d["yaw_deg"], d["of_q"])
    g_f, px, py, st_f = gpu.replay_flow(p, *args, d["ranges"])
    ox, oy = oracle.pose_integrate(*args)
    want, U = oracle.replay_flights(p, ox, oy, d["frame_yaw_deg"], d["ranges"])
    assert np.array_equal(g_f, want), first_diff(g_f, want)
    g_mm, qx, qy, st_mm = gpu.replay_flow_mm(p, *args, mm)
    assert np.array_equal(g_mm, want), first_diff(g_mm, want)
    assert st_mm == st_f and st_f["ray_cell_updates"] == U
    assert np.array_equal(qx.view(np.uint32), ox.view(np.uint32))
    cells = 40 * p.W * p.H
    for kw in ({"ranges_mm": mm}, {"ranges": d["ranges"]}):
        for chunk in (0, 7):                                      # one chunk / six chunks through the pipeline
            gpu.set_host_chunk(chunk)
            try:
                boxes, offs, packed, used, st_b = gpu.replay_flow_boxed(p, *args, **kw)
            finally:
                gpu.set_host_chunk(0)
            dense = gpu.unpack_boxed(p, boxes, offs, packed)
            assert np.array_equal(dense, want), first_diff(dense, want)
            assert st_b["ray_cell_updates"] == U and used < 0.6 * cells
            assert (boxes[3] == 0).all() and (boxes[0, 2] > boxes[0, 0]) and (boxes[:, 2] <= p.W).all()
    # the sub-tile engine has no touched boxes: the box is the whole grid, the bytes are the same
    gpu.set_engine(1, 0)
    try:
        boxes, offs, packed, used, _ = gpu.replay_flow_boxed(p, *args, ranges_mm=mm)
    finally:
        gpu.set_engine(0, 0)
    assert np.array_equal(gpu.unpack_boxed(p, boxes, offs, packed), want)
    # too small an output buffer is an error, never a truncated result
    with pytest.raises(gpu.UqsError) as e:
        gpu.replay_flow_boxed(p, *args, ranges_mm=mm, packed=np.empty(1000, np.int8))
    assert e.value.code == gpu.ERR_NOMEM


# ----------------------------------------------------------------------------------------------
# inputs outside the fast engines' assumptions: the unrestricted kernel (uqs_generic.cu)
# ----------------------------------------------------------------------------------------------
def test_unrestricted_kernel_equals_oracle_on_ordinary_logs(gpu, oracle, synth):
    w = synth.scaled(synth.CONFIGS["c3"], n_flights=5, n_samples=500)
    d = synth.generate(w)
    p = w.params()
    gpu.set_engine(3, 0)
    try:
        got, st = gpu.replay(p, d["x_true"], d["y_true"], d["frame_yaw_deg"], d["ranges"])
    finally:
        gpu.set_engine(0, 0)
    want, U = oracle_grids(oracle, p, d)
    assert np.array_equal(got, want), first_diff(got, want)
    assert st["ray_cell_updates"] == U and st["rays_accepted"] + st["rays_skipped"] == 5 * 500 * 32


def test_rays_longer_than_1024_cells(gpu, oracle):
    """3 mm cells: a 4 m ray is 1333 cells, beyond the fast engines' divide-free step -- routed, not refused."""
    rng = np.random.default_rng(31)
    n = 150
    p = gpu.make_params(3000, 3000, 0.003, 9.0)
    x = rng.uniform(-0.3, 0.3, n).astype(np.float32)
    y = rng.uniform(-0.3, 0.3, n).astype(np.float32)
    yaw = rng.uniform(-180, 180, n).astype(np.float32)
    r = rng.uniform(0.06, 4.3, (n, 32)).astype(np.float32)
    r[rng.random((n, 32)) < 0.05] = np.nan
    got, st = gpu.replay(p, x[None], y[None], yaw[None], r[None])
    want, U = oracle.replay(p, x, y, yaw, r)
    assert np.array_equal(got[0], want), first_diff(got[0], want)
    assert st["ray_cell_updates"] == U and st["domain_errors"] == 0
    wc, wo = oracle.beam_cells(p, x, y, yaw, r)
    ok = wc[..., 0] >= 0
    reach = np.maximum(np.abs(wc[..., 0] - wo[:, None, 0]), np.abs(wc[..., 1] - wo[:, None, 1]))[ok]
    assert reach.max() > 1100                                       # the input does contain such rays


def test_clamp_range_excluding_zero_and_out_of_range_start_grid(gpu, oracle, synth):
    import torch
    w = synth.scaled(synth.CONFIGS["c1"], n_samples=400)
    d = synth.generate(w)
    for lo_min, lo_max in ((3, 60), (-90, -5)):
        p = w.params()
        p.lo_min, p.lo_max = lo_min, lo_max
        got, _ = gpu.replay(p, d["x_true"], d["y_true"], d["frame_yaw_deg"], d["ranges"])
        want, _ = oracle_grids(oracle, p, d)
        assert np.array_equal(got, want), ((lo_min, lo_max), first_diff(got, want))
        assert (got == 0).any()                                    # untouched cells keep the 0 the reference leaves
    # accumulate into a grid holding values above lo_max: the reference clamps such a cell on its next update
    p = w.params()
    start = np.zeros((p.H, p.W), np.int8)
    start[150:250, 150:250] = 120
    start[::7, ::5] = -128
    want, _ = oracle.replay(p, d["x_true"][0], d["y_true"][0], d["frame_yaw_deg"][0], d["ranges"][0], grid=start.copy())
    dev = torch.device("cuda:0")
    t = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (d["x_true"], d["y_true"], d["frame_yaw_deg"], d["ranges"])]
    g = torch.from_numpy(start.copy()).to(dev)
    gpu.set_stream(torch.cuda.current_stream().cuda_stream)
    try:
        gpu.replay_dev(p, 1, w.n_frames, *(a.data_ptr() for a in t), g.data_ptr(), accumulate=True)
        torch.cuda.synchronize()
    finally:
        gpu.set_stream(None)
    out = g.cpu().numpy()
    assert np.array_equal(out, want), first_diff(out, want)
    assert (out == 120).any() and (out == 80).any()


# ----------------------------------------------------------------------------------------------
# size-independent properties (also hold at full BASELINE sizes -- see test_gpu_fullsize.py)
# ----------------------------------------------------------------------------------------------
def test_result_is_independent_of_subtile_size(gpu, oracle, synth):
    w = synth.scaled(synth.CONFIGS["c3"], n_flights=6, n_samples=900)
    d = synth.generate(w)
    p = w.params()
    want, _ = oracle_grids(oracle, p, d)
    gpu.set_engine(1, 0)
    try:
        for sw, sh in [(0, 0), (32, 32), (64, 48), (100, 100), (400, 20), (52, 200), (7, 13)]:
            gpu.set_tuning(sw, sh, 0)
            got, _ = gpu.replay(p, d["x_true"], d["y_true"], d["frame_yaw_deg"], d["ranges"])
            assert np.array_equal(got, want), ((sw, sh), first_diff(got, want))
    finally:
        gpu.set_tuning(0, 0, 0)
        gpu.set_engine(0, 0)


def test_time_slices_compose_exactly(gpu, oracle, synth):
    """A log cut into S concurrent time slices (slice 0 on values, later slices as clamp-add maps, composed
    per cell afterwards) must give the same bytes as the sequential replay, for any S and any tile size."""
    w = synth.scaled(synth.CONFIGS["c2"], n_samples=9000)
    d = synth.generate(w)
    p = w.params()
    want, U = oracle_grids(oracle, p, d)
    w1 = synth.scaled(synth.CONFIGS["c3"], n_flights=3, n_samples=1300)
    d1 = synth.generate(w1)
    p1 = w1.params()
    want1, _ = oracle_grids(oracle, p1, d1)
    gpu.set_engine(1, 0)
    try:
        for (sw, sh, S) in [(0, 0, 2), (0, 0, 5), (64, 24, 9), (0, 0, 64), (80, 80, 3)]:
            gpu.set_tuning(sw, sh, S)
            got, st = gpu.replay(p, d["x_true"], d["y_true"], d["frame_yaw_deg"], d["ranges"])
            assert np.array_equal(got, want), ((sw, sh, S), first_diff(got, want))
            assert st["ray_cell_updates"] == U
            got1, _ = gpu.replay(p1, d1["x_true"], d1["y_true"], d1["frame_yaw_deg"], d1["ranges"])
            assert np.array_equal(got1, want1), ((sw, sh, S), first_diff(got1, want1))
    finally:
        gpu.set_tuning(0, 0, 0)
        gpu.set_engine(0, 0)


def test_chained_replays_and_row_bands_via_device_api(gpu, oracle, synth):
    """accumulate: first half then second half == whole log; row bands owned by different 'GPUs'
    union to the whole grid (the config-4 partitioning), all through uqs_replay_dev on device pointers."""
    import torch
    sh = __import__("importlib").import_module("micro-quad-slam_b200.sharding")
    w = synth.scaled(synth.CONFIGS["c1"], n_samples=1400)
    d = synth.generate(w)
    p = w.params()
    want, _ = oracle_grids(oracle, p, d)
    dev = torch.device("cuda:0")
    tx, ty, tyaw = (torch.from_numpy(d[k]).to(dev) for k in ("x_true", "y_true", "frame_yaw_deg"))
    tr = torch.from_numpy(d["ranges"]).to(dev)
    gpu.set_stream(torch.cuda.current_stream().cuda_stream)
    try:
        g = torch.full((1, p.H, p.W), 55, dtype=torch.int8, device=dev)
        st = gpu.replay_dev(p, 1, 1400, tx.data_ptr(), ty.data_ptr(), tyaw.data_ptr(), tr.data_ptr(), g.data_ptr(), want_stats=True)
        torch.cuda.synchronize()
        assert np.array_equal(g.cpu().numpy(), want)
        # chained halves
        g2 = torch.zeros((1, p.H, p.W), dtype=torch.int8, device=dev)
        h = 700
        gpu.replay_dev(p, 1, h, tx.data_ptr(), ty.data_ptr(), tyaw.data_ptr(), tr.data_ptr(), g2.data_ptr(), accumulate=True)
        gpu.replay_dev(p, 1, 1400 - h, tx[:, h:].contiguous().data_ptr(), ty[:, h:].contiguous().data_ptr(),
                       tyaw[:, h:].contiguous().data_ptr(), tr[:, h:].contiguous().data_ptr(), g2.data_ptr(), accumulate=True)
        torch.cuda.synchronize()
        assert np.array_equal(g2.cpu().numpy(), want)
        # row bands
        g3 = torch.full((1, p.H, p.W), -7, dtype=torch.int8, device=dev)
        for r in range(3):
            r0, rows = sh.row_band(p.H, r, 3)
            gpu.replay_dev(p, 1, 1400, tx.data_ptr(), ty.data_ptr(), tyaw.data_ptr(), tr.data_ptr(), g3.data_ptr(), row0=r0, rows=rows)
        torch.cuda.synchronize()
        assert np.array_equal(g3.cpu().numpy(), want)
    finally:
        gpu.set_stream(None)


def test_host_pipeline_chunking_does_not_change_results(gpu, oracle, synth):
    """uqs_replay / uqs_replay_flow overlap H2D, kernels and D2H over chunks of flights: any chunk size
    (incl. a ragged last chunk) must give the same grids, poses and update count."""
    w = synth.scaled(synth.CONFIGS["c3"], n_flights=11, n_samples=500)
    d = synth.generate(w)
    p = w.params()
    args = (d["t_ms"], d["of_rate_x"], d["of_rate_y"], d["h_m"], d["yaw_deg"], d["of_q"])
    ox, oy = oracle.pose_integrate(*args)
    want, U = oracle_grids(oracle, p, d, ox, oy)
    try:
        for chunk in (0, 1, 3, 4, 11, 64):
            gpu.set_host_chunk(chunk)
            grids, px, py, st = gpu.replay_flow(p, *args, d["ranges"])
            assert np.array_equal(grids, want), (chunk, first_diff(grids, want))
            assert np.array_equal(px.view(np.uint32), ox.view(np.uint32)) and st["ray_cell_updates"] == U
            g2, st2 = gpu.replay(p, ox, oy, d["frame_yaw_deg"], d["ranges"])
            assert np.array_equal(g2, want) and st2["ray_cell_updates"] == U
    finally:
        gpu.set_host_chunk(0)


def test_unaligned_range_log_takes_the_plain_load_path(gpu, oracle, synth):
    """k_ray_setup stages a block's range readings with one TMA bulk copy, which needs a 16-byte aligned log; a
    log that starts 4 bytes off must give the same grids through the plain-load path (both engines)."""
    import torch
    w = synth.scaled(synth.CONFIGS["c3"], n_flights=3, n_samples=333)
    d = synth.generate(w)
    p = w.params()
    x, y = synth.frame_poses(d, d["x_true"], d["y_true"])
    want, U = oracle_grids(oracle, p, d, x, y)
    dev = torch.device("cuda:0")
    gpu.set_stream(torch.cuda.current_stream().cuda_stream)
    try:
        tx, ty, tyaw = (torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (x, y, d["frame_yaw_deg"]))
        flat = torch.from_numpy(np.ascontiguousarray(d["ranges"])).to(dev).reshape(-1)
        for shift in (0, 1, 2, 3):                       # floats: 0 -> aligned (TMA), 1..3 -> 4/8/12 bytes off
            buf = torch.empty(flat.numel() + 4, dtype=torch.float32, device=dev)
            view = buf[shift:shift + flat.numel()]
            view.copy_(flat)
            assert (view.data_ptr() % 16 == 0) == (shift == 0)
            for engine in (1, 2):
                gpu.set_engine(engine, 0)
                g = torch.zeros((w.n_flights, p.H, p.W), dtype=torch.int8, device=dev)
                st = gpu.replay_dev(p, w.n_flights, w.n_frames, tx.data_ptr(), ty.data_ptr(), tyaw.data_ptr(), view.data_ptr(),
                                    g.data_ptr(), want_stats=True)
                got = g.cpu().numpy()
                assert np.array_equal(got, want), (shift, engine, first_diff(got, want))
                assert st["ray_cell_updates"] == U
    finally:
        gpu.set_engine(0, 0)
        gpu.set_stream(None)


def _walk_cells(dx, dy):
    """Cells of the reference's Bresenham walk from (0,0) to (dx,dy) (uav_local_nav.c:246-274), in order."""
    x = y = 0
    sx = 1 if dx > 0 else -1
    sy = 1 if dy > 0 else -1
    ax, ay = abs(dx), -abs(dy)
    err = ax + ay
    out = [(0, 0)]
    while (x, y) != (dx, dy):
        e2 = 2 * err
        if e2 >= ay:
            err += ay; x += sx
        if e2 <= ax:
            err += ax; y += sy
        out.append((x, y))
    return out


def test_collision_bound_k0_is_safe_by_brute_force(gpu, synth):
    """The resident engine trusts K0 (k_ray_setup): two beams of a frame may share a cell only at steps k < K0, and
    frames flagged 'sorted' have their beams in circular angular order.  Both are checked here by brute force on the
    reference's own walk, for ordinary frames and for hostile ones (very short rays, narrow and wide fans, coarse and
    fine cells, dropouts), independently of any replay result."""
    from fractions import Fraction
    rng = np.random.default_rng(77)
    checked = shared_below = 0
    for W, res, fov, rmax in ((400, 0.05, 63.0, 4.0), (200, 0.10, 63.0, 4.0), (800, 0.025, 63.0, 4.0), (400, 0.05, 2.0, 4.0),
                              (400, 0.05, 170.0, 4.0), (400, 0.05, 89.0, 0.6), (300, 0.07, 120.0, 2.0)):
        p = gpu.make_params(W, W, res)
        p.fov_deg = fov
        p.max_range_m = rmax
        N = 100
        x = rng.uniform(-2, 2, N).astype(np.float32)
        y = rng.uniform(-2, 2, N).astype(np.float32)
        yaw = rng.uniform(-180, 180, N).astype(np.float32)
        ranges = rng.uniform(0.06, rmax * 1.1, (N, 32)).astype(np.float32)
        ranges[rng.random((N, 32)) < 0.1] = np.nan                        # dropouts
        short = rng.random(N) < 0.3
        ranges[short] = rng.uniform(0.06, 6 * res, (int(short.sum()), 32)).astype(np.float32)   # rays of a few cells
        cells, origin = gpu.beam_cells(p, x, y, yaw, ranges)
        k0, srt = gpu.frame_bounds(p, x, y, yaw, ranges)
        for f in range(N):
            if origin[f, 0] < 0:
                assert k0[f] == -1
                continue
            beams = [(b, int(cells[f, b, 0] - origin[f, 0]), int(cells[f, b, 1] - origin[f, 1])) for b in range(32) if cells[f, b, 0] >= 0]
            walks = {b: _walk_cells(dx, dy) for b, dx, dy in beams}
            last_shared = -1
            for i, (bi, _, _) in enumerate(beams):
                for bj, _, _ in beams[i + 1:]:
                    wi, wj = walks[bi], walks[bj]
                    for k in range(min(len(wi), len(wj))):
                        if wi[k] == wj[k]:
                            last_shared = max(last_shared, k)
            assert last_shared < k0[f], (W, fov, f, last_shared, int(k0[f]))
            checked += 1
            shared_below += last_shared >= 1
            if srt[f]:
                # circular order of the directions: exact position on the unit max-norm ring as a fraction in [0, 8)
                def sigma(dx, dy):
                    ax, ay = abs(dx), abs(dy)
                    if ax >= ay:
                        t = Fraction(ay, ax)
                        return (t if dy >= 0 else 8 - t) if dx > 0 else (4 - t if dy >= 0 else 4 + t)
                    t = Fraction(ax, ay)
                    return (2 - t if dx >= 0 else 2 + t) if dy > 0 else (6 + t if dx >= 0 else 6 - t)
                sig = [sigma(dx, dy) % 8 for _, dx, dy in beams if (dx, dy) != (0, 0)]
                wraps = sum(1 for a, b in zip(sig, sig[1:] + sig[:1]) if b < a)
                assert wraps <= 1, (W, fov, f, [float(v) for v in sig])
    assert checked > 500 and shared_below > 250


# ----------------------------------------------------------------------------------------------
# P0
# ----------------------------------------------------------------------------------------------
def test_pose_integration_exact_order_is_bit_identical(gpu, oracle, synth):
    w = synth.scaled(synth.CONFIGS["c3"], n_flights=33, n_samples=3000)
    d = synth.generate(w)
    d["of_rate_x"][3, 100:120] = np.nan          # NaN inputs contribute nothing
    d["of_q"][5, :500] = 10                      # low quality gated out
    args = (d["t_ms"], d["of_rate_x"], d["of_rate_y"], d["h_m"], d["yaw_deg"], d["of_q"])
    px, py = gpu.pose_integrate(*args, mode=0)
    ox, oy = oracle.pose_integrate(*args)
    assert np.array_equal(px.view(np.uint32), ox.view(np.uint32))
    assert np.array_equal(py.view(np.uint32), oy.view(np.uint32))
    assert np.abs(ox).max() > 1.0


def test_pose_integration_long_log_exact(gpu, oracle, synth):
    w = synth.scaled(synth.CONFIGS["c2"], n_samples=100001)
    d = synth.generate(w)
    args = (d["t_ms"], d["of_rate_x"], d["of_rate_y"], d["h_m"], d["yaw_deg"], d["of_q"])
    px, py = gpu.pose_integrate(*args, mode=0)
    ox, oy = oracle.pose_integrate(*args)
    assert np.array_equal(px.view(np.uint32), ox.view(np.uint32)) and np.array_equal(py.view(np.uint32), oy.view(np.uint32))


def test_pose_scan_variant_on_the_one_hour_log(gpu, oracle, synth):
    """North-star tolerance for the look-back scan (mode 1) on the config-2 log (360 000 samples): within 1e-6
    relative of a binary64 accumulation of the same binary32 increments, and in the same 1 cm grid cell as that
    accumulation for every sample.  The exact-order kernel (mode 0) is bit-identical to the binary32 serial sum
    instead (test above), which itself drifts from the binary64 sum by ~sqrt(n) * 2^-24 relative: the count of
    samples whose cell differs between the two modes is reported, not asserted to be zero."""
    w = synth.CONFIGS["c2"]
    d = synth.generate(w)
    args = (d["t_ms"], d["of_rate_x"], d["of_rate_y"], d["h_m"], d["yaw_deg"], d["of_q"])
    sx, sy = gpu.pose_integrate(*args, mode=1)
    rx, ry = oracle.pose_integrate_f64(*(a[0] for a in args))
    scale = max(np.abs(rx).max(), np.abs(ry).max())
    assert scale > 1.0
    err = max(np.abs(sx[0].astype(np.float64) - rx).max(), np.abs(sy[0].astype(np.float64) - ry).max())
    assert err <= 1e-6 * scale, f"scan differs from the binary64 sum by {err:.3e} m on a {scale:.2f} m path"
    p = w.params()

    def cells(x, y):      # world_to_grid (uav_local_nav.c:207-210) in binary32, as the reference computes it
        fx = (x.astype(np.float32) - np.float32(p.origin_x)) / np.float32(p.res_m)
        fy = (y.astype(np.float32) - np.float32(p.origin_y)) / np.float32(p.res_m)
        return np.rint(fx).astype(np.int64) + p.W // 2, np.rint(fy).astype(np.int64) + p.H // 2

    cs, cr = cells(sx[0], sy[0]), cells(rx, ry)
    differ = int(((cs[0] != cr[0]) | (cs[1] != cr[1])).sum())
    assert differ == 0, f"{differ} of {rx.size} samples fall in another 1 cm cell than the binary64 sum"
    ex, ey = gpu.pose_integrate(*args, mode=0)
    ce = cells(ex[0], ey[0])
    drift = int(((cs[0] != ce[0]) | (cs[1] != ce[1])).sum())
    rel = max(np.abs(sx - ex).max(), np.abs(sy - ey).max()) / scale
    print(f"scan vs exact-order binary32 sum: max relative difference {rel:.2e}, {drift} of {rx.size} samples in a different cell")
    assert rel <= 64 * np.sqrt(rx.size) * 2.0 ** -24


# ----------------------------------------------------------------------------------------------
# drop-in symbols
# ----------------------------------------------------------------------------------------------
def test_dropin_symbols_replay_like_log_tick(gpu, oracle, orc_mod, synth):
    """Drive map_update_from_beams()/raycast_update()/world_to_grid() exactly as uav_local_nav.c does
    (HOVER init :2187-2194, log_tick :1629-1635) and compare occ_grid with the reference's."""
    w = synth.scaled(synth.CONFIGS["c1"], n_samples=600)
    d = synth.generate(w)
    p = w.params()
    di = gpu.DropIn(p)
    di.set_beams(d["ranges"][0, 0])
    di.map_update_from_beams(0.0, 0.0, 0.0)          # before map_inited: silent no-op (:281)
    assert not di.world_to_grid(0.0, 0.0)[0]          # :206
    ox, oy = float(d["x_true"][0, 0]), float(d["y_true"][0, 0])
    di.hover_init(ox, oy)
    q = p.copy()
    q.origin_x, q.origin_y = np.float32(ox), np.float32(oy)
    want = np.zeros((p.H, p.W), np.int8)
    for i in range(600):
        di.set_beams(d["ranges"][0, i])
        di.map_update_from_beams(d["x_true"][0, i], d["y_true"][0, i], d["yaw_deg"][0, i])
        oracle.L.orc_frame(ctypes.byref(q), want.ctypes.data, d["x_true"][0, i], d["y_true"][0, i], d["yaw_deg"][0, i],
                           np.ascontiguousarray(d["ranges"][0, i]).ctypes.data)
        if i % 97 == 0:
            a = (float(d["x_true"][0, i]), float(d["y_true"][0, i]), float(d["x_true"][0, i]) + 1.5, float(d["y_true"][0, i]) - 0.7)
            di.raycast_update(*a, True)
            oracle.raycast(q, want, *a, True)
        if i == 300:
            mid = di.grid()                            # sync-on-read in the middle of the log
            assert np.array_equal(mid, want), first_diff(mid, want)
    got = di.grid()
    assert np.array_equal(got, want), first_diff(got, want)
    for (x, y) in [(ox, oy), (ox + 9.99, oy - 9.99), (ox + 10.1, oy), (ox + 0.025, oy + 0.075)]:
        assert di.world_to_grid(x, y) == oracle.world_to_grid(q, x, y) or not oracle.world_to_grid(q, x, y)[0]
    if have_ref(orc_mod, 400, 400, "0.05"):
        ref = orc_mod.Reference(400, 400, "0.05")
        ref.reset(ox, oy)
        for i in range(600):
            ref.frame(d["x_true"][0, i], d["y_true"][0, i], d["yaw_deg"][0, i], d["ranges"][0, i])
            if i % 97 == 0:
                ref.raycast_update(float(d["x_true"][0, i]), float(d["y_true"][0, i]), float(d["x_true"][0, i]) + 1.5, float(d["y_true"][0, i]) - 0.7, True)
        assert np.array_equal(got, ref.grid())


def test_dropin_symbols_on_a_grid_wider_than_1025_cells(gpu, oracle):
    """raycast_update() takes arbitrary end points: on a grid wider than 1025 cells a single call may be a ray of more
    than 1024 cells, so the drop-in queue is replayed by the unrestricted kernel (raw-ray entries included); an
    edited occ_grid with values outside [-80, 80] is honoured the way the reference's clamp does."""
    p = gpu.make_params(1400, 1200, 0.01, 14.0)
    di = gpu.DropIn(p)
    di.hover_init(0.5, -0.25)
    q = p.copy()
    q.origin_x, q.origin_y = np.float32(0.5), np.float32(-0.25)
    want = np.zeros((p.H, p.W), np.int8)
    rng = np.random.default_rng(77)
    for i in range(120):
        x, y, yaw = (np.float32(v) for v in (0.5 + rng.uniform(-2, 2), -0.25 + rng.uniform(-2, 2), rng.uniform(-180, 180)))
        beams = rng.uniform(0.03, 4.4, 32).astype(np.float32)
        beams[rng.random(32) < 0.1] = np.nan
        di.set_beams(beams)
        di.map_update_from_beams(x, y, yaw)
        oracle.L.orc_frame(ctypes.byref(q), want.ctypes.data, x, y, yaw, np.ascontiguousarray(beams).ctypes.data)
        if i % 10 == 0:
            a = (float(x), float(y), float(np.float32(0.5 + rng.uniform(-6.9, 6.9))), float(np.float32(-0.25 + rng.uniform(-5.9, 5.9))))
            hit = bool(i % 20)
            di.raycast_update(*a, hit)                 # up to ~1380 cells long
            oracle.raycast(q, want, *a, hit)
    got = di.grid()
    assert np.array_equal(got, want), first_diff(got, want)
    assert (want != 0).sum() > 20000


def test_unrestricted_kernel_with_row_bands_and_chained_replays(gpu, oracle, synth):
    import torch
    w = synth.scaled(synth.CONFIGS["c3"], n_flights=3, n_samples=400)
    d = synth.generate(w)
    p = w.params()
    want, _ = oracle_grids(oracle, p, d)
    dev = torch.device("cuda:0")
    t = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (d["x_true"], d["y_true"], d["frame_yaw_deg"], d["ranges"])]
    g = torch.full((3, p.H, p.W), 7, dtype=torch.int8, device=dev)
    gpu.set_stream(torch.cuda.current_stream().cuda_stream)
    gpu.set_engine(3, 0)
    try:
        for r0, rows in ((0, 120), (120, 164), (284, 116)):                                 # three owned bands, fresh
            gpu.replay_dev(p, 3, 400, *(a.data_ptr() for a in t), g.data_ptr(), row0=r0, rows=rows)
        torch.cuda.synchronize()
        assert np.array_equal(g.cpu().numpy(), want)
        g.zero_()
        h = 173
        t1 = [a[:, :h].contiguous() for a in t]                                             # kept alive until the sync
        gpu.replay_dev(p, 3, h, *(a.data_ptr() for a in t1), g.data_ptr(), accumulate=True)
        t2 = [a[:, h:].contiguous() for a in t]
        gpu.replay_dev(p, 3, 400 - h, *(a.data_ptr() for a in t2), g.data_ptr(), accumulate=True)   # chained halves
        torch.cuda.synchronize()
        assert np.array_equal(g.cpu().numpy(), want)
    finally:
        gpu.set_engine(0, 0)
        gpu.set_stream(None)


# ----------------------------------------------------------------------------------------------
# randomized stress: odd geometries and sensor constants against the oracle, every engine
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("seed", range(40))
def test_random_geometry_and_sensor_constants(gpu, oracle, seed):
    """fov from 2 deg (beams almost parallel: collisions far out, large K0) to 170 deg (fans overlap: beams NOT in
    angular order, the all-pairs K0 / table path), coarse and fine cells, widths not divisible by 4, moved origins,
    short and long ranges, hovering and jumping poses -- sub-tile, time-sliced and resident engines must all equal
    the oracle."""
    rng = np.random.default_rng(1000 + seed)
    W = int(rng.choice([62, 90, 101, 128, 250, 333, 400]))
    H = int(rng.choice([58, 96, 127, 200, 260, 400]))
    res = float(rng.choice([0.02, 0.05, 0.1, 0.25]))
    p = gpu.make_params(W, H, res, W * res)
    p.fov_deg = np.float32(rng.choice([2.0, 10.0, 63.0, 90.0, 120.0, 170.0]))
    p.max_range_m = np.float32(rng.choice([1.0, 4.0, 6.0]))
    p.origin_x, p.origin_y = np.float32(rng.uniform(-1, 1)), np.float32(rng.uniform(-1, 1))
    p.lo_free, p.lo_occ = int(rng.choice([1, 2, 5])), int(rng.choice([3, 6, 20]))
    p.lo_min, p.lo_max = int(rng.choice([-80, -128, -5])), int(rng.choice([80, 127, 7]))
    F, N = int(rng.choice([1, 2, 5])), int(rng.choice([40, 333, 700]))
    half_x, half_y = 0.5 * W * res, 0.5 * H * res
    mode = rng.integers(3)
    if mode == 0:      # hover
        x = (p.origin_x + 0.01 * rng.standard_normal((F, N))).astype(np.float32)
        y = (p.origin_y + 0.01 * rng.standard_normal((F, N))).astype(np.float32)
    elif mode == 1:    # jumps all over (and beyond) the map
        x = rng.uniform(-1.2 * half_x, 1.2 * half_x, (F, N)).astype(np.float32)
        y = rng.uniform(-1.2 * half_y, 1.2 * half_y, (F, N)).astype(np.float32)
    else:              # smooth drift
        t = np.linspace(0, 1, N, dtype=np.float32)
        x = (np.float32(0.7 * half_x) * np.cos(6 * t) + np.zeros((F, 1), np.float32)).astype(np.float32)
        y = (np.float32(0.7 * half_y) * np.sin(4 * t) + np.zeros((F, 1), np.float32)).astype(np.float32)
    yaw = rng.uniform(-360, 360, (F, N)).astype(np.float32)
    r = rng.uniform(0.0, float(p.max_range_m) * 1.1, (F, N, 32)).astype(np.float32)
    r[rng.random(r.shape) < 0.1] = np.nan
    r[rng.random(r.shape) < 0.1] = np.float32(rng.uniform(0.04, 3 * res))       # very short rays
    want, U = oracle.replay_flights(p, x, y, yaw, r)
    cases = [(1, 0, 1, 0), (1, 0, 3, 0), (0, 0, 0, 0)]
    if max(W, H) <= 400:
        cases += [(2, 4, 0, 0), (2, 16, 0, 0), (2, 4, 0, 1), (2, 8, 0, 1), (2, 32, 0, 1), (2, 4, 0, 2), (2, 8, 0, 2), (2, 16, 0, 2)]
    try:
        for engine, nw, slices, fan in cases:                      # fan: 0 = 32-beam layout, 1 = fan layout, 2 = dedicated decode warp
            gpu.set_engine(engine, nw)
            gpu.set_fan_layout(1 if fan == 1 else 0)
            gpu.set_decode_warp(1 if fan == 2 else 0)
            gpu.set_tuning(0, 0, slices)
            got, st = gpu.replay(p, x, y, yaw, r)
            assert np.array_equal(got, want), ((engine, nw, slices, fan), dict(W=W, H=H, res=res, fov=float(p.fov_deg)), first_diff(got, want))
            assert st["ray_cell_updates"] == U
    finally:
        gpu.set_engine(0, 0)
        gpu.set_fan_layout(-1)
        gpu.set_decode_warp(-1)
        gpu.set_tuning(0, 0, 0)
