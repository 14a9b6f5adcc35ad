"""BASELINE.json's five configurations at FULL size.  The CPU oracle cannot replay these in seconds, so
parity is shown through size-independent properties -- both replay engines / several sub-tile sizes must
give identical bytes, a log replayed in two chained halves must equal the whole, owned row bands must union
to the whole grid -- plus byte comparison with the oracle on a random sample of flights or a log prefix."""
import importlib

import numpy as np
import pytest
import torch

from conftest import first_diff

pytestmark = pytest.mark.gpu


def _dev_inputs(d, x, y, dev):
    return tuple(torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (x, y, d["frame_yaw_deg"], d["ranges"]))


def _replay(gpu, p, F, N, t, grids, **kw):
    return gpu.replay_dev(p, F, N, t[0].data_ptr(), t[1].data_ptr(), t[2].data_ptr(), t[3].data_ptr(), grids.data_ptr(), **kw)


@pytest.fixture()
def torch_stream(gpu):
    gpu.set_stream(torch.cuda.current_stream().cuda_stream)
    yield torch.device("cuda:0")
    gpu.set_stream(None)
    gpu.set_engine(0, 0)
    gpu.set_tuning(0, 0, 0)


def test_c3_full_ensemble_4096_flights(gpu, oracle, synth, torch_stream):
    dev = torch_stream
    w = synth.CONFIGS["c3"]
    d = synth.generate(w)
    p = w.params()
    F, N = w.n_flights, w.n_frames
    args = [torch.from_numpy(d[k].view(np.int32) if k == "t_ms" else d[k]).to(dev) for k in ("t_ms", "of_rate_x", "of_rate_y", "h_m", "yaw_deg", "of_q")]
    tx = torch.empty((F, N), dtype=torch.float32, device=dev)
    ty = torch.empty_like(tx)
    gpu.pose_integrate_dev(F, N, *(a.data_ptr() for a in args), tx.data_ptr(), ty.data_ptr(), 0)
    t = (tx, ty, args[4], torch.from_numpy(d["ranges"]).to(dev))
    g1 = torch.empty((F, p.H, p.W), dtype=torch.int8, device=dev)
    g2 = torch.empty_like(g1)
    gpu.set_engine(1, 0)
    st = _replay(gpu, p, F, N, t, g1, want_stats=True)
    gpu.set_engine(2, 16)
    _replay(gpu, p, F, N, t, g2)
    torch.cuda.synchronize()
    assert torch.equal(g1, g2), "sub-tile and resident engines disagree at full size"
    gpu.set_engine(1, 0)
    gpu.set_tuning(56, 100, 0)
    _replay(gpu, p, F, N, t, g2)
    torch.cuda.synchronize()
    assert torch.equal(g1, g2), "result depends on the sub-tile size"
    assert st["frames"] == F * N and st["domain_errors"] == 0 and st["ray_cell_updates"] > 2e10
    # oracle on a random sample of ensemble members
    rng = np.random.default_rng(3)
    px, py = tx.cpu().numpy(), ty.cpu().numpy()
    for f in rng.choice(F, 6, replace=False):
        ox, oy = oracle.pose_integrate(d["t_ms"][f], d["of_rate_x"][f], d["of_rate_y"][f], d["h_m"][f], d["yaw_deg"][f], d["of_q"][f])
        assert np.array_equal(ox.view(np.uint32), px[f].view(np.uint32)) and np.array_equal(oy.view(np.uint32), py[f].view(np.uint32))
        want, _ = oracle.replay(p, ox, oy, d["yaw_deg"][f], d["ranges"][f])
        got = g1[f].cpu().numpy()
        assert np.array_equal(got, want), (int(f), first_diff(got, want))


def test_c2_full_one_hour_log(gpu, oracle, synth, torch_stream):
    dev = torch_stream
    w = synth.CONFIGS["c2"]
    d = synth.generate(w)
    p = w.params()
    N = w.n_frames
    px, py = gpu.pose_integrate(d["t_ms"], d["of_rate_x"], d["of_rate_y"], d["h_m"], d["yaw_deg"], d["of_q"], mode=0)
    ox, oy = oracle.pose_integrate(d["t_ms"], d["of_rate_x"], d["of_rate_y"], d["h_m"], d["yaw_deg"], d["of_q"])
    assert np.array_equal(px.view(np.uint32), ox.view(np.uint32)) and np.array_equal(py.view(np.uint32), oy.view(np.uint32))
    t = _dev_inputs(d, px, py, dev)
    g1 = torch.empty((1, p.H, p.W), dtype=torch.int8, device=dev)
    g2 = torch.zeros_like(g1)
    st = _replay(gpu, p, 1, N, t, g1, want_stats=True)
    # two chained halves == whole
    h = N // 2 + 7
    _replay(gpu, p, 1, h, t, g2, accumulate=True)
    t2 = tuple(a[:, h:].contiguous() for a in t)
    _replay(gpu, p, 1, N - h, t2, g2, accumulate=True)
    torch.cuda.synchronize()
    assert torch.equal(g1, g2), "chained halves differ from the whole log"
    gpu.set_tuning(48, 48, 0)
    _replay(gpu, p, 1, N, t, g2)
    torch.cuda.synchronize()
    assert torch.equal(g1, g2), "result depends on the sub-tile size"
    assert st["ray_cell_updates"] > 2.5e9
    # oracle on a 20 000-frame prefix
    n = 20000
    g3 = torch.empty_like(g1)
    _replay(gpu, p, 1, n, tuple(a[:, :n].contiguous() for a in t), g3)
    want, _ = oracle.replay(p, ox[0, :n], oy[0, :n], d["frame_yaw_deg"][0, :n], d["ranges"][0, :n])
    got = g3[0].cpu().numpy()
    assert np.array_equal(got, want), first_diff(got, want)


def test_c4_full_building_sweep_row_bands(gpu, oracle, synth, torch_stream):
    dev = torch_stream
    sh = importlib.import_module("micro-quad-slam_b200.sharding")
    w = synth.CONFIGS["c4"]
    d = synth.generate(w)
    p = w.params()
    N = w.n_frames
    x, y = synth.frame_poses(d, d["x_true"], d["y_true"])
    t = _dev_inputs(d, x, y, dev)
    g1 = torch.empty((1, p.H, p.W), dtype=torch.int8, device=dev)
    st = _replay(gpu, p, 1, N, t, g1, want_stats=True)
    assert st["frames"] == 2 * 1048576 and st["domain_errors"] == 0
    # eight owned row bands (the 8-GPU partitioning, executed band after band on one GPU) union to the whole grid
    g2 = torch.full_like(g1, 9)
    for r in range(8):
        r0, rows = sh.row_band(p.H, r, 8)
        _replay(gpu, p, 1, N, t, g2, row0=r0, rows=rows)
    torch.cuda.synchronize()
    assert torch.equal(g1, g2), "union of owned row bands differs from the whole grid"
    # the balanced cuts of a 2/4/8-rank run (equal shares of the log, not of the rows): same union, and every band carries work
    for world in (2, 4, 8):
        edges = gpu.balanced_row_bands_dev(p, N, t[0].data_ptr(), t[1].data_ptr(), world)
        assert edges[0] == 0 and edges[-1] == p.H and all(b > a and a % 4 == 0 for a, b in zip(edges, edges[1:]))
        if world == 8:
            equal = [sh.row_band(p.H, r, 8)[0] for r in range(8)] + [p.H]
            assert edges != equal                          # the sweep covers the middle of the grid: equal bands are not balanced
            g2.fill_(9)
            for a, b in zip(edges, edges[1:]):
                _replay(gpu, p, 1, N, t, g2, row0=a, rows=b - a)
            torch.cuda.synchronize()
            assert torch.equal(g1, g2), "union of balanced row bands differs from the whole grid"
    # oracle on an 8 000-frame prefix
    n = 8000
    g3 = torch.empty_like(g1)
    _replay(gpu, p, 1, n, tuple(a[:, :n].contiguous() for a in t), g3)
    want, _ = oracle.replay(p, x[0, :n], y[0, :n], d["frame_yaw_deg"][0, :n], d["ranges"][0, :n])
    got = g3[0].cpu().numpy()
    assert np.array_equal(got, want), first_diff(got, want)


def test_c5_full_sweep_256_configs_x_64_flights(gpu, oracle, orc_mod, synth, torch_stream):
    """16 resolutions x 16 range noises x 64 flights; per config both engines agree (where the grid fits a CTA)
    and one flight per config row is compared with the oracle (the reference's own code where built)."""
    dev = torch_stream
    rng = np.random.default_rng(5)
    total_U = 0
    for i_res in range(16):
        check_sigma = int(rng.integers(16))
        for i_sigma in range(16):
            w = synth.c5_workload(i_res, i_sigma)
            d = synth.generate(w)
            p = w.params()
            t = _dev_inputs(d, d["x_true"], d["y_true"], dev)
            g1 = torch.empty((w.n_flights, p.H, p.W), dtype=torch.int8, device=dev)
            gpu.set_engine(1, 0)
            st = _replay(gpu, p, w.n_flights, w.n_frames, t, g1, want_stats=True)
            total_U += st["ray_cell_updates"]
            if p.W <= 476 and i_sigma % 4 == 0:
                g2 = torch.empty_like(g1)
                gpu.set_engine(2, 16)
                _replay(gpu, p, w.n_flights, w.n_frames, t, g2)
                torch.cuda.synchronize()
                assert torch.equal(g1, g2), (w.name, "engines disagree")
            if i_sigma == check_sigma:
                f = int(rng.integers(w.n_flights))
                want, _ = oracle.replay(p, d["x_true"][f], d["y_true"][f], d["frame_yaw_deg"][f], d["ranges"][f])
                got = g1[f].cpu().numpy()
                assert np.array_equal(got, want), (w.name, f, first_diff(got, want))
    assert total_U > 5e10
