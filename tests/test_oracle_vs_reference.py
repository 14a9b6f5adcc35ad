"""Pin the CPU restatement (oracle/uqs_oracle.c) against the reference's OWN mapping code
(oracle/_ref, extracted from uav_local_nav.c:181-385 by oracle/build_ref.sh).

The reference has no tests or golden vectors for this path; these
comparisons, plus the Appendix-D known answer, are the pin."""
import os

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from conftest import first_diff, have_ref

GEOMS = [(500, "0.10", 50.0), (400, "0.05", 20.0), (2000, "0.01", 20.0), (668, "0.03", 20.0), (236, "0.085", 20.0)]


def _need_ref(orc_mod, W, res):
    if not have_ref(orc_mod, W, W, res):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    return orc_mod.Reference(W, W, res)


def kat_inputs():
    """SURVEY.md Appendix D probe: float arithmetic exactly as the C probe wrote it."""
    f = np.float32
    b = np.zeros((4, 8), f)
    for d in range(4):
        for c in range(8):
            b[d, c] = f(1.0) + f(0.1) * f(c) + f(0.5) * f(d)
    b[2, 3] = np.nan
    b[1, 1] = 4.5
    b[0, 0] = 3.97
    k = np.arange(100, dtype=f)
    return f(0.01) * k, f(-0.02) * k, f(3.0) * k, np.broadcast_to(b.reshape(1, 32), (100, 32)).copy()


def test_reference_known_answer(orc_mod, oracle):
    """sum=-40173 nz=4658 min=-80 max=42 fnv1a32=c4787900 at the reference's native 500x500 @ 0.10."""
    ref = _need_ref(orc_mod, 500, "0.10")
    x, y, yaw, r = kat_inputs()
    g = ref.replay(x, y, yaw, r)
    assert (int(g.sum(dtype=np.int64)), int(np.count_nonzero(g)), int(g.min()), int(g.max())) == (-40173, 4658, -80, 42)
    assert oracle.fnv1a32(g) == 0xC4787900


def test_oracle_known_answer(pkg, oracle):
    """The restatement reproduces the same known answer without the reference present."""
    p = pkg.make_params(500, 500, 0.10, 50.0)
    x, y, yaw, r = kat_inputs()
    g, U = oracle.replay(p, x, y, yaw, r)
    assert (int(g.sum(dtype=np.int64)), int(np.count_nonzero(g)), int(g.min()), int(g.max())) == (-40173, 4658, -80, 42)
    assert oracle.fnv1a32(g) == 0xC4787900
    assert U > 0


@pytest.mark.parametrize("W,res,size", GEOMS)
def test_oracle_equals_reference_on_synthetic_flight(orc_mod, oracle, pkg, synth, W, res, size):
    ref = _need_ref(orc_mod, W, res)
    n = 1500 if W < 1000 else 400
    w = synth.scaled(synth.Workload("t", 77, 1, n, W, res, size, 50.0), n_samples=n)
    d = synth.generate(w)
    p = w.params()
    assert np.float32(ref.res) == np.float32(p.res_m)
    x, y, yaw, r = d["x_true"][0], d["y_true"][0], d["yaw_deg"][0], d["ranges"][0]
    want = ref.replay(x, y, yaw, r)
    got, U = oracle.replay(p, x, y, yaw, r)
    assert np.array_equal(got, want), first_diff(got, want)
    assert not ref.recentered()
    assert U > n


def test_oracle_equals_reference_with_moved_origin(orc_mod, oracle, pkg, synth):
    ref = _need_ref(orc_mod, 400, "0.05")
    w = synth.scaled(synth.CONFIGS["c1"], n_samples=500)
    d = synth.generate(w)
    p = w.params()
    p.origin_x, p.origin_y = np.float32(1.2345), np.float32(-0.777)
    x, y, yaw, r = d["x_true"][0], d["y_true"][0], d["yaw_deg"][0], d["ranges"][0]
    want = ref.replay(x, y, yaw, r, ox=p.origin_x, oy=p.origin_y)
    got, _ = oracle.replay(p, x, y, yaw, r)
    assert np.array_equal(got, want), first_diff(got, want)


finite = st.floats(width=32, allow_nan=False, allow_infinity=False, min_value=-30, max_value=30)
ranges_st = st.lists(st.one_of(st.floats(width=32, min_value=0.0, max_value=5.0), st.just(float("nan")),
                               st.sampled_from([0.05, 0.0500001, 3.95, 3.9499998, 4.0, 4.0000005, 0.02])),
                     min_size=32, max_size=32)


@settings(max_examples=300, deadline=None)
@given(x=finite, y=finite, yaw=st.floats(width=32, min_value=-720, max_value=720), r=ranges_st)
def test_oracle_equals_reference_hypothesis_frames(orc_mod, oracle, pkg, x, y, yaw, r):
    """Random poses (on and off the grid), ranges with NaN / <=0.05 / >=3.95 / >4.0 edge values."""
    ref = _need_ref(orc_mod, 400, "0.05")
    p = pkg.make_params(400, 400, 0.05, 20.0)
    r = np.asarray(r, np.float32)
    ref.reset(0, 0)
    g0 = (np.arange(160000) % 161 - 80).astype(np.int8).reshape(400, 400)   # non-trivial start: exercises both clamps
    np.ctypeslib.as_array(ref.L.ref_grid(), shape=(400, 400))[:] = g0
    ref.frame(x, y, yaw, r)
    want = ref.grid()
    got = g0.copy()
    oracle.L.orc_frame(__import__("ctypes").byref(p), got.ctypes.data, np.float32(x), np.float32(y), np.float32(yaw), r.ctypes.data)
    assert np.array_equal(got, want), first_diff(got, want)


@settings(max_examples=500, deadline=None)
@given(x=st.floats(width=32, allow_nan=True, allow_infinity=True), y=finite)
def test_world_to_grid_matches_reference(orc_mod, oracle, pkg, x, y):
    """A3 incl. NaN/Inf/huge x: lrintf's 'integer indefinite' and the int truncation."""
    ref = _need_ref(orc_mod, 400, "0.05")
    ref.reset(0, 0)
    p = pkg.make_params(400, 400, 0.05, 20.0)
    a = ref.world_to_grid(x, y)
    b = oracle.world_to_grid(p, x, y)
    assert a[0] == b[0]
    if a[0]:
        assert a == b


def test_half_cell_ties_round_to_even(orc_mod, oracle, pkg):
    """lrintf ties-to-even exactly at half-cell offsets."""
    ref = _need_ref(orc_mod, 400, "0.05")
    ref.reset(0, 0)
    p = pkg.make_params(400, 400, 0.05, 20.0)
    for k in range(-50, 50):
        x = np.float32(0.025) + np.float32(0.05) * np.float32(k)
        assert ref.world_to_grid(x, -x) == oracle.world_to_grid(p, x, -x)


def test_raycast_edge_cases(orc_mod, oracle, pkg):
    """single-cell ray (end rule only), max-range end cell unchanged (LO_FREE_DEC/2 == 0), off-grid end dropped."""
    ref = _need_ref(orc_mod, 400, "0.05")
    p = pkg.make_params(400, 400, 0.05, 20.0)
    cases = [(0, 0, 0.01, 0.01, True), (0, 0, 0.01, 0.01, False), (0, 0, 3.0, 1.0, False), (0, 0, 3.0, 1.0, True),
             (0, 0, 30.0, 0.0, True), (30.0, 0, 0.0, 0.0, True), (9.97, 9.97, 9.0, -9.99, True), (-10.0, -10.0, -9.0, -9.0, True),
             (1, 1, 1, -2.5, True), (1, 1, -2.5, 1, True), (1, 1, 3, 3, True), (1, 1, -1, 3, True)]
    ref.reset(0, 0)
    got = np.zeros((400, 400), np.int8)
    for rep in range(30):     # repeat to drive cells into both clamps
        for (x0, y0, x1, y1, hit) in cases:
            ref.raycast_update(x0, y0, x1, y1, hit)
            oracle.raycast(p, got, x0, y0, x1, y1, hit)
    want = ref.grid()
    assert np.array_equal(got, want), first_diff(got, want)
    assert want.max() == 80 and want.min() < -25
