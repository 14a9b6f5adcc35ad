"""Golden vectors generated from the reference's own code (tests/golden/make_golden.py): the oracle restatement (CPU)
and the device path (GPU) must reproduce them without /root/reference or oracle/_ref being present."""
import glob
import os

import numpy as np
import pytest

from conftest import first_diff

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURES = sorted(glob.glob(os.path.join(HERE, "golden", "*.npz")))


def load(path, pkg):
    z = np.load(path)
    W = int(z["W"])
    p = pkg.make_params(W, W, float(str(z["res"])), float(z["size"]), origin=tuple(z["origin"]))
    want = np.zeros(W * W, np.int8)
    want[z["nz_index"]] = z["nz_value"]
    return z, p, want.reshape(W, W)


def test_fixtures_exist():
    assert len(FIXTURES) >= 4


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p)[:-4] for p in FIXTURES])
def test_oracle_reproduces_reference_golden(path, pkg, oracle):
    z, p, want = load(path, pkg)
    if bool(z["recenter"]):
        got, origin, n_ev, _ = oracle.replay_recentering(p, z["x"], z["y"], z["yaw"], z["ranges"])
        assert n_ev >= 1 and np.float32(origin[0]) == z["final_origin"][0] and np.float32(origin[1]) == z["final_origin"][1]
    else:
        got, _ = oracle.replay(p, z["x"], z["y"], z["yaw"], z["ranges"])
    assert np.array_equal(got, want), first_diff(got, want)
    assert oracle.fnv1a32(got) == int(z["fnv"])


@pytest.mark.gpu
@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p)[:-4] for p in FIXTURES])
def test_device_reproduces_reference_golden(path, gpu, oracle):
    z, p, want = load(path, gpu)
    if bool(z["recenter"]):
        got, origin, events, _ = gpu.replay_recentering(p, z["x"], z["y"], z["yaw"], z["ranges"])
        assert len(events) >= 1 and np.float32(origin[0]) == z["final_origin"][0]
    else:
        for engine in (1, 2):
            gpu.set_engine(engine, 0)
            try:
                if engine == 2 and p.W > 476:
                    continue
                got = gpu.replay(p, z["x"], z["y"], z["yaw"], z["ranges"])[0][0]
                assert np.array_equal(got, want), (engine, first_diff(got, want))
            finally:
                gpu.set_engine(0, 0)
        got = gpu.replay(p, z["x"], z["y"], z["yaw"], z["ranges"])[0][0]
    assert np.array_equal(got, want), first_diff(got, want)
    assert oracle.fnv1a32(got) == int(z["fnv"])
