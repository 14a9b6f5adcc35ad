"""pytest configuration: the `gpu` marker, shared fixtures, in-tree builds of the checkers."""
import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: exhaustive sweep, a minute or more")


@pytest.fixture(scope="session")
def orc_mod():
    from oracle import orc
    orc.build()
    return orc


@pytest.fixture(scope="session")
def oracle(orc_mod):
    return orc_mod.Oracle()


@pytest.fixture(scope="session")
def pkg():
    import __graft_entry__ as ge
    ge.build_host_only()
    return importlib.import_module("micro-quad-slam_b200")


@pytest.fixture(scope="session")
def synth(pkg):
    return importlib.import_module("micro-quad-slam_b200.synth")


@pytest.fixture(scope="session")
def gpu(pkg):
    """Initialised library on cuda:0; fails (not skips) if the CUDA library cannot run."""
    pkg.init(0)
    return pkg


def have_ref(orc_mod, W, H, res):
    return os.path.exists(orc_mod.ref_lib_path(W, H, res))


def first_diff(a: np.ndarray, b: np.ndarray) -> str:
    d = np.argwhere(a != b)
    if d.size == 0:
        return "identical"
    i = tuple(d[0])
    return f"{d.shape[0]} cells differ; first at {i}: got {a[i]} want {b[i]}"
