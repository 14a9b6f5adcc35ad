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


# ---- background jobs: the reference's own code on the two long single logs (tests/ref_jobs.py) --------------------
_BG = {}


def pytest_collection_finish(session):
    """The whole-log parity tests of configs 2 and 4 need ~15 s / ~80 s of the reference's code on one core each;
    start them as soon as it is known that they will run, so that they finish behind the rest of the session."""
    import subprocess
    import tempfile
    wanted = {"c2": "test_c2_the_whole_one_hour_log", "c4": "test_c4_the_whole_building_sweep"}
    names = [it.name for it in session.items]
    if not any(n in names for n in wanted.values()) or session.config.option.collectonly:
        return
    from oracle import orc
    tmp = tempfile.mkdtemp(prefix="uqs_ref_")
    geom = {"c2": (2000, "0.01"), "c4": (16384, "0.01")}
    for key, test in wanted.items():
        W, res = geom[key]
        if test in names and os.path.exists(orc.ref_lib_path(W, W, res)):
            path = os.path.join(tmp, key + ".json")
            proc = subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "ref_jobs.py"), key, path],
                                    stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            _BG[key] = (proc, path)


def pytest_sessionfinish(session, exitstatus):
    for proc, _ in _BG.values():
        if proc.poll() is None:
            proc.kill()


def background_reference_job(name):
    return _BG.get(name)


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
