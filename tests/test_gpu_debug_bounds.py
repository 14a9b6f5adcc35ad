"""The replay kernels under the bounds-checked build (libuqs_mapping_dbg.so, -DUQS_DEBUG_BOUNDS): every shared-memory
access is checked against the region it is meant for (a flight's resident box, a warp's collision table, the decode
ring, a warp's sub-tile, its candidate queue) and traps otherwise.  compute-sanitizer is closed on the pool, and a
stray write into a neighbouring region need not show in the grids.  Runs the randomized stress, the ragged /
degenerate logs, the short-range (collision-heavy) and the time-slice cases in a subprocess bound to that library."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DBG = os.path.join(ROOT, "micro-quad-slam_b200", "libuqs_mapping_dbg.so")


def test_stress_and_ragged_logs_under_the_bounds_checked_build(pkg):
    assert os.path.exists(DBG), "libuqs_mapping_dbg.so missing: __graft_entry__.build() builds it (make debug)"
    sel = ("random_geometry or ragged or short_ranges or saturation_hazards or time_slices or subtile_size or "
           "c3_drift or c2_long or c4_multizone or millimetre or every_yaw")
    env = dict(os.environ, UQS_LIBRARY=DBG)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-q", "-x", "-m", "gpu",
                        "-k", sel, "-p", "no:cacheprovider"], capture_output=True, text=True, env=env, timeout=1500, cwd=ROOT)
    tail = r.stdout[-3000:] + r.stderr[-2000:]
    assert r.returncode == 0 and "UQS_DEBUG_BOUNDS" not in r.stdout + r.stderr, tail
    assert " passed" in r.stdout and "failed" not in r.stdout, tail


def test_the_bounds_check_has_teeth(pkg):
    """The debug library really contains the trap path (SASS), the release library does not."""
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    n_dbg = subprocess.check_output([cuobjdump, "-sass", DBG], text=True).count("BPT.TRAP")
    n_rel = subprocess.check_output([cuobjdump, "-sass", os.path.join(ROOT, "micro-quad-slam_b200", "libuqs_mapping.so")], text=True).count("BPT.TRAP")
    assert n_dbg > 0 and n_rel == 0
