"""The drop-in boundary against the file it cites: the REAL /root/reference/uav_local_nav.c, with its mapping
block (:188-216, :229, :241-385, :108) cut out at build time by oracle/link_reference.sh, compiles against
include/uqs_mapping.h and links against libuqs_mapping.so.  The un-vendored common/mavlink.h (:48) is a
do-nothing stub (oracle/stub); the program is linked, never run."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/uav_local_nav.c"
LINK = os.path.join(ROOT, "oracle", "_ref", "link")

# what the rest of uav_local_nav.c uses from the block that was cut (callers: log_tick :1629-1635, frontier turn
# choice :1718-1720 and :2228-2231, HOVER map init :2187-2194, scan record :1572-1573, beam computation :1353)
EXPECTED = {"map_update_from_beams", "map_recentre_if_needed", "frontier_score_dir", "map_reset", "map_inited",
            "map_origin_x", "map_origin_y", "tof_beams_m", "pending_kf_flags", "uqs_init", "uqs_dropin_configure",
            "uqs_params_default", "uqs_last_error"}


def _syms(args):
    out = subprocess.check_output(["nm"] + args, text=True)
    return {ln.split()[-1] for ln in out.splitlines() if ln.strip()}


@pytest.mark.skipif(not os.path.exists(REF), reason="the reference is not present on this machine (GPU box)")
def test_real_reference_translation_unit_links_against_the_library(pkg):
    r = subprocess.run(["bash", os.path.join(ROOT, "oracle", "link_reference.sh")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    obj, exe = os.path.join(LINK, "uav_local_nav_cut.o"), os.path.join(LINK, "uav_local_nav_b200")
    assert os.path.exists(obj) and os.path.exists(exe)
    assert not os.path.exists(os.path.join(LINK, "uav_local_nav_cut.c")), "edited reference text must not outlive the compile"
    undefined = _syms(["-u", obj])
    exported = _syms(["-D", "--defined-only", os.path.join(ROOT, "micro-quad-slam_b200", "libuqs_mapping.so")])
    mapping = {s for s in undefined if s in exported}
    assert EXPECTED <= mapping, f"the cut translation unit no longer references {sorted(EXPECTED - mapping)}"
    # nothing mapping-related is left undefined, and nothing of the removed block is still defined in the object
    leftovers = {"occ_grid", "occ_grid_tmp", "world_to_grid", "raycast_update", "clamp_lo", "map_recenter_shift"} & _syms([obj]) - undefined
    assert not leftovers, f"still defined inside the reference object: {sorted(leftovers)}"
    ldd = subprocess.check_output(["ldd", exe], text=True)
    assert "libuqs_mapping.so" in ldd and "not found" not in ldd
    still_undefined = {ln.split()[-1] for ln in subprocess.check_output(["nm", "-u", exe], text=True).splitlines()
                       if "@" not in ln and ln.split()[-1] in EXPECTED | {"occ_grid", "world_to_grid", "raycast_update"}}
    # (dynamic symbols resolved at load time appear in `nm -u`; what matters is that the library defines every one of them)
    assert still_undefined <= exported
