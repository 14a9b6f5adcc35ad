"""oracle/orc.py -- TEST INFRASTRUCTURE: ctypes bindings of the CPU checkers.

* ``Oracle``     liborc.so, the plain-C restatement (oracle/uqs_oracle.c), run-time geometry.
* ``Reference``  oracle/_ref/libref_<W>x<H>_<res>.so, the reference's OWN mapping code extracted and
                 compiled by oracle/build_ref.sh (one library per geometry; macros are compile-time).

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this module.
The product (libuqs_mapping.so and its Python binding) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
ORC_LIB = os.path.join(HERE, "liborc.so")


def build(force: bool = False):
    """Compile liborc.so and (when /root/reference exists) the oracle/_ref libraries."""
    if force or not os.path.exists(ORC_LIB) or os.path.getmtime(ORC_LIB) < os.path.getmtime(os.path.join(HERE, "uqs_oracle.c")):
        subprocess.check_call(["make", "-C", HERE, "liborc.so"], stdout=subprocess.DEVNULL)
    subprocess.check_call(["bash", os.path.join(HERE, "build_ref.sh")], stdout=subprocess.DEVNULL)
    # the real reference translation unit against the product library (needs the library: skipped until it is built)
    if os.path.exists(os.path.join(os.path.dirname(HERE), "micro-quad-slam_b200", "libuqs_mapping.so")):
        subprocess.check_call(["bash", os.path.join(HERE, "link_reference.sh")], stdout=subprocess.DEVNULL)


def _vp(a: Optional[np.ndarray]):
    return None if a is None else C.c_void_p(a.ctypes.data)


class Oracle:
    def __init__(self):
        if not os.path.exists(ORC_LIB):
            build()
        L = self.L = C.CDLL(ORC_LIB)
        L.orc_replay.restype = C.c_long
        L.orc_replay.argtypes = [C.c_void_p, C.c_void_p, C.c_long] + [C.c_void_p] * 4
        L.orc_frame.restype = C.c_long
        L.orc_frame.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_void_p]
        L.orc_raycast.restype = C.c_long
        L.orc_raycast.argtypes = [C.c_void_p, C.c_void_p] + [C.c_float] * 4 + [C.c_int]
        L.orc_world_to_grid.argtypes = [C.c_void_p, C.c_float, C.c_float, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.orc_beam_cells.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_beam_cells.restype = None
        L.orc_pose_integrate.argtypes = [C.c_long] + [C.c_void_p] * 8
        L.orc_pose_integrate.restype = None
        L.orc_pose_integrate_f64.argtypes = [C.c_long] + [C.c_void_p] * 8
        L.orc_pose_integrate_f64.restype = None
        L.orc_sincosf_sweep.restype = C.c_long
        L.orc_sincosf_sweep.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.POINTER(C.c_uint32)]
        L.orc_libm_sincosf.argtypes = [C.c_long, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_libm_sincosf.restype = None
        L.orc_bresenham_closed_form_check.restype = C.c_long
        L.orc_magic_division_check.restype = C.c_long
        L.orc_clamp_monoid_check.restype = C.c_long
        L.orc_clamp_monoid_check.argtypes = [C.c_long, C.c_uint64]
        L.orc_beams_from_scan.argtypes = [C.c_void_p, C.c_float, C.c_void_p, C.c_void_p]
        L.orc_beams_from_scan.restype = None
        L.orc_replay_recentering.argtypes = [C.c_void_p, C.c_void_p, C.c_long] + [C.c_void_p] * 4 + [C.POINTER(C.c_long)]
        L.orc_frontier_score_dir.argtypes = [C.c_void_p, C.c_void_p] + [C.c_float] * 4
        L.orc_slice_map_check.restype = C.c_long
        L.orc_slice_map_check.argtypes = [C.c_long, C.c_uint64]
        L.orc_grid_hash64.restype = C.c_uint64
        L.orc_grid_hash64.argtypes = [C.c_void_p, C.c_size_t]
        L.orc_fnv1a32.restype = C.c_uint32
        L.orc_fnv1a32.argtypes = [C.c_void_p, C.c_size_t]

    # --- mapping ---------------------------------------------------------------------------
    def replay(self, p, x, y, yaw_deg, ranges, grid: Optional[np.ndarray] = None):
        """One flight: returns (grid [H,W] int8, ray-cell updates U).  ``grid`` given = accumulate into it."""
        x, y, yaw_deg = (np.ascontiguousarray(a, np.float32).ravel() for a in (x, y, yaw_deg))
        n = x.size
        ranges = np.ascontiguousarray(ranges, np.float32).reshape(n, 32)
        if grid is None:
            grid = np.zeros((p.H, p.W), np.int8)
        U = self.L.orc_replay(C.byref(p), _vp(grid), n, _vp(x), _vp(y), _vp(yaw_deg), _vp(ranges))
        return grid, int(U)

    def replay_flights(self, p, x, y, yaw_deg, ranges):
        F = x.shape[0]
        grids = np.zeros((F, p.H, p.W), np.int8)
        U = 0
        for f in range(F):
            _, u = self.replay(p, x[f], y[f], yaw_deg[f], ranges[f], grids[f])
            U += u
        return grids, U

    def raycast(self, p, grid, x0, y0, x1, y1, hit) -> int:
        return int(self.L.orc_raycast(C.byref(p), _vp(grid), x0, y0, x1, y1, int(bool(hit))))

    def world_to_grid(self, p, x, y):
        gx, gy = C.c_int(-1), C.c_int(-1)
        ok = self.L.orc_world_to_grid(C.byref(p), np.float32(x), np.float32(y), C.byref(gx), C.byref(gy))
        return bool(ok), gx.value, gy.value

    def beam_cells(self, p, x, y, yaw_deg, ranges):
        x, y, yaw_deg = (np.ascontiguousarray(a, np.float32).ravel() for a in (x, y, yaw_deg))
        n = x.size
        ranges = np.ascontiguousarray(ranges, np.float32).reshape(n, 32)
        cells = np.empty((n, 32, 2), np.int32)
        origin = np.empty((n, 2), np.int32)
        for i in range(n):
            self.L.orc_beam_cells(C.byref(p), x[i], y[i], yaw_deg[i], _vp(ranges[i]), _vp(cells[i]), _vp(origin[i]))
        return cells, origin

    # --- next rows N1-N3 ---------------------------------------------------------------------
    def beams_from_scans(self, raw, max_range=4.0):
        raw = np.ascontiguousarray(raw, np.uint8).reshape(-1, 512)
        n = raw.shape[0]
        beams, dmin = np.empty((n, 32), np.float32), np.empty((n, 4), np.float32)
        for i in range(n):
            self.L.orc_beams_from_scan(_vp(raw[i]), max_range, _vp(beams[i]), _vp(dmin[i]))
        return beams, dmin

    def replay_recentering(self, p, x, y, yaw_deg, ranges):
        """log_tick() with recentering: returns (grid, origin (x,y), n_events, updates); p is not modified."""
        q = type(p)()
        C.memmove(C.byref(q), C.byref(p), C.sizeof(q))
        x, y, yaw_deg = (np.ascontiguousarray(a, np.float32).ravel() for a in (x, y, yaw_deg))
        ranges = np.ascontiguousarray(ranges, np.float32).reshape(x.size, 32)
        grid = np.zeros((p.H, p.W), np.int8)
        U = C.c_long(0)
        n_ev = self.L.orc_replay_recentering(C.byref(q), _vp(grid), x.size, _vp(x), _vp(y), _vp(yaw_deg), _vp(ranges), C.byref(U))
        return grid, (float(q.origin_x), float(q.origin_y)), int(n_ev), int(U.value)

    def frontier_score(self, p, grid, x, y, yaw_deg, offset_deg) -> int:
        g = np.ascontiguousarray(grid, np.int8)
        return int(self.L.orc_frontier_score_dir(C.byref(p), _vp(g), x, y, yaw_deg, offset_deg))

    # --- P0 (builder-defined) --------------------------------------------------------------
    def pose_integrate(self, t_ms, rx, ry, h, yaw_deg, q):
        t_ms = np.ascontiguousarray(t_ms, np.uint32)
        flat = t_ms.ndim == 1
        if flat:
            t_ms = t_ms[None]
        F, N = t_ms.shape
        rx, ry, h, yaw_deg = (np.ascontiguousarray(a, np.float32).reshape(F, N) for a in (rx, ry, h, yaw_deg))
        q = np.ascontiguousarray(q, np.uint8).reshape(F, N)
        xo, yo = np.empty((F, N), np.float32), np.empty((F, N), np.float32)
        for f in range(F):
            self.L.orc_pose_integrate(N, _vp(t_ms[f]), _vp(rx[f]), _vp(ry[f]), _vp(h[f]), _vp(yaw_deg[f]), _vp(q[f]),
                                      _vp(xo[f]), _vp(yo[f]))
        return (xo[0], yo[0]) if flat else (xo, yo)

    def pose_integrate_f64(self, t_ms, rx, ry, h, yaw_deg, q):
        """The binary32 increments of the P0 spec accumulated in binary64 (one log): yardstick for the scan variant."""
        t_ms = np.ascontiguousarray(t_ms, np.uint32).ravel()
        N = t_ms.size
        rx, ry, h, yaw_deg = (np.ascontiguousarray(a, np.float32).ravel() for a in (rx, ry, h, yaw_deg))
        q = np.ascontiguousarray(q, np.uint8).ravel()
        xo, yo = np.empty(N, np.float64), np.empty(N, np.float64)
        self.L.orc_pose_integrate_f64(N, _vp(t_ms), _vp(rx), _vp(ry), _vp(h), _vp(yaw_deg), _vp(q), _vp(xo), _vp(yo))
        return xo, yo

    # --- arithmetic KATs --------------------------------------------------------------------
    def libm_sincosf(self, a):
        a = np.ascontiguousarray(a, np.float32).ravel()
        s, c = np.empty_like(a), np.empty_like(a)
        self.L.orc_libm_sincosf(a.size, _vp(a), _vp(s), _vp(c))
        return s, c

    def sincosf_sweep(self, lo_bits, hi_bits, stride=1, threads=1):
        first = C.c_uint32(0)
        bad = self.L.orc_sincosf_sweep(lo_bits, hi_bits, stride, threads, C.byref(first))
        return int(bad), first.value

    def grid_hash64(self, a: np.ndarray) -> int:
        """sum over cells of splitmix64(i << 8 | byte): the digest the product computes on the device."""
        a = np.ascontiguousarray(a)
        return int(self.L.orc_grid_hash64(_vp(a), a.nbytes))

    def fnv1a32(self, a: np.ndarray) -> int:
        a = np.ascontiguousarray(a)
        return int(self.L.orc_fnv1a32(_vp(a), a.nbytes))


def ref_lib_path(W: int, H: int, res: str) -> str:
    return os.path.join(REF_DIR, f"libref_{W}x{H}_{res}.so")


class Reference:
    """The reference's own mapping code for one compile-time geometry."""

    def __init__(self, W: int, H: int, res: str):
        path = ref_lib_path(W, H, res)
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} not built (oracle/build_ref.sh needs /root/reference)")
        L = self.L = C.CDLL(path)
        f = C.c_float
        L.ref_reset.argtypes = [f, f]
        L.ref_grid.restype = C.POINTER(C.c_int8)
        L.ref_map_res.restype = f
        L.ref_map_size.restype = f
        L.ref_origin_x.restype = f
        L.ref_origin_y.restype = f
        L.ref_frame.argtypes = [f, f, f, C.c_void_p, C.c_int]
        L.ref_replay.argtypes = [C.c_long] + [C.c_void_p] * 4 + [C.c_int]
        L.ref_world_to_grid.argtypes = [f, f, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.ref_raycast_update.argtypes = [f, f, f, f, C.c_int]
        L.ref_recenter_shift.argtypes = [C.c_int, C.c_int]
        L.ref_recentre_if_needed.argtypes = [f, f]
        L.ref_frontier_score_dir.argtypes = [f, f, f, f]
        L.ref_beams_from_frame.argtypes = [C.c_void_p, C.c_void_p]
        self.W, self.H = L.ref_map_w(), L.ref_map_h()
        assert (self.W, self.H) == (W, H)
        self.res = float(L.ref_map_res())

    def reset(self, ox=0.0, oy=0.0):
        self.L.ref_reset(ox, oy)

    def grid(self) -> np.ndarray:
        return np.ctypeslib.as_array(self.L.ref_grid(), shape=(self.H, self.W)).copy()

    def frame(self, x, y, yaw_deg, beams32, allow_recenter=False):
        b = np.ascontiguousarray(beams32, np.float32).ravel()
        self.L.ref_frame(np.float32(x), np.float32(y), np.float32(yaw_deg), _vp(b), int(allow_recenter))

    def replay(self, x, y, yaw_deg, ranges, ox=0.0, oy=0.0, reset=True, allow_recenter=False):
        x, y, yaw_deg = (np.ascontiguousarray(a, np.float32).ravel() for a in (x, y, yaw_deg))
        ranges = np.ascontiguousarray(ranges, np.float32).reshape(x.size, 32)
        if reset:
            self.reset(ox, oy)
        self.L.ref_replay(x.size, _vp(x), _vp(y), _vp(yaw_deg), _vp(ranges), int(allow_recenter))
        return self.grid()

    def world_to_grid(self, x, y):
        gx, gy = C.c_int(-1), C.c_int(-1)
        ok = self.L.ref_world_to_grid(np.float32(x), np.float32(y), C.byref(gx), C.byref(gy))
        return bool(ok), gx.value, gy.value

    def raycast_update(self, x0, y0, x1, y1, hit):
        self.L.ref_raycast_update(np.float32(x0), np.float32(y0), np.float32(x1), np.float32(y1), int(bool(hit)))

    def recentered(self) -> bool:
        return bool(self.L.ref_recentered())

    def origin(self):
        return float(self.L.ref_origin_x()), float(self.L.ref_origin_y())

    def frontier_score(self, x, y, yaw_deg, offset_deg) -> int:
        return int(self.L.ref_frontier_score_dir(np.float32(x), np.float32(y), np.float32(yaw_deg), np.float32(offset_deg)))

    def set_grid(self, grid):
        np.ctypeslib.as_array(self.L.ref_grid(), shape=(self.H, self.W))[:] = grid

    def beams_from_frame(self, raw512) -> np.ndarray:
        """compute_beams_and_minima on a 518-byte wire frame built around the 512 raw bytes."""
        frame = np.zeros(518, np.uint8)
        frame[0] = 0xA5
        frame[5:517] = np.ascontiguousarray(raw512, np.uint8).ravel()
        out = np.empty(32, np.float32)
        self.L.ref_beams_from_frame(_vp(frame), _vp(out))
        return out
