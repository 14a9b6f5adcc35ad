"""Seeded synthetic flight logs (binding of libuqs_synth.so) and the BASELINE.json configurations.

The reference has no simulator or recorded log; SURVEY.md section 8(d) defines the synthetic
inputs.  ``generate`` fills the SoA arrays the replay consumes; ``CONFIGS`` holds the five
BASELINE.json configurations at full size, ``scaled`` shrinks them for tests.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, replace
from typing import Dict, Optional

import numpy as np

from . import SYNTH_LIB_PATH, Params, make_params


class SynthCfg(C.Structure):
    _fields_ = [
        ("config_id", C.c_int32), ("n_samples", C.c_int32), ("frames_per_sample", C.c_int32),
        ("traj_kind", C.c_int32), ("rate_hz", C.c_float),
        ("room_w", C.c_float), ("room_h", C.c_float), ("room_x0", C.c_float), ("room_y0", C.c_float),
        ("traj_ax", C.c_float), ("traj_ay", C.c_float), ("traj_period_s", C.c_float),
        ("speed_mps", C.c_float), ("line_spacing_m", C.c_float), ("yaw_rate_dps", C.c_float),
        ("max_range_m", C.c_float), ("sigma_r", C.c_float), ("sigma_f", C.c_float), ("sigma_b", C.c_float),
        ("p_dropout", C.c_float), ("p_lowq", C.c_float), ("h_m", C.c_float), ("shared_truth", C.c_int32),
        ("range_mm", C.c_int32),
    ]


@dataclass(frozen=True)
class Workload:
    """One BASELINE.json configuration: geometry + log shape + sensor model."""
    name: str
    config_id: int
    n_flights: int
    n_samples: int
    W: int
    res: str                 # decimal string: the literal the reference would be compiled with
    size_m: float
    rate_hz: float
    frames_per_sample: int = 1
    traj_kind: int = 0
    room: tuple = (12.0, 9.0, -6.0, -4.5)
    traj: tuple = (4.6, 3.2, 47.0)       # ax, ay, period
    speed_mps: float = 1.0
    line_spacing_m: float = 1.0
    sigma_r: float = 0.01
    sigma_f: float = 0.02
    sigma_b: float = 0.005
    shared_truth: int = 0
    range_mm: int = 0        # 1: ranges on the sensor's 1 mm lattice, (float)mm * 0.001f (uav_local_nav.c:1328)

    @property
    def n_frames(self) -> int:
        return self.n_samples * self.frames_per_sample

    def params(self) -> Params:
        return make_params(self.W, self.W, float(self.res), self.size_m)

    def cfg(self) -> SynthCfg:
        c = SynthCfg()
        c.config_id, c.n_samples, c.frames_per_sample = self.config_id, self.n_samples, self.frames_per_sample
        c.traj_kind, c.rate_hz = self.traj_kind, self.rate_hz
        c.room_w, c.room_h, c.room_x0, c.room_y0 = self.room
        c.traj_ax, c.traj_ay, c.traj_period_s = self.traj
        c.speed_mps, c.line_spacing_m, c.yaw_rate_dps = self.speed_mps, self.line_spacing_m, 20.0
        c.max_range_m, c.sigma_r, c.sigma_f, c.sigma_b = 4.0, self.sigma_r, self.sigma_f, self.sigma_b
        c.p_dropout, c.p_lowq, c.h_m, c.shared_truth = 0.02, 0.01, 0.5, self.shared_truth
        c.range_mm = self.range_mm
        return c


C5_RES = ["0.02", "0.025", "0.03", "0.035", "0.04", "0.045", "0.05", "0.055", "0.06", "0.065", "0.07", "0.075",
          "0.08", "0.085", "0.09", "0.10"]
C5_SIGMA_R = [round(i * 0.10 / 15, 6) for i in range(16)]


def c5_width(res: str) -> int:
    w = round(20 / float(res))
    return w + (w % 2)


CONFIGS: Dict[str, Workload] = {
    # single 60 s flight, 50 Hz flow+ToF, 400x400 @ 5 cm (the reference's CPU-runnable case)
    "c1": Workload("c1", 1, 1, 3000, 400, "0.05", 20.0, 50.0),
    # 1 h at 100 Hz into 2000x2000 @ 1 cm
    "c2": Workload("c2", 2, 1, 360000, 2000, "0.01", 20.0, 100.0, traj=(4.6, 3.2, 61.0)),
    # 4096-flight optical-flow drift ensemble, one 400x400 grid each
    "c3": Workload("c3", 3, 4096, 3000, 400, "0.05", 20.0, 50.0, shared_truth=1),
    # 64-beam building sweep into 16384^2 @ 1 cm (two 32-beam reference frames per sample)
    "c4": Workload("c4", 4, 1, 1048576, 16384, "0.01", 163.84, 100.0, frames_per_sample=2, traj_kind=1,
                   room=(10.0, 10.0, -5.0, -5.0), traj=(48.0, 48.0, 0.0), sigma_b=0.0002),
}


def c5_workload(i_res: int, i_sigma: int, n_flights: int = 64, n_samples: int = 3000) -> Workload:
    """Config 5: resolution x range-noise sweep; (i_res, i_sigma) in [0,16)^2, 64 flights each."""
    res = C5_RES[i_res]
    return Workload(f"c5_r{i_res}_s{i_sigma}", 5 * 256 + i_res * 16 + i_sigma, n_flights, n_samples, c5_width(res),
                    res, 20.0, 50.0, sigma_r=C5_SIGMA_R[i_sigma])


def scaled(w: Workload, n_flights: Optional[int] = None, n_samples: Optional[int] = None) -> Workload:
    return replace(w, n_flights=n_flights or w.n_flights, n_samples=n_samples or w.n_samples)


_lib = None


def _synth_lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SYNTH_LIB_PATH):
            raise RuntimeError(f"{SYNTH_LIB_PATH} missing: run __graft_entry__.build()")
        _lib = C.CDLL(SYNTH_LIB_PATH)
        _lib.uqs_synth_generate.argtypes = [C.POINTER(SynthCfg), C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 9
    return _lib


def generate(w: Workload, flight_id0: int = 0, n_flights: Optional[int] = None, n_threads: Optional[int] = None,
             out: Optional[dict] = None) -> dict:
    """Logs for flights [flight_id0, flight_id0+n_flights) of workload ``w`` as numpy SoA arrays.

    Keys: t_ms u32, of_rate_x/of_rate_y/h_m/yaw_deg f32, of_q u8, x_true/y_true f32 (all [F, n_samples]);
    ranges f32 [F, n_frames, 32]; frame_yaw_deg f32 [F, n_frames] (sample yaw, +45 deg on the second frame
    of a 64-beam sample); frame_sample i32 [n_frames] (sample index of each frame).
    """
    F = n_flights if n_flights is not None else w.n_flights
    N, fps = w.n_samples, w.frames_per_sample
    n_threads = n_threads or min(os.cpu_count() or 1, 64)
    d = out if out is not None else {}

    def buf(name, shape, dt):
        a = d.get(name)
        if a is None or a.shape != shape or a.dtype != dt:
            a = np.empty(shape, dt)
            d[name] = a
        return a

    t = buf("t_ms", (F, N), np.uint32)
    rx, ry = buf("of_rate_x", (F, N), np.float32), buf("of_rate_y", (F, N), np.float32)
    h, yaw = buf("h_m", (F, N), np.float32), buf("yaw_deg", (F, N), np.float32)
    q = buf("of_q", (F, N), np.uint8)
    ranges = buf("ranges", (F, N * fps, 32), np.float32)
    xt, yt = buf("x_true", (F, N), np.float32), buf("y_true", (F, N), np.float32)
    cfg = w.cfg()
    rc = _synth_lib().uqs_synth_generate(C.byref(cfg), flight_id0, F, n_threads, *(C.c_void_p(a.ctypes.data) for a in
                                         (t, rx, ry, h, yaw, q, ranges, xt, yt)))
    if rc != 0:
        raise RuntimeError(f"uqs_synth_generate failed with {rc}")
    d["frame_sample"] = np.repeat(np.arange(N, dtype=np.int32), fps)
    if fps == 1:
        d["frame_yaw_deg"] = yaw
    else:
        fy = np.repeat(yaw, fps, axis=1)
        fy[:, 1::2] = fy[:, 1::2] + np.float32(45.0)      # binary32 add, as a caller of the reference would do
        d["frame_yaw_deg"] = fy
    return d


def on_mm_lattice(w: Workload) -> Workload:
    """The same workload with its ranges on the ToF sensor's millimetre lattice (what a real scan log holds)."""
    return replace(w, range_mm=1)


def ranges_to_mm(ranges: np.ndarray, out: Optional[np.ndarray] = None) -> np.ndarray:
    """u16 millimetre form of ranges generated with ``range_mm=1``: NaN (no return) -> 0xFFFF.  Lossless:
    ``mm_to_ranges(ranges_to_mm(r))`` reproduces r bit for bit (checked by the caller or the tests)."""
    r = np.asarray(ranges, np.float32)
    mm = out if out is not None else np.empty(r.shape, np.uint16)
    blk = 1 << 24
    rf, mf = r.reshape(-1), mm.reshape(-1)
    for a in range(0, rf.size, blk):
        v = rf[a:a + blk]
        nan = np.isnan(v)
        q = np.rint(np.where(nan, np.float32(0), v) * np.float32(1000.0))
        q[nan] = 65535
        mf[a:a + blk] = q.astype(np.uint16)
    return mm


def mm_to_ranges(mm: np.ndarray) -> np.ndarray:
    """(float)mm * 0.001f in binary32, 0xFFFF -> NaN: the conversion the device applies (uav_local_nav.c:1327-1328)."""
    r = mm.astype(np.float32) * np.float32(0.001)
    r[mm == 0xFFFF] = np.nan
    return r


def frame_poses(d: dict, x: np.ndarray, y: np.ndarray):
    """Expand per-sample poses to per-frame poses (64-beam samples replay two frames at one pose)."""
    idx = d["frame_sample"]
    if idx.size == x.shape[-1]:
        return x, y
    return np.ascontiguousarray(x[..., idx]), np.ascontiguousarray(y[..., idx])
