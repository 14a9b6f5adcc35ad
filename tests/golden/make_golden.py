"""Generate tests/golden/*.npz from the reference's OWN mapping code (oracle/_ref, built by oracle/build_ref.sh from
/root/reference/uav_local_nav.c).  Run in the build container, where the reference exists:

    python tests/golden/make_golden.py

Each fixture holds seeded inputs (poses, yaw, ranges) and the reference's resulting grid (as its non-zero cells), so
that the oracle restatement and the device path can be checked against the reference on machines that have neither
/root/reference nor the oracle/_ref libraries.
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import orc  # noqa: E402

synth = importlib.import_module("micro-quad-slam_b200.synth")
HERE = os.path.dirname(os.path.abspath(__file__))


def save(name, W, res, size, x, y, yaw, ranges, origin=(0.0, 0.0), recenter=False):
    ref = orc.Reference(W, W, res)
    grid = ref.replay(x, y, yaw, ranges, ox=origin[0], oy=origin[1], allow_recenter=recenter)
    nz = np.flatnonzero(grid)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), W=W, res=res, size=size, origin=np.asarray(origin, np.float32),
                        x=x.astype(np.float32), y=y.astype(np.float32), yaw=yaw.astype(np.float32),
                        ranges=ranges.astype(np.float32), nz_index=nz.astype(np.int32), nz_value=grid.ravel()[nz],
                        final_origin=np.asarray(ref.origin(), np.float32), recenter=recenter,
                        fnv=np.uint32(orc.Oracle().fnv1a32(grid)))
    print(name, "cells", nz.size, "fnv %08x" % orc.Oracle().fnv1a32(grid))


def main():
    orc.build()
    # 1. the SURVEY Appendix-D probe at the reference's native geometry
    f = np.float32
    b = np.zeros((4, 8), f)
    for d in range(4):
        for c in range(8):
            b[d, c] = f(1.0) + f(0.1) * f(c) + f(0.5) * f(d)
    b[2, 3] = np.nan; b[1, 1] = 4.5; b[0, 0] = 3.97
    k = np.arange(100, dtype=f)
    save("kat_appendix_d_500_0.10", 500, "0.10", 50.0, f(0.01) * k, f(-0.02) * k, f(3.0) * k, np.broadcast_to(b.reshape(1, 32), (100, 32)).copy())
    # 2. a 12 s slice of the config-1 flight (400x400 @ 0.05)
    w = synth.scaled(synth.CONFIGS["c1"], n_samples=600)
    d = synth.generate(w, n_threads=1)
    save("c1_slice_400_0.05", 400, "0.05", 20.0, d["x_true"][0], d["y_true"][0], d["yaw_deg"][0], d["ranges"][0])
    # 3. a fine-grid slice (2000x2000 @ 0.01), moved origin
    w = synth.scaled(synth.CONFIGS["c2"], n_samples=40)
    d = synth.generate(w, n_threads=1)
    save("c2_slice_2000_0.01", 2000, "0.01", 20.0, d["x_true"][0], d["y_true"][0], d["yaw_deg"][0], d["ranges"][0], origin=(0.37, -0.21))
    # 4. a log that leaves the map (recentering, N2)
    w = synth.Workload("wander", 99, 1, 900, 400, "0.05", 20.0, 50.0)
    d = synth.generate(w, n_threads=1)
    t = np.arange(900, dtype=f) / f(50.0)
    save("recenter_400_0.05", 400, "0.05", 20.0, d["x_true"][0] + f(1.4) * t, d["y_true"][0] - f(0.9) * t, d["yaw_deg"][0], d["ranges"][0], recenter=True)


if __name__ == "__main__":
    main()
