"""How loose is the collision bound K0?  Brute force on C3 frames: last step at which two beams really share a cell."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
m = importlib.import_module("micro-quad-slam_b200"); syn = importlib.import_module("micro-quad-slam_b200.synth")
m.init(0)
w = syn.scaled(syn.CONFIGS["c3"], n_flights=1, n_samples=600); d = syn.generate(w); p = w.params()
x, y = d["x_true"][0], d["y_true"][0]
cells, origin = m.beam_cells(p, x, y, d["frame_yaw_deg"][0], d["ranges"][0])
k0, srt = m.frame_bounds(p, x, y, d["frame_yaw_deg"][0], d["ranges"][0])
def walk(dx, dy):
    x = y = 0; sx = 1 if dx > 0 else -1; sy = 1 if dy > 0 else -1; ax, ay = abs(dx), -abs(dy); err = ax + ay; out = [(0, 0)]
    while (x, y) != (dx, dy):
        e2 = 2 * err
        if e2 >= ay: err += ay; x += sx
        if e2 <= ax: err += ax; y += sy
        out.append((x, y))
    return out
ls = []; steps_with_collision = []
for f in range(0, 600, 3):
    beams = [(b, int(cells[f, b, 0] - origin[f, 0]), int(cells[f, b, 1] - origin[f, 1])) for b in range(32) if cells[f, b, 0] >= 0]
    wk = {b: walk(dx, dy) for b, dx, dy in beams}
    last = -1; coll = set()
    for i, (bi, _, _) in enumerate(beams):
        for bj, _, _ in beams[i + 1:]:
            for k in range(min(len(wk[bi]), len(wk[bj]), 12)):
                if wk[bi][k] == wk[bj][k]: last = max(last, k); coll.add(k)
    ls.append(last); steps_with_collision.append(len(coll))
ls = np.array(ls); kk = k0[0:600:3]
print("K0 mean", kk.mean(), "last shared step mean", ls.mean(), "slack (K0-1-last) mean", (kk - 1 - ls).mean(), "hist last:", np.bincount(ls + 1), "steps with any collision mean", np.mean(steps_with_collision))
