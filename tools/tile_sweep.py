"""Sub-tile shape sweep (value mode, no slices) for a single-flight big-grid config."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
m = importlib.import_module("micro-quad-slam_b200"); syn = importlib.import_module("micro-quad-slam_b200.synth")
m.init(0); dev = torch.device("cuda:0"); m.set_stream(torch.cuda.current_stream().cuda_stream)
name = sys.argv[1]
shapes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[2:]]
w = syn.CONFIGS[name]; d = syn.generate(w); p = w.params()
x, y = syn.frame_poses(d, d["x_true"], d["y_true"])
t = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (x, y, d["frame_yaw_deg"], d["ranges"])]
g = torch.empty((w.n_flights, p.H, p.W), dtype=torch.int8, device=dev)
m.set_engine(1, 0)
for (a, b) in shapes:
    m.set_tuning(a, b, 1)
    try:
        m.replay_dev(p, w.n_flights, w.n_frames, *(v.data_ptr() for v in t), g.data_ptr())
    except Exception as e:
        print(f"{name} tile {a}x{b}: {e}"); continue
    best = 1e9
    for _ in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); m.replay_dev(p, w.n_flights, w.n_frames, *(v.data_ptr() for v in t), g.data_ptr()); e1.record()
        torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    print(f"{name} tile {a}x{b}: {best:.2f} ms", flush=True)
