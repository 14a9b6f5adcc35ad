"""Scratch: sub-tile shape sweep on a C3 slice (device-resident)."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
m = importlib.import_module("micro-quad-slam_b200"); syn = importlib.import_module("micro-quad-slam_b200.synth")
m.init(0); dev = torch.device("cuda:0"); m.set_stream(torch.cuda.current_stream().cuda_stream)
w = syn.scaled(syn.CONFIGS["c3"], n_flights=2048); d = syn.generate(w); p = w.params()
t = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (d["x_true"], d["y_true"], d["frame_yaw_deg"], d["ranges"])]
g = torch.empty((w.n_flights, p.H, p.W), dtype=torch.int8, device=dev)
m.set_engine(1, 0)
for (a, b) in [(80, 80), (100, 100), (100, 80), (80, 100), (134, 80), (80, 134), (68, 68), (100, 68), (134, 100), (200, 50), (50, 200), (400, 20)]:
    m.set_tuning(a, b, 1)
    m.replay_dev(p, w.n_flights, w.n_frames, *(x.data_ptr() for x in t), g.data_ptr())
    m.set_profiling(True); m.profile_collect()
    for _ in range(2):
        m.replay_dev(p, w.n_flights, w.n_frames, *(x.data_ptr() for x in t), g.data_ptr())
    ms, cnt = m.profile_collect(); m.set_profiling(False)
    print(f"tile {a}x{b}: replay {ms[2]/2:.2f} ms setup {ms[1]/2:.2f} ms", flush=True)
