"""What one collision-checked step per frame costs the resident engine: config 3 with K0 biased by +0..+3 steps
(uqs_set_k0_bias; identical bytes).  tools/k0_stats.py: K0 is 7.2 on average while beams really share cells up to step
4.9, so an exact bound would remove ~1.3 steps per frame."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
m = importlib.import_module("micro-quad-slam_b200"); syn = importlib.import_module("micro-quad-slam_b200.synth")
m.init(0); dev = torch.device("cuda:0"); m.set_stream(torch.cuda.current_stream().cuda_stream)
w = syn.CONFIGS["c3"]; d = syn.generate(w); p = w.params()
t = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (d["x_true"], d["y_true"], d["frame_yaw_deg"], d["ranges"])]
g = torch.empty((w.n_flights, p.H, p.W), dtype=torch.int8, device=dev)
ref = None
for bias in (0, 1, 2, 3, 0):
    m.set_k0_bias(bias)
    m.replay_dev(p, w.n_flights, w.n_frames, *(a.data_ptr() for a in t), g.data_ptr())
    m.set_profiling(True); m.profile_collect()
    for _ in range(3):
        m.replay_dev(p, w.n_flights, w.n_frames, *(a.data_ptr() for a in t), g.data_ptr())
    ms, cnt = m.profile_collect(); m.set_profiling(False)
    h = m.grid_hashes_dev(g.data_ptr(), w.n_flights, p.W * p.H).sum(dtype=np.uint64)
    ref = h if ref is None else ref
    print(f"K0 bias +{bias}: ray set-up {ms[1]/3:.2f} ms  replay {ms[2]/3:.2f} ms  {'same bytes' if h == ref else 'DIFFERENT BYTES'}", flush=True)
m.set_k0_bias(0)
