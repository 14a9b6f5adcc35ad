"""Replay-kernel time alone (profiling spans) on a C3 slice, resident engine; used with UQS_LIBRARY variants."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
m = importlib.import_module("micro-quad-slam_b200"); syn = importlib.import_module("micro-quad-slam_b200.synth")
m.init(0); dev = torch.device("cuda:0"); m.set_stream(torch.cuda.current_stream().cuda_stream)
nf = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
w = syn.scaled(syn.CONFIGS["c3"], n_flights=nf); d = syn.generate(w); p = w.params()
t = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (d["x_true"], d["y_true"], d["frame_yaw_deg"], d["ranges"])]
g = torch.zeros((w.n_flights, p.H, p.W), dtype=torch.int8, device=dev)
for nw in [int(a) for a in sys.argv[2:]] or [4]:
    m.set_engine(2, nw)
    m.replay_dev(p, w.n_flights, w.n_frames, *(a.data_ptr() for a in t), g.data_ptr())
    m.set_profiling(True); m.profile_collect()
    for _ in range(3):
        m.replay_dev(p, w.n_flights, w.n_frames, *(a.data_ptr() for a in t), g.data_ptr())
    ms, cnt = m.profile_collect(); m.set_profiling(False)
    print(f"F={nf} nw={nw}: setup {ms[1]/3:.3f} ms  replay {ms[2]/3:.3f} ms", flush=True)
