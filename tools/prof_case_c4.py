"""Fixed C4 slice for ncu captures of the sub-tile engine on a large grid."""
import importlib, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
m = importlib.import_module("micro-quad-slam_b200"); syn = importlib.import_module("micro-quad-slam_b200.synth")
m.init(0); dev = torch.device("cuda:0"); m.set_stream(torch.cuda.current_stream().cuda_stream)
name = sys.argv[1] if len(sys.argv) > 1 else "c4"
ns = int(sys.argv[2]) if len(sys.argv) > 2 else 300000
w = syn.scaled(syn.CONFIGS[name], n_samples=ns); d = syn.generate(w); p = w.params()
x, y = syn.frame_poses(d, d["x_true"], d["y_true"])
t = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (x, y, d["frame_yaw_deg"], d["ranges"])]
g = torch.zeros((1, p.H, p.W), dtype=torch.int8, device=dev)
for _ in range(2):
    m.replay_dev(p, 1, w.n_frames, *(a.data_ptr() for a in t), g.data_ptr())
torch.cuda.synchronize(); print("ok")
