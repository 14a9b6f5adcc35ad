"""Small fixed workload for ncu captures: C3 slice, device-resident."""
import importlib, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
m = importlib.import_module("micro-quad-slam_b200"); syn = importlib.import_module("micro-quad-slam_b200.synth")
m.init(0)
dev = torch.device("cuda:0")
m.set_stream(torch.cuda.current_stream().cuda_stream)
nf = int(sys.argv[1]) if len(sys.argv) > 1 else 592
m.set_engine(int(sys.argv[2]) if len(sys.argv) > 2 else 0, int(sys.argv[3]) if len(sys.argv) > 3 else 0)
m.set_fan_layout(int(sys.argv[4]) if len(sys.argv) > 4 else -1)
w = syn.scaled(syn.CONFIGS["c3"], n_flights=nf)
d = syn.generate(w); p = w.params()
tx, ty, tyaw, tr = (torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (d["x_true"], d["y_true"], d["frame_yaw_deg"], d["ranges"]))
g = torch.zeros((w.n_flights, p.H, p.W), dtype=torch.int8, device=dev)
for _ in range(2):
    m.replay_dev(p, w.n_flights, w.n_frames, tx.data_ptr(), ty.data_ptr(), tyaw.data_ptr(), tr.data_ptr(), g.data_ptr())
torch.cuda.synchronize()
print("ok", int(g.to(torch.int64).sum().item()))
