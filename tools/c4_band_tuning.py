"""Config 4 sharded over 8 GPUs, emulated band by band on one GPU: time of every rank's owned row band (ray set-up +
replay of the band, uqs_replay_dev row0/rows) under several sub-tile / time-slice tunings.  The N-GPU step is the
slowest band (+ the gather), so the tuning that minimises the maximum is the one the banded replay should pick."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
m = importlib.import_module("micro-quad-slam_b200"); syn = importlib.import_module("micro-quad-slam_b200.synth")
m.init(0); dev = torch.device("cuda:0"); m.set_stream(torch.cuda.current_stream().cuda_stream)
w = syn.on_mm_lattice(syn.CONFIGS["c4"]); d = syn.generate(w); p = w.params()
x, y = syn.frame_poses(d, d["x_true"], d["y_true"])
t = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (x, y, d["frame_yaw_deg"], d["ranges"])]
g = torch.empty((1, p.H, p.W), dtype=torch.int8, device=dev)
N = w.n_frames
for world in (8, 4, 2):
    edges = m.balanced_row_bands_dev(p, N, t[0].data_ptr(), t[1].data_ptr(), world)
    print(f"world={world} balanced cuts {edges}", flush=True)
    for label, (sw, sh, sl) in (("auto", (0, 0, 0)), ("80x100 no slices", (80, 100, 1)), ("56x56 no slices", (56, 56, 1)), ("40x48 no slices", (40, 48, 1)),
                                ("32x32 no slices", (32, 32, 1)), ("56x56 2 slices", (56, 56, 2)), ("40x40 3 slices", (40, 40, 3))):
        m.set_tuning(sw, sh, sl)
        times = []
        for a, b in zip(edges, edges[1:]):
            m.replay_dev(p, 1, N, *(q.data_ptr() for q in t), g.data_ptr(), row0=a, rows=b - a)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); m.replay_dev(p, 1, N, *(q.data_ptr() for q in t), g.data_ptr(), row0=a, rows=b - a); e1.record()
            torch.cuda.synchronize(); times.append(e0.elapsed_time(e1))
        print(f"  {label:18s}: max {max(times):6.2f} ms  sum {sum(times):6.2f}  bands " + " ".join(f"{v:.1f}" for v in times), flush=True)
m.set_tuning(0, 0, 0)
