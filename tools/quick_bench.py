"""Scratch timing script for early GPU runs (replaced by bench.py)."""
import importlib, time, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
m = importlib.import_module("micro-quad-slam_b200"); syn = importlib.import_module("micro-quad-slam_b200.synth")
m.init(0)
dev = torch.device("cuda:0")
m.set_stream(torch.cuda.current_stream().cuda_stream)
print("rmw peak updates/s", m.measure_rmw_peak())
def run(w, tun=(0,0), reps=3):
    d = syn.generate(w); p = w.params()
    x, y = syn.frame_poses(d, d["x_true"], d["y_true"])
    tx, ty, tyaw, tr = (torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (x, y, d["frame_yaw_deg"], d["ranges"]))
    g = torch.zeros((w.n_flights, p.H, p.W), dtype=torch.int8, device=dev)
    m.set_tuning(tun[0], tun[1], 0)
    st = m.replay_dev(p, w.n_flights, w.n_frames, tx.data_ptr(), ty.data_ptr(), tyaw.data_ptr(), tr.data_ptr(), g.data_ptr(), want_stats=True)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); m.replay_dev(p, w.n_flights, w.n_frames, tx.data_ptr(), ty.data_ptr(), tyaw.data_ptr(), tr.data_ptr(), g.data_ptr()); e1.record()
        torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    U = st["ray_cell_updates"]
    print(f"{w.name} F={w.n_flights} N={w.n_frames} tun={tun}: {best:.2f} ms  U={U:.3e}  {U/best/1e6:.1f} G upd/s  frames/s={w.n_flights*w.n_frames/best*1e3:.3e}", flush=True)
for tun in [(0,0),(40,40),(56,56),(100,100),(200,40),(400,16)]:
    run(syn.scaled(syn.CONFIGS["c3"], n_flights=1024), tun)
run(syn.CONFIGS["c1"])
for tun in [(0,0),(48,48),(128,128)]:
    run(syn.scaled(syn.CONFIGS["c2"], n_samples=60000), tun)
run(syn.scaled(syn.CONFIGS["c4"], n_samples=100000))
