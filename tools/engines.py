"""Scratch: time both replay engines on C3 slices of several sizes (device-resident inputs)."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
m = importlib.import_module("micro-quad-slam_b200"); syn = importlib.import_module("micro-quad-slam_b200.synth")
m.init(0); dev = torch.device("cuda:0"); m.set_stream(torch.cuda.current_stream().cuda_stream)
for nf in [int(a) for a in sys.argv[1:]] or [1184]:
    w = syn.scaled(syn.CONFIGS["c3"], n_flights=nf); d = syn.generate(w); p = w.params()
    tx, ty, tyaw, tr = (torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (d["x_true"], d["y_true"], d["frame_yaw_deg"], d["ranges"]))
    g = torch.zeros((w.n_flights, p.H, p.W), dtype=torch.int8, device=dev)
    for eng, nw in [(1,0),(2,4),(2,8),(2,16)]:
        m.set_engine(eng, nw)
        st = m.replay_dev(p, w.n_flights, w.n_frames, tx.data_ptr(), ty.data_ptr(), tyaw.data_ptr(), tr.data_ptr(), g.data_ptr(), want_stats=True)
        best=1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); m.replay_dev(p, w.n_flights, w.n_frames, tx.data_ptr(), ty.data_ptr(), tyaw.data_ptr(), tr.data_ptr(), g.data_ptr()); e1.record()
            torch.cuda.synchronize(); best=min(best,e0.elapsed_time(e1))
        U=st["ray_cell_updates"]; print(f"F={nf} engine={eng} nw={nw}: {best:.2f} ms {U/best/1e6:.1f} G upd/s", flush=True)
