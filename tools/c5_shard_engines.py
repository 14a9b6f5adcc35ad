"""Config 5 sharded over 8 / 4 GPUs leaves 128 / 256 flights per call: automatic engine choice against the resident
engine forced (2) with 8 and 16 warps, per resolution.  Decides the flight-count threshold of the automatic choice."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
m = importlib.import_module("micro-quad-slam_b200"); syn = importlib.import_module("micro-quad-slam_b200.synth")
m.init(0); dev = torch.device("cuda:0"); m.set_stream(torch.cuda.current_stream().cuda_stream)
for nfl in (128, 256):
    tot = {}
    for ir in range(16):
        w = syn.c5_workload(ir, 5, n_flights=nfl)
        d = syn.generate(w); p = w.params()
        t = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (d["x_true"], d["y_true"], d["frame_yaw_deg"], d["ranges"])]
        g = torch.empty((nfl, p.H, p.W), dtype=torch.int8, device=dev)
        row = []
        for eng, nw in ((0, 0), (1, 0), (2, 8), (2, 16)):
            m.set_engine(eng, nw)
            try:
                m.replay_dev(p, nfl, w.n_frames, *(a.data_ptr() for a in t), g.data_ptr())
            except m.UqsError:
                row.append(None); continue
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(2):
                m.replay_dev(p, nfl, w.n_frames, *(a.data_ptr() for a in t), g.data_ptr())
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 2
            row.append(ms); tot[(eng, nw)] = tot.get((eng, nw), 0.0) + ms
        print(f"flights={nfl} res={w.res} W={p.W}: auto {row[0]:.2f}  tiles {row[1]:.2f}  res8 {row[2] if row[2] is None else round(row[2],2)}  res16 {row[3] if row[3] is None else round(row[3],2)}", flush=True)
    print(f"flights={nfl} totals: " + ", ".join(f"{k}: {v:.1f} ms" for k, v in tot.items()), flush=True)
m.set_engine(0, 0)
