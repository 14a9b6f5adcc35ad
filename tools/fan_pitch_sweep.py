"""Both lane layouts of the resident engine against the row pitch of the resident box (words mod 32) on config 3:
tools/banksim.py says the 8 beams x 4 steps layout needs a pitch of 5, 6, 7, 9 or 27 words (mod 32) to beat the 32-beam
layout; the built-in pitch for the ensemble's 248-cell box is 63 words (31 mod 32), where it is worse."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
m = importlib.import_module("micro-quad-slam_b200"); syn = importlib.import_module("micro-quad-slam_b200.synth")
m.init(0); dev = torch.device("cuda:0"); m.set_stream(torch.cuda.current_stream().cuda_stream)
w = syn.CONFIGS["c3"]; d = syn.generate(w); p = w.params()
t = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (d["x_true"], d["y_true"], d["frame_yaw_deg"], d["ranges"])]
g = torch.empty((w.n_flights, p.H, p.W), dtype=torch.int8, device=dev)
ref = None
for mod in (-1, 27, 1, 3, 5, 6, 7, 9, 11, 13, 15):
    for fan in (0, 1):
        m.set_engine(2, 8); m.set_fan_layout(fan); m.set_resident_pitch_mod(mod)
        try:
            st = m.replay_dev(p, w.n_flights, w.n_frames, *(a.data_ptr() for a in t), g.data_ptr(), want_stats=True)
        except m.UqsError as e:
            print(f"pitch mod {mod} fan {fan}: {e}"); continue
        m.set_profiling(True); m.profile_collect()
        for _ in range(3):
            m.replay_dev(p, w.n_flights, w.n_frames, *(a.data_ptr() for a in t), g.data_ptr())
        ms, cnt = m.profile_collect(); m.set_profiling(False)
        h = m.grid_hashes_dev(g.data_ptr(), w.n_flights, p.W * p.H).sum(dtype=np.uint64)
        ref = h if ref is None else ref
        print(f"pitch mod {mod:3d} fan {fan}: replay {ms[2]/3:.2f} ms  {'same bytes' if h == ref else 'DIFFERENT BYTES'}", flush=True)
m.set_engine(0, 0); m.set_fan_layout(-1); m.set_resident_pitch_mod(-1)
