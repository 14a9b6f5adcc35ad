"""A/B of the resident engine's lane layouts (uqs_set_fan_layout) on config 3 and on the config-5 geometries:
replay-kernel time (events on the launching stream), identical digests required."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
m = importlib.import_module("micro-quad-slam_b200"); syn = importlib.import_module("micro-quad-slam_b200.synth")
m.init(0); dev = torch.device("cuda:0"); m.set_stream(torch.cuda.current_stream().cuda_stream)

def run(w, label, warps=(0, 4, 8, 16), reps=3):
    d = syn.generate(w); p = w.params()
    t = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (d["x_true"], d["y_true"], d["frame_yaw_deg"], d["ranges"])]
    g = torch.empty((w.n_flights, p.H, p.W), dtype=torch.int8, device=dev)
    ref = None
    for nw in warps:
        for fan in (0, 2):                                   # 0 = shared decode, 2 = dedicated decode warp (1 = fan layout: see r2_fan_layout_ab.log)
            m.set_engine(2 if nw else 0, nw); m.set_fan_layout(1 if fan == 1 else 0); m.set_decode_warp(1 if fan == 2 else 0)
            try:
                st = m.replay_dev(p, w.n_flights, w.n_frames, *(a.data_ptr() for a in t), g.data_ptr(), want_stats=True)
            except m.UqsError as e:
                print(f"{label} nw={nw} variant={fan}: {e}"); continue
            m.replay_dev(p, w.n_flights, w.n_frames, *(a.data_ptr() for a in t), g.data_ptr())
            m.set_profiling(True); m.profile_collect()
            for _ in range(reps):
                m.replay_dev(p, w.n_flights, w.n_frames, *(a.data_ptr() for a in t), g.data_ptr())
            ms, cnt = m.profile_collect(); m.set_profiling(False)
            h = m.grid_hashes_dev(g.data_ptr(), w.n_flights, p.W * p.H).sum(dtype=np.uint64)
            ref = h if ref is None else ref
            print(f"{label} nw={nw} variant={fan}: setup {ms[1]/reps:.2f} ms replay {ms[2]/reps:.2f} ms  {st['ray_cell_updates']/(ms[2]/reps)/1e6:.0f} G upd/s  "
                  f"{'same bytes' if h == ref else 'DIFFERENT BYTES'}", flush=True)
    m.set_engine(0, 0); m.set_fan_layout(-1); m.set_decode_warp(-1)

which = sys.argv[1:] or ["c3", "c5"]
if "c3" in which:
    run(syn.CONFIGS["c3"], "c3 4096 flights")
    run(syn.scaled(syn.CONFIGS["c3"], n_flights=512), "c3 512 flights", warps=(0, 8, 16))
if "c5" in which:
    for ir in (0, 2, 4, 6, 9, 12, 15):
        w = syn.c5_workload(ir, 5, n_flights=1024)
        run(w, f"c5 res {w.res} W={w.W} 1024 flights", warps=(0,), reps=2)
