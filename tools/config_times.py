"""Device-resident replay time of every BASELINE configuration at full size (ray set-up + replay kernels)."""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
m = importlib.import_module("micro-quad-slam_b200"); syn = importlib.import_module("micro-quad-slam_b200.synth")
m.init(0); dev = torch.device("cuda:0"); m.set_stream(torch.cuda.current_stream().cuda_stream)
def run(w, label, reps=3):
    d = syn.generate(w); p = w.params()
    x, y = syn.frame_poses(d, d["x_true"], d["y_true"])
    t = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (x, y, d["frame_yaw_deg"], d["ranges"])]
    g = torch.empty((w.n_flights, p.H, p.W), dtype=torch.int8, device=dev)
    st = m.replay_dev(p, w.n_flights, w.n_frames, *(a.data_ptr() for a in t), g.data_ptr(), want_stats=True)
    m.set_profiling(True); m.profile_collect()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); m.replay_dev(p, w.n_flights, w.n_frames, *(a.data_ptr() for a in t), g.data_ptr()); e1.record()
        torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    ms, cnt = m.profile_collect(); m.set_profiling(False)
    U = st["ray_cell_updates"]
    print(f"{label}: flights={w.n_flights} frames={w.n_frames} grid={p.W} U={U:.4e} time={best:.2f} ms "
          f"(setup {ms[1]/reps:.2f} + replay {ms[2]/reps:.2f}) -> {U/best/1e6:.1f} G updates/s, {w.n_flights*w.n_frames/best*1e3:.3e} frames/s", flush=True)
    return U, best
which = sys.argv[1:] or ["c1", "c2", "c3", "c4", "c5"]
for name in which:
    if name == "c5":
        # the 16 range-noise levels of one resolution share the grid geometry: one call of 16 x 64 flights each
        tot_u = tot_t = 0
        for ir in range(16):
            ws = [syn.c5_workload(ir, isg) for isg in range(16)]
            ds = [syn.generate(w) for w in ws]
            p = ws[0].params(); F = 64 * 16; N = ws[0].n_frames
            cat = lambda k: np.concatenate([d[k] for d in ds], axis=0)
            t = [torch.from_numpy(np.ascontiguousarray(cat(k))).to(dev) for k in ("x_true", "y_true", "frame_yaw_deg", "ranges")]
            g = torch.empty((F, p.H, p.W), dtype=torch.int8, device=dev)
            st = m.replay_dev(p, F, N, *(a.data_ptr() for a in t), g.data_ptr(), want_stats=True)
            best = 1e9
            for _ in range(2):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); m.replay_dev(p, F, N, *(a.data_ptr() for a in t), g.data_ptr()); e1.record()
                torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
            print(f"c5[res {ws[0].res} m, W={p.W}]: 16 noise levels x 64 flights: {best:.2f} ms, {st['ray_cell_updates']/best/1e6:.1f} G updates/s", flush=True)
            tot_u += st["ray_cell_updates"]; tot_t += best
        print(f"c5 all 256 configs x 64 flights: {tot_t:.1f} ms total, {tot_u/tot_t/1e6:.1f} G updates/s, U={tot_u:.4e}")
    else:
        run(syn.CONFIGS[name], name)
