"""C5 resolutions whose touched box fits only one resident CTA per SM: resident engine (8/16/32 warps) against the sub-tile engine."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
m = importlib.import_module("micro-quad-slam_b200"); syn = importlib.import_module("micro-quad-slam_b200.synth")
m.init(0); dev = torch.device("cuda:0"); m.set_stream(torch.cuda.current_stream().cuda_stream)
for ir in [int(a) for a in sys.argv[1:]] or [1, 2]:
    ws = [syn.c5_workload(ir, isg) for isg in range(16)]
    ds = [syn.generate(w) for w in ws]
    p = ws[0].params(); F = 64 * 16; N = ws[0].n_frames
    cat = lambda k: np.concatenate([d[k] for d in ds], axis=0)
    t = [torch.from_numpy(np.ascontiguousarray(cat(k))).to(dev) for k in ("x_true", "y_true", "frame_yaw_deg", "ranges")]
    g = torch.empty((F, p.H, p.W), dtype=torch.int8, device=dev)
    ref = None
    for eng, nw in [(0, 0), (2, 4), (2, 8), (2, 16), (2, 32)]:
        m.set_engine(eng, nw)
        try:
            st = m.replay_dev(p, F, N, *(a.data_ptr() for a in t), g.data_ptr(), want_stats=True)
        except Exception as e:
            print(f"W={p.W} engine={eng} nw={nw}: {e}"); continue
        if ref is None: ref = g.clone()
        same = bool(torch.equal(ref, g))
        best = 1e9
        for _ in range(2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); m.replay_dev(p, F, N, *(a.data_ptr() for a in t), g.data_ptr()); e1.record()
            torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
        print(f"W={p.W} engine={eng} nw={nw}: {best:.2f} ms {st['ray_cell_updates']/best/1e6:.1f} G upd/s same={same}", flush=True)
