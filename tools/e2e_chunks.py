"""Host-buffer replay time (uqs_replay_flow, C3 ensemble) against the pipeline's chunk size."""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
m = importlib.import_module("micro-quad-slam_b200"); syn = importlib.import_module("micro-quad-slam_b200.synth")
m.init(0); dev = torch.device("cuda:0")
F = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
w = syn.scaled(syn.CONFIGS["c3"], n_flights=F); p = w.params(); N = w.n_samples
def pinned(shape, dt): return torch.empty(shape, dtype=dt, pin_memory=True)
host = {"t_ms": pinned((F, N), torch.int32), "of_rate_x": pinned((F, N), torch.float32), "of_rate_y": pinned((F, N), torch.float32),
        "h_m": pinned((F, N), torch.float32), "yaw_deg": pinned((F, N), torch.float32), "of_q": pinned((F, N), torch.uint8),
        "ranges": pinned((F, N, 32), torch.float32), "x_true": pinned((F, N), torch.float32), "y_true": pinned((F, N), torch.float32)}
views = {k: v.numpy() for k, v in host.items()}; views["t_ms"] = views["t_ms"].view(np.uint32)
d = syn.generate(w, out=dict(views))
hg = pinned((F, p.H, p.W), torch.int8)
m.set_stream(torch.cuda.current_stream().cuda_stream)
def run():
    t0 = time.perf_counter()
    m.replay_flow(p, d["t_ms"], d["of_rate_x"], d["of_rate_y"], d["h_m"], d["yaw_deg"], d["of_q"], d["ranges"], want_poses=False, out=hg.numpy())
    return (time.perf_counter() - t0) * 1e3
for chunk in [int(a) for a in sys.argv[2:]] or [0]:
    m.set_host_chunk(chunk)
    run()
    print(f"chunk={chunk}: " + " ".join(f"{run():.1f}" for _ in range(4)) + " ms", flush=True)
