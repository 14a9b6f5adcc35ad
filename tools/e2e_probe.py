"""Scratch: where does the host-buffer pipeline spend its time?"""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
m = importlib.import_module("micro-quad-slam_b200"); syn = importlib.import_module("micro-quad-slam_b200.synth")
m.init(0); dev = torch.device("cuda:0")
F = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
w = syn.scaled(syn.CONFIGS["c3"], n_flights=F); p = w.params(); N = w.n_samples
def pinned(shape, dt): return torch.empty(shape, dtype=dt, pin_memory=True)
host = {"t_ms": pinned((F, N), torch.int32), "of_rate_x": pinned((F, N), torch.float32), "of_rate_y": pinned((F, N), torch.float32),
        "h_m": pinned((F, N), torch.float32), "yaw_deg": pinned((F, N), torch.float32), "of_q": pinned((F, N), torch.uint8),
        "ranges": pinned((F, N, 32), torch.float32), "x_true": pinned((F, N), torch.float32), "y_true": pinned((F, N), torch.float32)}
views = {k: v.numpy() for k, v in host.items()}; views["t_ms"] = views["t_ms"].view(np.uint32)
d = syn.generate(w, out=dict(views))
print("pinned reused:", all(d[k].ctypes.data == views[k].ctypes.data for k in views))
hg = pinned((F, p.H, p.W), torch.int8)
# raw copy bandwidths
t = torch.empty_like(host["ranges"], device=dev); torch.cuda.synchronize()
for _ in range(2):
    t0 = time.perf_counter(); t.copy_(host["ranges"], non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"H2D ranges {host['ranges'].numel()*4/1e9:.2f} GB in {dt*1e3:.1f} ms = {host['ranges'].numel()*4/dt/1e9:.1f} GB/s")
g = torch.empty((F, p.H, p.W), dtype=torch.int8, device=dev)
for _ in range(2):
    t0 = time.perf_counter(); hg.copy_(g, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"D2H grids {g.numel()/1e9:.2f} GB in {dt*1e3:.1f} ms = {g.numel()/dt/1e9:.1f} GB/s")
for stream_mode in ("torch",):
    m.set_stream(None if stream_mode == "own" else torch.cuda.current_stream().cuda_stream)
    for chunk in (740, 888, 0, 0):
        m.set_host_chunk(chunk)
        for rep in range(2):
            t0 = time.perf_counter()
            m.replay_flow(p, d["t_ms"], d["of_rate_x"], d["of_rate_y"], d["h_m"], d["yaw_deg"], d["of_q"], d["ranges"], want_poses=False, out=hg.numpy())
            dt = time.perf_counter() - t0
        print(f"stream={stream_mode} chunk={chunk}: {dt*1e3:.1f} ms", flush=True)
