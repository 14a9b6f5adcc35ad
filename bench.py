#!/usr/bin/env python
"""bench.py -- ray-cell updates/s of the post-flight mapping path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c3|c1|c2|c4]

One "step" = one pass of the hot path (P0 pose integration -> ray set-up -> grid replay) over one
batch of synthetic logs.  Default workload: BASELINE config 3, the 4096-flight optical-flow drift
ensemble (3000 frames x 32 beams each, one 400x400 int8 grid per flight) -- the configuration the
metric is quoted on at 1/2/4/8 GPUs.  Flights are independent, so N GPUs run N x 4096 flights
(weak scaling, one process per GPU, no data-path collective).

Printed JSON line (rank 0):
  value        whole-job ray-cell updates/s, inputs already resident in HBM, CUDA-event timed
  e2e          same metric through the C ABI with HOST (pinned) buffers: H2D of the logs and D2H of
               the grids inside the timed region
  roofline     replay kernel: algorithmic bytes per launch / its event-timed duration vs the measured
               HBM copy bandwidth (MEASURED_PEAKS.json); plus the measured on-chip RMW ceiling
  cpu_baseline the reference's own mapping code (oracle/_ref) on the box's host cores, bounded sample
  --impl reference   times that CPU implementation as the main line (rank 0 only)
"""
from __future__ import annotations

import argparse
import ctypes
import importlib
import json
import multiprocessing as mp
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "ray_cell_updates_per_s"


# ------------------------------------------------------------------------------------------------
# workload
# ------------------------------------------------------------------------------------------------
def pick_workload(synth, name: str, flights):
    if name == "c3":
        w = synth.CONFIGS["c3"]
    elif name in synth.CONFIGS:
        w = synth.CONFIGS[name]
    else:
        raise SystemExit(f"unknown workload {name}")
    if flights:
        w = synth.scaled(w, n_flights=flights)
    return w


def describe(w, n_gpus):
    return {
        "workload": f"BASELINE config {w.config_id}: {w.name}",
        "flights_per_gpu": w.n_flights, "frames_per_flight": w.n_frames, "beams_per_frame": 32,
        "grid": f"{w.W}x{w.W}", "res_m": float(w.res), "global_flights": w.n_flights * n_gpus,
        "parallelism": f"flight-sharded x{n_gpus}, no collective",
        "l2_policy": "inputs (logs + ray records, >1.7 GB per GPU at full size) exceed the 126 MB L2; no flush needed",
    }


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's own mapping code, one forked worker per host core
# ------------------------------------------------------------------------------------------------
_CPU_JOB = {}          # inherited by forked workers (no pickling of the logs)


def _cpu_worker(c, cores, barrier, ret):
    J = _CPU_JOB
    from oracle import orc
    o = orc.Oracle()
    W, res = J["W"], J["res"]
    ref = orc.Reference(W, W, res) if J["kind"] == "reference" else None
    p = importlib.import_module("micro-quad-slam_b200").make_params(W, W, float(res))
    t_ms, rx, ry, h, yaw, q, ranges = (J[k] for k in ("t_ms", "of_rate_x", "of_rate_y", "h_m", "yaw_deg", "of_q", "ranges"))
    barrier.wait()
    t0 = time.perf_counter()
    for f in range(c, J["n"], cores):
        px, py = o.pose_integrate(t_ms[f], rx[f], ry[f], h[f], yaw[f], q[f])      # P0 (builder-defined), CPU statement
        if ref is not None:
            ref.reset(0.0, 0.0)
            ref.L.ref_replay(px.size, px.ctypes.data, py.ctypes.data, yaw[f].ctypes.data, ranges[f].ctypes.data, 0)
        else:
            o.replay(p, px, py, yaw[f], ranges[f])
    ret[c] = time.perf_counter() - t0


def cpu_replay_rate(w, d, updates_per_flight, cores, flights_per_core):
    """Replay cores*flights_per_core flights of workload w, one forked process per core (the reference's
    static grid forbids threads).  Returns (updates/s, frames/s, kind, flights, processes, wall seconds)."""
    from oracle import orc
    kind = "reference" if os.path.exists(orc.ref_lib_path(w.W, w.W, w.res)) else "port"
    n = min(w.n_flights, cores * flights_per_core)
    cores = min(cores, n)
    _CPU_JOB.clear()
    _CPU_JOB.update({k: d[k] for k in ("t_ms", "of_rate_x", "of_rate_y", "h_m", "yaw_deg", "of_q", "ranges")})
    _CPU_JOB.update(W=w.W, res=w.res, kind=kind, n=n)
    ctx = mp.get_context("fork")
    barrier = ctx.Barrier(cores)
    ret = ctx.Array("d", cores)
    procs = [ctx.Process(target=_cpu_worker, args=(c, cores, barrier, ret)) for c in range(cores)]
    for pr in procs:
        pr.start()
    for pr in procs:
        pr.join()
    if any(pr.exitcode != 0 for pr in procs):
        raise RuntimeError("CPU baseline worker failed")
    wall = max(ret[:])
    return updates_per_flight * n / wall, n * w.n_frames / wall, kind, n, cores, wall


# ------------------------------------------------------------------------------------------------
# clocks during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def bind_to_gpu_numa_node(gpu_index):
    """Run this rank (and its pinned allocations, by first touch) on the CPUs local to its GPU; on multi-socket
    boxes this keeps H2D/D2H traffic of the e2e path off the inter-socket link.  Best effort."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = {64 * i + b for i, w in enumerate(mask) for b in range(64) if (w >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def measured_hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def recorded_capture(w):
    """Figures of the committed ncu --set full capture of the replay kernel, when it is for this workload."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        if t.get("workload") == w.name and t.get("flights") == w.n_flights and t.get("frames") == w.n_frames:
            return t
    except Exception:
        pass
    return {}


def recorded_traffic(w):
    """dram bytes per replay launch from that capture (None when it is for another workload)."""
    return recorded_capture(w).get("dram_bytes_per_launch")


# ------------------------------------------------------------------------------------------------
def run_reference_arm(args, synth, out_fd):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = pick_workload(synth, args.workload, args.flights)
    cores = os.cpu_count() or 1
    fpc = max(1, min(args.cpu_flights_per_core, max(1, w.n_flights // cores)))
    n = min(w.n_flights, cores * fpc)
    ws = synth.scaled(w, n_flights=n)
    d = synth.generate(ws)
    from oracle import orc
    o = orc.Oracle()
    px, py = o.pose_integrate(d["t_ms"][0], d["of_rate_x"][0], d["of_rate_y"][0], d["h_m"][0], d["yaw_deg"][0], d["of_q"][0])
    _, U0 = o.replay(ws.params(), px, py, d["frame_yaw_deg"][0], d["ranges"][0])
    # per-flight update counts differ slightly (drift); count them all once with the port (untimed)
    U = 0
    for f in range(n):
        px, py = o.pose_integrate(d["t_ms"][f], d["of_rate_x"][f], d["of_rate_y"][f], d["h_m"][f], d["yaw_deg"][f], d["of_q"][f])
        U += o.replay(ws.params(), px, py, d["frame_yaw_deg"][f], d["ranges"][f])[1]
    upf = U / n
    times, kind, used = [], "port", cores
    for i in range(args.warmup + args.steps):
        ups, fps, kind, nn, used, wall = cpu_replay_rate(ws, d, upf, cores, fpc)
        if i >= args.warmup:
            times.append(wall)
    wall = sum(times) / len(times)
    value = U / wall
    sample = f"{n} of {w.n_flights} flights per step ({fpc} per worker process), P0 + mapping, logs in RAM"
    out_fd.emit(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "updates/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": wall * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "i8", "data": "synthetic", "config": describe(w, args.gpus), "frames_per_s": n * w.n_frames / wall,
        "cpu_baseline": {"value": value, "unit": "updates/s", "cores": used, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


class OneLineStdout:
    """Everything any library prints to fd 1 (NCCL's version banner, ...) goes to stderr; only emit() writes to the
    real stdout -- the driver reads ONE JSON line from it."""

    def __init__(self):
        sys.stdout.flush()
        self.real = os.dup(1)
        os.dup2(2, 1)

    def emit(self, line: str):
        sys.stdout.flush()
        os.write(self.real, (line + "\n").encode())


def main():
    out_fd = OneLineStdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--flights", type=int, default=0, help="flights per GPU (default: the config's)")
    ap.add_argument("--cpu-flights-per-core", type=int, default=48)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the single-flight configs (c1, c2) reported for information")
    ap.add_argument("--engine", type=int, default=0, help="0 auto, 1 sub-tiles, 2 resident (tuning experiments)")
    ap.add_argument("--flight-warps", type=int, default=0, help="warps per resident CTA (0 = library default)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    m = importlib.import_module("micro-quad-slam_b200")
    synth = importlib.import_module("micro-quad-slam_b200.synth")
    if args.impl == "reference":
        run_reference_arm(args, synth, out_fd)
        return

    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with python -m torch.distributed.run "
                         f"--nnodes=1 --nproc-per-node {args.gpus} --master-addr 127.0.0.1 bench.py --gpus {args.gpus}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    local_cpus = bind_to_gpu_numa_node(local) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    m.init(local)                                   # fails loudly if the CUDA library cannot run
    m.set_stream(torch.cuda.current_stream().cuda_stream)
    m.set_engine(args.engine, args.flight_warps)

    w = pick_workload(synth, args.workload, args.flights)
    p = w.params()
    F, N, NF = w.n_flights, w.n_samples, w.n_frames

    # ---- synthetic logs, generated straight into pinned host memory --------------------------------
    def pinned(shape, dt):
        return torch.empty(shape, dtype=dt, pin_memory=True)
    host = {"t_ms": pinned((F, N), torch.int32), "of_rate_x": pinned((F, N), torch.float32), "of_rate_y": pinned((F, N), torch.float32),
            "h_m": pinned((F, N), torch.float32), "yaw_deg": pinned((F, N), torch.float32), "of_q": pinned((F, N), torch.uint8),
            "ranges": pinned((F, NF, 32), torch.float32), "x_true": pinned((F, N), torch.float32), "y_true": pinned((F, N), torch.float32)}
    views = {k: v.numpy() for k, v in host.items()}
    views["t_ms"] = views["t_ms"].view(np.uint32)
    d = synth.generate(w, flight_id0=rank * F, out=dict(views))
    if w.frames_per_sample != 1:
        raise SystemExit("bench: flow-driven workloads only (one frame per sample)")
    h_grids = pinned((F, p.H, p.W), torch.int8)

    dv = {k: host[k].to(dev, non_blocking=True) for k in ("t_ms", "of_rate_x", "of_rate_y", "h_m", "yaw_deg", "of_q", "ranges")}
    d_x = torch.empty((F, N), dtype=torch.float32, device=dev)
    d_y = torch.empty((F, N), dtype=torch.float32, device=dev)
    d_grids = torch.empty((F, p.H, p.W), dtype=torch.int8, device=dev)
    torch.cuda.synchronize()

    def step_device(want_stats=False):
        m.pose_integrate_dev(F, N, dv["t_ms"].data_ptr(), dv["of_rate_x"].data_ptr(), dv["of_rate_y"].data_ptr(), dv["h_m"].data_ptr(),
                             dv["yaw_deg"].data_ptr(), dv["of_q"].data_ptr(), d_x.data_ptr(), d_y.data_ptr(), 0)
        return m.replay_dev(p, F, NF, d_x.data_ptr(), d_y.data_ptr(), dv["yaw_deg"].data_ptr(), dv["ranges"].data_ptr(),
                            d_grids.data_ptr(), want_stats=want_stats)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    st = step_device(want_stats=True)
    U = st["ray_cell_updates"]
    for _ in range(args.warmup - 1):
        step_device()
    barrier()

    # ---- timed: K steps, device-resident inputs ---------------------------------------------------------
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    m.set_profiling(True)
    m.profile_collect()
    l0 = m.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step_device()
    e1.record()
    barrier()
    launches = m.kernel_launches() - l0
    kms, kcnt = m.profile_collect()
    m.set_profiling(False)
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    tot_U = torch.tensor([float(U)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot_U)
    ms_per_step = float(ms.item()) / args.steps
    total_updates = float(tot_U.item())
    value = total_updates / (ms_per_step * 1e-3)

    # ---- e2e: the C-ABI call with host buffers, H2D + D2H inside the timed region -----------------------------
    e2e = None
    if not args.no_e2e:
        def step_e2e():
            m.replay_flow(p, d["t_ms"], d["of_rate_x"], d["of_rate_y"], d["h_m"], d["yaw_deg"], d["of_q"], d["ranges"],
                          want_poses=False, out=h_grids.numpy())
        step_e2e()
        barrier()
        t0 = time.perf_counter()
        e0.record()
        for _ in range(args.steps):
            step_e2e()
        e1.record()
        barrier()
        ems = torch.tensor([max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3)], device=dev)
        if world > 1:
            dist.all_reduce(ems, op=dist.ReduceOp.MAX)
        e_ms = float(ems.item()) / args.steps
        h2d = sum(d[k].nbytes for k in ("t_ms", "of_rate_x", "of_rate_y", "h_m", "yaw_deg", "of_q", "ranges"))
        e2e = {"value": total_updates / (e_ms * 1e-3), "unit": "updates/s", "ms_per_step": e_ms,
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(h_grids.numel()),
               "api": "uqs_replay_flow (host pointers, pinned)"}
        # the e2e grids must equal the device-resident ones
        if not torch.equal(h_grids, d_grids.cpu()):
            raise SystemExit("bench: e2e grids differ from device-resident grids")
    clk = clocks.stop() if rank == 0 else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (replay) --------------------------------------------------------------
    peak, peak_src = measured_hbm_peak()
    n_rays = F * NF * 32
    b_alg = 2 * U + 8 * n_rays + 16 * F * NF + p.W * p.H * F          # per replay launch(es) of one step, this rank
    replay_ms = kms[2] / max(args.steps, 1)                           # summed over the step's replay launches
    achieved = b_alg / (replay_ms * 1e-3) / 1e9
    rmw_peak = m.measure_rmw_peak()
    cap = recorded_capture(w)
    roofline = {"bound": "hbm", "kernel": "k_replay_tiles/k_replay_flights", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "peak_source": peak_src, "traffic": recorded_traffic(w),
                "algorithmic_bytes_per_step": int(b_alg), "kernel_ms_per_step": replay_ms,
                "kernel_share_of_step": replay_ms / ms_per_step,
                "setup_kernel_ms_per_step": kms[1] / args.steps, "pose_kernels_ms_per_step": kms[0] / args.steps,
                "smem_pipe_recorded": {"pct_of_peak": cap.get("smem_pipe_pct_of_peak"), "issue_slots_pct": cap.get("issue_slots_pct_of_peak"),
                                       "note": "shared-memory wavefronts of the replay kernel in the committed ncu capture "
                                               "(profiles/r1_ncu_full_k_replay_flights_c3.txt): the pipe this kernel is bound by"},
                "onchip_rmw": {"achieved_updates_per_s": U / (replay_ms * 1e-3), "peak_updates_per_s": rmw_peak,
                               "frac": U / (replay_ms * 1e-3) / rmw_peak,
                               "note": "peak = conflict-free shared-memory byte RMW microbenchmark on this GPU"}}

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        cores = os.cpu_count() or 1
        fpc = max(1, min(args.cpu_flights_per_core, max(1, F // cores)))
        ns = min(F, cores * fpc)
        # pageable copies of the sample: CUDA-pinned pages are not inherited by forked workers
        ds = {k: np.array(d[k][:ns]) for k in ("t_ms", "of_rate_x", "of_rate_y", "h_m", "yaw_deg", "of_q", "ranges")}
        ups, fps, kind, n, used, wall = cpu_replay_rate(synth.scaled(w, n_flights=ns), ds, U / F, cores, fpc)
        cpu = {"value": ups, "unit": "updates/s", "cores": used, "kind": kind, "frames_per_s": fps,
               "sample": f"{n} of {F} flights ({fpc} per worker process), P0 + mapping, logs in RAM, {wall:.2f} s wall"}

    # ---- the single-flight BASELINE configurations, device-resident (information only; parity: tests/test_gpu_fullsize.py) ----
    others = []
    if world == 1 and args.workload == "c3" and not args.flights and not args.no_other_configs:
        for name in ("c1", "c2"):
            wo = synth.CONFIGS[name]
            do = synth.generate(wo)
            po = wo.params()
            xo, yo = synth.frame_poses(do, do["x_true"], do["y_true"])
            to = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (xo, yo, do["frame_yaw_deg"], do["ranges"])]
            go = torch.empty((wo.n_flights, po.H, po.W), dtype=torch.int8, device=dev)
            so = m.replay_dev(po, wo.n_flights, wo.n_frames, *(a.data_ptr() for a in to), go.data_ptr(), want_stats=True)
            for _ in range(2):
                m.replay_dev(po, wo.n_flights, wo.n_frames, *(a.data_ptr() for a in to), go.data_ptr())
            torch.cuda.synchronize()
            e0.record()
            for _ in range(args.steps):
                m.replay_dev(po, wo.n_flights, wo.n_frames, *(a.data_ptr() for a in to), go.data_ptr())
            e1.record()
            torch.cuda.synchronize()
            t_ms = e0.elapsed_time(e1) / args.steps
            others.append({"workload": f"BASELINE config {wo.config_id}: {wo.name}", "frames": wo.n_frames, "grid": f"{po.W}x{po.H}",
                           "ms_per_replay": t_ms, "updates_per_s": so["ray_cell_updates"] / (t_ms * 1e-3),
                           "frames_per_s": wo.n_frames / (t_ms * 1e-3), "what": "ray set-up + replay from poses in HBM"})

    out = {"metric": METRIC, "value": value, "unit": "updates/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "i8",
           "data": "synthetic", "config": describe(w, world), "frames_per_s": F * NF * world / (ms_per_step * 1e-3),
           "ray_cell_updates_per_step": total_updates, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
           "gpu_launches": int(launches), "clocks": clk, "host": {"cpus": os.cpu_count(), "rank_local_cpus": local_cpus},
           "other_configs": others}
    out_fd.emit(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
