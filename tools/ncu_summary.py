"""Summarise an .ncu-rep: headline raw metrics + hot SASS regions (needs ncu on PATH; no GPU).
usage: ncu_summary.py <report.ncu-rep> [full] [kernel=<regex>]   (kernel= restricts a multi-kernel report to one kernel)"""
import csv, io, subprocess, sys
rep = sys.argv[1]
ksel = [a.split("=", 1)[1] for a in sys.argv[2:] if a.startswith("kernel=")]
sys.argv = [a for a in sys.argv if not a.startswith("kernel=")]
kflt = ["-k", "regex:" + ksel[0]] if ksel else []
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"] + kflt, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sectors.sum", "lts__t_sectors.sum.per_second",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
for r in rows[2:]:
    print("== launch:", r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "")
    for h, u, v in zip(hdr, units, r):
        if h in keys or h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
            print(f"  {h} [{u}] = {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + kflt, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
hdr = rows[hi]
isrc, ie, it, isamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Avg. Predicated-On Threads Executed"), hdr.index("# Samples")
data = [r for r in rows[hi + 1:] if len(r) > ie and r[ie].isdigit()]
half = len(data) // 2
if half and len(data) % 2 == 0 and [r[isrc] + r[ie] for r in data[:half]] == [r[isrc] + r[ie] for r in data[half:]]:
    data = data[:half]          # a report holding two launches lists the module's source once per launch
tot = sum(int(r[ie]) for r in data) or 1
tots = sum(int(r[isamp]) for r in data) or 1
print(f"== SASS regions (total warp-instructions {tot}, samples {tots})")
cur, regions = None, []
for i, r in enumerate(data):
    e = int(r[ie])
    if e / tot < 0.0001:
        continue
    if cur and abs(e - cur["e"]) <= 0.05 * max(e, cur["e"]) and i == cur["end"] + 1:
        cur["n"] += 1; cur["sum"] += e; cur["samp"] += int(r[isamp]); cur["end"] = i; cur["thr"] += float(r[it])
    else:
        cur = {"start": i, "end": i, "e": e, "n": 1, "sum": e, "samp": int(r[isamp]), "thr": float(r[it]), "first": r[isrc].strip()[:48]}
        regions.append(cur)
for g in regions:
    print(f"  [{g['start']:4d}-{g['end']:4d}] n={g['n']:3d} exec={g['e']/1e6:9.2f}M inst%={100*g['sum']/tot:5.1f} stall-samples%={100*g['samp']/tots:5.1f} lanes={g['thr']/g['n']:4.1f}  {g['first']}")
if len(sys.argv) > 2:
    for i, r in enumerate(data):
        print(f"{i:4d} {int(r[ie])/1e6:9.2f}M s={100*int(r[isamp])/tots:5.2f}% thr={r[it]:>5} {r[isrc].strip()[:100]}")
