"""Regenerates the round-2 table of profiles/README.md (between the R2_TABLE markers) from profiles/r2_bench_n*.json."""
import json, os, re
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")
D = {n: json.load(open(os.path.join(P, f"r2_bench_n{n}.json"))) for n in (1, 2, 4, 8)}
ref = json.load(open(os.path.join(P, "r2_bench_n1_reference_arm.json")))
ref8 = json.load(open(os.path.join(P, "r2_bench_n8_reference_arm.json")))
hdr = "| | N=1 | N=2 | N=4 | N=8 |\n|---|---|---|---|---|"
row = lambda name, fn: "| " + name + " | " + " | ".join(fn(D[n]) for n in (1, 2, 4, 8)) + " |"
v1 = D[1]["value"]
ok = lambda d: d["hash"] == d["hash_n1"] and all(c.get("hash_matches_n1", True) for c in d["configs"].values())
lines = [hdr,
 row("`value`: device-resident updates/s (ms/step)", lambda d: f"**{d['value']:.3e}** ({d['ms_per_step']:.2f} ms)"),
 row("strong-scaling efficiency of `value`", lambda d: f"{d['value']/v1/d['n_gpus']:.2f}"),
 row("`e2e`: `uqs_replay_flow`, float ranges, dense grids (ms/step)", lambda d: f"{d['e2e']['value']:.3e} ({d['e2e']['ms_per_step']:.1f} ms)"),
 row("`e2e.copy_floor_ms`: the same call, kernels off", lambda d: f"{d['e2e']['copy_floor_ms']:.1f} ms"),
 row("`e2e_variants.mm`: u16 millimetre ranges in", lambda d: f"{d['e2e_variants']['mm']['value']:.3e} ({d['e2e_variants']['mm']['ms_per_step']:.1f} ms)"),
 row("`e2e_variants.mm_boxed`: + touched boxes out", lambda d: f"{d['e2e_variants']['mm_boxed']['value']:.3e} ({d['e2e_variants']['mm_boxed']['ms_per_step']:.1f} ms)"),
 row("`hash == hash_n1` (every config)", lambda d: "—" if d["n_gpus"] == 1 else ("yes" if ok(d) else "NO")),
 row("`configs.c3_weak`: 4096 flights per GPU", lambda d: "= headline" if d["n_gpus"] == 1 else f"{d['configs']['c3_weak']['value']:.3e} ({d['configs']['c3_weak']['ms_per_step']:.1f} ms; e2e {d['configs']['c3_weak']['e2e']['ms_per_step']:.0f} ms, floor {d['configs']['c3_weak']['e2e']['copy_floor_ms']:.0f})"),
 row("`configs.c1`: one 60 s flight, P0 included", lambda d: f"{d['configs']['c1']['ms_per_step']:.2f} ms"),
 row("`configs.c2`: one-hour log, P0 included", lambda d: f"{d['configs']['c2']['ms_per_step']:.1f} ms ({d['configs']['c2']['value']:.2e})"),
 row("`configs.c4`: 16384² grid, owned bands + NCCL gather", lambda d: f"{d['configs']['c4']['ms_per_step']:.1f} ms ({d['configs']['c4']['value']:.2e}); e2e {d['configs']['c4']['e2e']['ms_per_step']:.1f} ms"),
 row("`configs.c5`: 256 configs × 64 flights", lambda d: f"{d['configs']['c5']['ms_per_step']:.1f} ms ({d['configs']['c5']['value']:.2e}); e2e {d['configs']['c5']['e2e']['ms_per_step']:.0f} ms"),
]
r = D[1]["roofline"]
c48 = D[8]["configs"]["c4"]
extra = f"""
N=1 details: replay kernel {r['kernel_ms_per_step']:.2f} ms ({100*r['kernel_share_of_step']:.1f} % of the step), ray set-up {r['setup_kernel_ms_per_step']:.2f} ms, P0 {r['pose_kernels_ms_per_step']:.2f} ms;
`roofline` vs the measured HBM copy bandwidth: {r['achieved']:.0f} GB/s algorithmic → **{r['frac']:.3f}** (DRAM traffic per launch {r['traffic']/1e9:.2f} GB: HBM is not the limiter);
on-chip: {r['onchip_rmw']['achieved_updates_per_s']:.3e} updates/s in the replay kernel against {r['onchip_rmw']['peak_updates_per_s']:.2e} for the conflict-free byte load/clamp/store ceiling → **{r['onchip_rmw']['frac']:.2f}**
(`ATOMS.ADD` ceiling {r['onchip_rmw']['atomics_updates_per_s']:.2e}); clocks {D[1]['clocks']['sm_mhz']:.0f} of {D[1]['clocks']['sm_max_mhz']:.0f} MHz, no throttle reason;
`cpu_baseline` (the reference's own code, {D[1]['cpu_baseline']['cores']} host cores): {D[1]['cpu_baseline']['value']:.2e} updates/s; `--impl reference`: {ref['value']:.2e} on that box, {ref8['value']:.2e} on the {ref8['cpu_baseline']['cores']}-core host of the 8-GPU box.
Config 4 on eight GPUs: kernels per rank {c48['kernel_ms_per_rank']} ms, step without the gather {c48['ms_per_step_without_gather']:.2f} ms, with it {c48['ms_per_step']:.2f} ms, cuts {D[8]['configs']['c4']['partition'].split('cut at rows ')[1].split(' (')[0]}.
N=1 from a 1-GPU box, N=2 from a 2-GPU box, N=4 and N=8 from one 8-GPU box (`r2_bench_n*.json`), all on the final build.
"""
p = os.path.join(P, "README.md")
s = open(p).read()
block = "<!-- R2_TABLE_BEGIN -->\n" + "\n".join(lines) + "\n" + extra + "<!-- R2_TABLE_END -->"
if "<!-- R2_TABLE_BEGIN -->" in s:
    s = re.sub(r"<!-- R2_TABLE_BEGIN -->.*?<!-- R2_TABLE_END -->", lambda m: block, s, flags=re.S)
else:
    a = s.index("| | N=1 | N=2 | N=4 | N=8 |")
    b = s.index("### What changed in round 2")
    s = s[:a] + block + "\n\n" + s[b:]
open(p, "w").write(s)
print(block)
