"""gpurun_out/<tag>_* (written by tools/profile_round.sh on the GPU box) -> profiles/<tag>_* (committed):
launch list + per-kernel share of the timed window, DRAM traffic of the fused step kernel, key ncu metrics."""
import collections, csv, json, os, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
commit = subprocess.run(["git", "-C", ROOT, "log", "-1", "--format=%h %s"], capture_output=True, text=True).stdout.strip()
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")

# ---- launch list
rows = [r for r in csv.reader(open(os.path.join(G, tag + "_launches.csv"))) if len(r) > 5]
hdr = rows[0]; ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
seq = [(r[ik].split("(")[0].replace("mjb::", "").replace("void ", ""), float(r[iv].replace(",", ""))) for r in rows[1:]]
step = [i for i, (n, _) in enumerate(seq) if n.startswith("k_step_warp")]
first, last = step[303], step[322]          # 300 settle + 3 warm-up launches precede the 20 timed steps
win = seq[first - 1:last + 2]
tot = sum(ns for _, ns in win)
by = collections.Counter()
for k, ns in win:
    by[k[:48]] += ns
share = {"command": "ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv python bench.py --steps 20 --warmup 3",
         "build": commit, "window": "the 20 timed steps (fused-step launches #303..#322 after 300 settle + 3 warm-up steps)",
         "per_kernel_ms": {k: round(v / 1e6, 4) for k, v in by.items()}, "share": {k: round(v / tot, 4) for k, v in by.items()},
         "fused_step_avg_ms": round(sum(v for k, v in by.items() if k.startswith("k_step_warp")) / 20 / 1e6, 4),
         "note": "per-launch times under ncu are cold-cache and serialised; the SHARE of the fused step kernel is what must agree with bench.py (launch_ms = ms_per_step)"}
json.dump(share, open(os.path.join(P, tag + "_launch_share.json"), "w"), indent=1)
shutil.copy(os.path.join(G, tag + "_launches.csv"), os.path.join(P, tag + "_launches.csv"))

# ---- full capture
raw = list(csv.reader(open(os.path.join(G, tag + "_step_full_raw.csv"))))
h, u, v = raw[0], raw[1], raw[2]
m = {a: (c, b) for a, b, c in zip(h, u, v)}
def val(k):
    return float(m[k][0].replace(",", "")) if k in m else float("nan")
def unit_scale(k):
    un = m[k][1] if k in m else ""
    return {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}.get(un, 1.0)
rd, wr = val("dram__bytes_read.sum") * unit_scale("dram__bytes_read.sum"), val("dram__bytes_write.sum") * unit_scale("dram__bytes_write.sum")
json.dump({"dram_bytes_per_launch": rd + wr, "dram_read_bytes": rd, "dram_write_bytes": wr, "kernel": m["Kernel Name"][0],
           "build": commit, "source": "profiles/%s_k_step_warp_full.md (ncu --set full --clock-control none, launch #316 of `bench.py --steps 20 --warmup 3`, B200)" % tag},
          open(os.path.join(P, tag + "_traffic.json"), "w"), indent=1)
stalls = sorted(((float(c), a) for a, c in zip(h, v) if a.startswith("smsp__average_warp") and "issue_stalled" in a and a.endswith(".ratio") and "not_issued" not in a), reverse=True)[:6]
keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_static",
        "launch__occupancy_limit_registers", "sm__maximum_warps_per_active_cycle_pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio"]
with open(os.path.join(P, tag + "_k_step_warp_full.md"), "w") as f:
    f.write("# %s - fused step kernel, full ncu capture (1xB200, sm_100a)\n\nBuild: `%s`\n\nCommand (under gpurun, one GPU): `ncu --set full --clock-control none --import-source on -k regex:k_step_warp -s 315 -c 1 "
            "python bench.py --steps 20 --warmup 3` (launch #316 = inside the timed window of the bench workload).\nRead with `ncu -i ... --page raw --csv` (tools/summarize_profile.py); the .ncu-rep is not committed.\n\n"
            "Kernel: `%s`\n\n| metric | value |\n|---|---|\n" % (tag, commit, m["Kernel Name"][0]))
    for k in keys:
        if k in m:
            f.write("| %s | %s %s |\n" % (k, m[k][0], m[k][1]))
    f.write("| DRAM traffic / launch | %.1f MB |\n" % ((rd + wr) / 1e6))
    f.write("\nTop issue-stall reasons (warps per issue-active cycle): " + ", ".join("%s %.2f" % (a.split("issue_stalled_")[1].split("_per_")[0], c) for c, a in stalls) + "\n")
for fn in (tag + "_bench.json", tag + "_bench_ref.json"):
    if os.path.exists(os.path.join(G, fn)):
        shutil.copy(os.path.join(G, fn), os.path.join(P, fn.replace("_bench.json", "_bench_n1.json").replace("_bench_ref.json", "_bench_reference_arm.json")))
print(json.dumps(share["share"]), rd + wr)
