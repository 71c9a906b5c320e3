"""gpurun_out/r02_* (written on the GPU box by tools/gpu_profiles_r02.sh) -> profiles/r02_*: launch shares of the bench
command, one record per kernel of the split step from the full ncu capture, the table in r02_split_full.md"""
import csv, json, os, shutil, subprocess, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
build = subprocess.run(["git", "rev-parse", "--short", "HEAD"], cwd=ROOT, capture_output=True, text=True).stdout.strip()

# ---- launch list -> shares
rows = [r for r in csv.reader(open(os.path.join(G, "r02_launches.csv"), errors="replace")) if len(r) > 14 and r[0].isdigit()]
by = collections.defaultdict(list)
for r in rows:
    by[r[4]].append(float(r[14]) / 1e3)
tot = sum(sum(v) for v in by.values())
share = {"source": "ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1500 -c 400 python bench.py --steps 20 --warmup 3 "
                   "(serialised, cold-cache times: shares, not absolutes), build " + build,
         "kernels": {k: {"launches": len(v), "mean_us": sum(v) / len(v), "share_of_captured_time": sum(v) / tot} for k, v in by.items()}}
json.dump(share, open(os.path.join(P, "r02_launch_share.json"), "w"), indent=1)
shutil.copy(os.path.join(G, "r02_launches.csv"), os.path.join(P, "r02_launches.csv"))

# ---- full capture -> one record per kernel
raw = list(csv.reader(open(os.path.join(G, "r02_split_full_raw.csv"), errors="replace")))
hdr, units = raw[0], raw[1]
col = {h: i for i, h in enumerate(hdr)}
SCALE = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}   # -> us, bytes
def num(r, name):
    if name not in col: return None
    try: return float(r[col[name]].replace(",", "")) * SCALE.get(units[col[name]], 1.0)
    except ValueError: return None
out = {"source": "ncu --set full --clock-control none, one capture per kernel after 300 settle steps (tools/gpu_profiles_r02.sh), build " + build}
keyof = {"<0, 32, 0, 0>": "redo", "k_set_control": "set_control", "<0, 16, 0, 1>": "part1", "k_pgs4": "pgs4", "<0, 16, 0, 2>": "part2"}
for r in raw[2:]:
    name = r[col["Kernel Name"]]
    key = next((v for k, v in keyof.items() if k in name), None)
    if key is None or key in out: continue
    stalls = sorted(((h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")], num(r, h)) for h in hdr
                     if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and num(r, h) is not None),
                    key=lambda t: -t[1])[:3]
    inst = num(r, "smsp__inst_executed.sum")
    thr = num(r, "smsp__thread_inst_executed.sum")
    out[key] = {
        "kernel": name, "duration_us": num(r, "gpu__time_duration.sum") or 0, "warp_instructions": inst,
        "issue_active": (num(r, "smsp__issue_active.avg.pct_of_peak_sustained_active") or 0) / 100,
        "threads_per_instruction": (thr / inst) if inst and thr else num(r, "smsp__thread_inst_executed_per_inst_executed.ratio"),
        "warps_active_pct": num(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
        "sm_cycles_active_frac": (num(r, "sm__cycles_active.avg") or 0) / max(1.0, num(r, "sm__cycles_elapsed.max") or num(r, "sm__cycles_elapsed.avg") or 1.0),
        "fp64_pipe_frac": (num(r, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active") or 0) / 100,
        "l1_wavefront_pipe_frac": (num(r, "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed") or 0) / 100,
        "dram_bytes_per_launch": (num(r, "dram__bytes_read.sum") or 0) + (num(r, "dram__bytes_write.sum") or 0),
        "l1_hit_pct": num(r, "l1tex__t_sector_hit_rate.pct"), "l2_hit_pct": num(r, "lts__t_sector_hit_rate.pct"),
        "registers": num(r, "launch__registers_per_thread"),
        "local_ld_requests": num(r, "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum"),
        "local_st_requests": num(r, "l1tex__t_requests_pipe_lsu_mem_local_op_st.sum"),
        "top_stalls": stalls}
json.dump(out, open(os.path.join(P, "r02_kernels.json"), "w"), indent=1)

# ---- the table
L = ["# round 2: full ncu capture of the kernels of one split step (humanoid x4096, PGS, after 300 settle steps), build " + build, "",
     "| kernel | duration us | warp instr | issue active % | lanes/instr | warps active % of peak | SM active / elapsed | fp64 pipe % | L1 wavefront pipe % | DRAM rd+wr MB | L1 hit % | L2 hit % | regs | top stalls (warps per issue) |",
     "|---|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
for key in ("redo", "set_control", "part1", "pgs4", "part2"):
    k = out.get(key)
    if not k: continue
    f = lambda v, s="%.1f": "-" if v is None else s % v
    L.append("| %s | %.1f | %.2e | %.1f | %s | %s | %.2f | %.1f | %.1f | %.1f | %s | %s | %d | %s |" % (
        key, k["duration_us"], k["warp_instructions"] or 0, 100 * k["issue_active"], f(k["threads_per_instruction"]), f(k["warps_active_pct"]),
        k["sm_cycles_active_frac"], 100 * k["fp64_pipe_frac"], 100 * k["l1_wavefront_pipe_frac"], k["dram_bytes_per_launch"] / 1e6, f(k["l1_hit_pct"]),
        f(k["l2_hit_pct"]), int(k["registers"] or 0), ", ".join("%s %.2f" % t for t in k["top_stalls"])))
L += ["", "Local-memory traffic (spills and lane-private arrays), `l1tex__t_requests_pipe_lsu_mem_local_op_{ld,st}.sum`:", "",
      "| kernel | local loads | local stores | share of warp instructions |", "|---|---|---|---|"]
for key in ("part1", "pgs4", "part2"):
    k = out.get(key)
    if not k: continue
    ld, st = k["local_ld_requests"] or 0, k["local_st_requests"] or 0
    L.append("| %s | %.2e | %.2e | %.2f %% |" % (key, ld, st, 100 * (ld + st) / max(1.0, k["warp_instructions"] or 1)))
dram = sum(out[k]["dram_bytes_per_launch"] for k in ("part1", "pgs4", "part2") if k in out)
L += ["", "DRAM per step (three kernels): %.1f MB." % (dram / 1e6)]
open(os.path.join(P, "r02_split_full.md"), "w").write("\n".join(L) + "\n")
for f in ("r02_bench_n1.json", "r02_bench_n1_s20.json", "r02_bench_reference_arm.json", "r02_res_usage.txt", "r02_gputests.txt"):
    if os.path.exists(os.path.join(G, f)): shutil.copy(os.path.join(G, f), os.path.join(P, f))
print(open(os.path.join(P, "r02_split_full.md")).read())
