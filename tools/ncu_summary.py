#!/usr/bin/env python3
"""Summarise an ncu run for profiles/: launch list (gpu__time_duration per launch) + key raw metrics + SASS opcode mix.
usage: tools/ncu_summary.py <launches.csv> <prof.ncu-rep> <out.md> [title]"""
import collections, csv, subprocess, sys

launch_csv, rep, out, title = sys.argv[1], sys.argv[2], sys.argv[3], (sys.argv[4] if len(sys.argv) > 4 else "ncu summary")
lines = [f"# {title}", ""]
rows = [r for r in csv.reader(open(launch_csv)) if len(r) > 5]
hdr = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
H, data = rows[hdr], rows[hdr + 1:]
ki, vi, gi, bi = H.index("Kernel Name"), H.index("Metric Value"), H.index("Grid Size"), H.index("Block Size")
d = collections.defaultdict(list)
for r in data:
    d[(r[ki].split("(")[0], r[gi], r[bi])].append(float(r[vi].replace(",", "")))
tot = sum(sum(v) for v in d.values())
lines += ["## Launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`; cold-cache, serialised: compare shares)", "",
          "| kernel | grid | block | launches | avg us | min us | max us | share of GPU time |", "|---|---|---|---|---|---|---|---|"]
for (k, g, b), v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
    lines.append(f"| `{k}` | {g} | {b} | {len(v)} | {sum(v)/len(v)/1e3:.2f} | {min(v)/1e3:.2f} | {max(v)/1e3:.2f} | {100*sum(v)/tot:.1f}% |")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
RH = rr[0]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "lts__t_sector_hit_rate.pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_warps", "launch__waves_per_multiprocessor", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_no_instruction_per_warp_active.pct",
        "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_wait_per_warp_active.pct",
        "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct"]
lines += ["", "## `ncu --set full` raw metrics (per captured launch)", "", "| metric | unit | " + " | ".join(f"launch {i}" for i in range(len(rr) - 2)) + " |",
          "|---|---|" + "---|" * (len(rr) - 2)]
for w in want:
    if w in RH:
        i = RH.index(w)
        lines.append(f"| {w} | {rr[1][i]} | " + " | ".join(r[i] for r in rr[2:]) + " |")
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
sr = list(csv.reader(sass.splitlines()))
SH = sr[1]
ie, si = SH.index("Instructions Executed"), SH.index("Source")
blk = []
for r in sr[2:]:
    if len(r) < len(SH):
        if blk:
            break
        continue
    blk.append(r)
totI = sum(float(r[ie]) for r in blk if r[ie])
h = collections.Counter()
for r in blk:
    if r[ie]:
        t = r[si].split()
        h[(t[1] if t[0].startswith("@") else t[0]).split(".")[0]] += float(r[ie])
lines += ["", f"## SASS opcode mix of the first captured launch ({len(blk)} SASS instructions, {int(totI)} warp-instructions executed)", "",
          "| opcode | warp-instructions | share |", "|---|---|---|"]
for k, v in h.most_common(18):
    lines.append(f"| {k} | {int(v)} | {100*v/totI:.1f}% |")
open(out, "w").write("\n".join(lines) + "\n")
print("wrote", out)
