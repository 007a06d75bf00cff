"""Turn gpurun_out/launches_<tag>.csv and prof_<tag>_*.ncu-rep into tracked summaries under profiles/."""
import collections
import csv
import os
import re
import subprocess
import sys

tag = sys.argv[1]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)
out = [f"# ncu summary {tag}", "",
       "Command: `ncu --metrics gpu__time_duration.sum --clock-control none --csv python bench.py --minutes <M> --steps 1 --warmup 1 --no-cpu-baseline`",
       "(per-launch times are cold-cache and serialised: compare SHARES with bench.py's `kernels_ms_per_step`, not absolutes)", ""]
src = os.path.join(G, f"launches_{tag}.csv")
if os.path.exists(src):
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"<.*", "", re.sub(r"\(.*", "", row["Kernel Name"].replace("<unnamed>::", ""))).replace("void ", "")
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        ms = v / 1e6 if u.startswith("ns") else v / 1e3 if u.startswith("us") else v
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ms
    tot = sum(a[1] for a in agg.values())
    out += ["| kernel | launches | total ms | share |", "|---|---:|---:|---:|"]
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| {k} | {n} | {ms:.3f} | {100 * ms / tot:.1f}% |")
    out += ["", f"total kernel time under ncu: {tot:.1f} ms over {sum(a[0] for a in agg.values())} launches", ""]
    subprocess.call(["cp", src, os.path.join(P, f"launches_{tag}.csv")])
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "smsp__inst_executed.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]
for f in sorted(os.listdir(G)):
    if f.startswith(f"prof_{tag}_") and f.endswith(".ncu-rep"):
        raw = subprocess.run(["ncu", "-i", os.path.join(G, f), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        r = list(csv.reader(raw.splitlines()))
        if len(r) < 3:
            continue
        hdr, units, vals = r[0], r[1], r[2]
        out += [f"## `ncu --set full` : {f[len('prof_' + tag + '_'):-8]} (one launch)", "", "| metric | value |", "|---|---:|"]
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                out.append(f"| {w} | {vals[i]} {units[i]} |")
        out.append("")
open(os.path.join(P, f"ncu_{tag}.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out[:40]))
