"""Turns ncu outputs under gpurun_out/ into the committed summaries under profiles/.
usage: python scripts/ncu_summary.py <round-tag> <launches.csv> <full.ncu-rep> [kernel-regex]"""
import csv
import json
import os
import re
import subprocess
import sys
from collections import defaultdict

tag, launches, rep = sys.argv[1], sys.argv[2], sys.argv[3]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)

# ---- launch list: per-kernel totals and share of the step --------------------------------
rows = [r for r in csv.reader(open(launches)) if len(r) > 10]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
d = defaultdict(list)
for r in rows[1:]:
    try:
        d[re.sub(r"\(.*", "", r[ki])[:70]].append(float(r[vi].replace(",", "")))
    except ValueError:
        pass
tot = sum(sum(v) for v in d.values())
lines = [f"# ncu launch list ({tag}): `ncu --metrics gpu__time_duration.sum --clock-control none` around `python bench.py --steps 5 --warmup 3 --no-cpu --no-extras` (burst probe 3 + 20, warm-up 3, timed 5 config-3 launches; then the chunk-pipelined e2e leg: 16 (b,h)-chunk launches per host step)",
         "# per-launch times are cold-cache and serialised: compare SHARES, not absolutes", "",
         f"{'kernel':70s} {'launches':>8s} {'total_us':>10s} {'avg_us':>9s} {'share':>6s}"]
for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
    lines.append(f"{k:70s} {len(v):8d} {sum(v) / 1e3:10.1f} {sum(v) / len(v) / 1e3:9.1f} {sum(v) / tot:6.3f}")
open(os.path.join(out_dir, f"launches_{tag}.txt"), "w").write("\n".join(lines) + "\n")
print("\n".join(lines))

# ---- full capture: the metrics the north-star asks for -----------------------------------
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h, u, vals = rr[0], rr[1], rr[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__cycles_active.avg", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum"]
m = {}
for i, name in enumerate(h):
    if name in want:
        m[name] = (vals[i], u[i])
kname = vals[h.index("Kernel Name")] if "Kernel Name" in h else "?"
lines = [f"# ncu --set full summary ({tag}) of `{kname[:90]}`", f"# report: {os.path.basename(rep)} (kept under gpurun_out/, not committed)", ""]
for name in want:
    if name in m:
        lines.append(f"{name:75s} {m[name][0]:>16s} {m[name][1]}")
open(os.path.join(out_dir, f"ncu_full_{tag}.txt"), "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
def num(name):
    return float(m[name][0].replace(",", ""))
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
traffic = num("dram__bytes_read.sum") * scale[m["dram__bytes_read.sum"][1]] + num("dram__bytes_write.sum") * scale[m["dram__bytes_write.sum"][1]]
json.dump({"kernel": "fa_fwd_kernel<128>", "dram_bytes_per_launch": traffic, "source": f"profiles/ncu_full_{tag}.txt",
           "algorithmic_bytes_per_launch": 4 * 8 * 32 * 4096 * 128 * 2}, open(os.path.join(out_dir, "fa_fwd_traffic.json"), "w"))
print("dram traffic per launch", traffic)
