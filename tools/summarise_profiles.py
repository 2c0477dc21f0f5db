#!/usr/bin/env python3
"""Turn ncu captures in gpurun_out/ into the small tracked summaries under profiles/ (dev tool).
usage: summarise_profiles.py <round-tag>   e.g. r01"""
import csv
import json
import os
import subprocess
import sys
from collections import defaultdict

tag = sys.argv[1]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GO, PR = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(PR, exist_ok=True)
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sectors_srcunit_tex.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg",
        "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
out = {}
for f in sorted(os.listdir(GO)):
    if not (f.endswith(".ncu-rep") and tag in f):
        continue
    raw = subprocess.run(["ncu", "-i", os.path.join(GO, f), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) < 3:
        continue
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {"kernel": vals[hdr.index("Kernel Name")][:120]}
    for h, u, v in zip(hdr, units, vals):
        if h in KEYS:
            d[h] = f"{v} {u}".strip()
    out[f] = d
json.dump(out, open(os.path.join(PR, f"{tag}_ncu_full_summary.json"), "w"), indent=1)

# launch list: per-kernel totals and shares
ll = os.path.join(GO, f"launches_{tag}.csv")
if os.path.exists(ll):
    rows = [r for r in csv.reader(open(ll)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        name = r[ki].split("(")[0][:90]
        agg[name][0] += 1
        agg[name][1] += float(r[vi].replace(",", ""))
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(PR, f"{tag}_launch_list_summary.txt"), "w") as fo:
        fo.write(f"# ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none  python bench.py --steps 5 --warmup 3 --no-graph --no-cpu-baseline --path composed  (the timed steps only; the unprofiled bench picks the composed path at every UpFlow level)\n")
        fo.write(f"# per-launch times are cold-cache and serialised: compare SHARES\n")
        for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            fo.write(f"{t/1e3:10.1f} us  {100*t/tot:5.1f}%  launches {n:4d}  avg {t/n/1e3:8.1f} us  {name}\n")
    import shutil
    shutil.copy(ll, os.path.join(PR, f"{tag}_launch_list.csv"))
print(json.dumps(out, indent=1)[:3000])
