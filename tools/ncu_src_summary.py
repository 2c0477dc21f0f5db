#!/usr/bin/env python3
"""Summarise an ncu source-page CSV (ncu -i X.ncu-rep --page source --csv): top stalled SASS
instructions and per-opcode totals.  Dev tool."""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
tot = sum(int(r[ix["# Samples"]]) for r in data)
exe = sum(int(r[ix["Instructions Executed"]]) for r in data)
print(f"instructions: {len(data)} SASS lines, {exe} warp-instr executed, {tot} samples")
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = Counter()
for r in data:
    for s in stalls:
        agg[s] += int(r[ix[s]] or 0)
print("stall totals:", ", ".join(f"{k[6:]}={v}" for k, v in agg.most_common(10)))
byop, byop_exec = Counter(), Counter()
for r in data:
    op = r[ix["Source"]].split()[0] if not r[ix["Source"]].strip().startswith("@") else r[ix["Source"]].split()[1]
    op = op.split(".")[0]
    byop[op] += int(r[ix["# Samples"]])
    byop_exec[op] += int(r[ix["Instructions Executed"]])
print("by opcode (samples, executed):")
for op, v in byop.most_common(14):
    print(f"  {op:12s} {v:7d} {100*v/tot:5.1f}%   exec {byop_exec[op]:10d} {100*byop_exec[op]/exe:5.1f}%")
print("top instructions:")
order = sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]]))[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]
for i in sorted(order):
    r = data[i]
    top = sorted(((int(r[ix[s]] or 0), s[6:]) for s in stalls), reverse=True)[:3]
    print(f"  [{i:5d}] {r[ix['Source']].strip():70s} samp {r[ix['# Samples']]:>5s} exec {r[ix['Instructions Executed']]:>8s}  "
          + " ".join(f"{n}:{v}" for v, n in top if v))
# shared-memory conflict summary
exc = sum(int(r[ix["L1 Wavefronts Shared Excessive"]] or 0) for r in data)
wf = sum(int(r[ix["L1 Wavefronts Shared"]] or 0) for r in data)
print(f"shared wavefronts {wf}, excessive {exc}")
for i, r in enumerate(data):
    e = int(r[ix["L1 Wavefronts Shared Excessive"]] or 0)
    if e > exc * 0.05 and exc:
        print(f"  conflict [{i}] {r[ix['Source']].strip():60s} wf {r[ix['L1 Wavefronts Shared']]} ideal {r[ix['L1 Wavefronts Shared Ideal']]}")
