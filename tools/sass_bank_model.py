#!/usr/bin/env python3
"""Static estimate of FFMA register-bank conflicts in a kernel's SASS (dev tool).
Model (matches the measured loop time of the tiled correlation within 3 %): the register file has
two banks (even / odd register index); an instruction that must fetch two distinct registers of the
same parity -- operands held in the reuse cache excepted -- takes two issue cycles.
usage: sass_bank_model.py <lib.so> <kernel-name-regex>"""
import re
import subprocess
import sys

so, pat = sys.argv[1], sys.argv[2]
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
blocks = re.split(r"\n\s*Function : ", sass)
for blk in blocks[1:]:
    name = blk.split("\n", 1)[0]
    if not re.search(pat, name):
        continue
    prev, tot, cyc, conf, other = {}, 0, 0, 0, 0
    inloop = False
    for l in blk.split("\n"):
        m = re.search(r"/\*[0-9a-f]{4}\*/\s+(@!?U?P\d+\s+)?(\S+)\s*([^;]*);", l)
        if not m:
            continue
        op, args = m.group(2), m.group(3)
        if not op.startswith("FFMA"):
            prev = {}
            continue
        ops = [o.strip() for o in args.split(",")][1:]
        srcs, new = set(), {}
        for slot, o in enumerate(ops):
            mm = re.match(r"-?\|?(R\d+)\|?(\.reuse)?", o)
            if not mm:
                continue
            if mm.group(2):
                new[slot] = mm.group(1)
            if prev.get(slot) == mm.group(1):
                continue
            srcs.add(int(mm.group(1)[1:]))
        ev = len([r for r in srcs if r % 2 == 0])
        c = max(ev, len(srcs) - ev, 1)
        tot += 1; cyc += c; conf += c > 1
        prev = new
    print(f"{name[:110]}\n   FFMA {tot}  conflicting {conf} ({100.0*conf/max(tot,1):.0f}%)  est FFMA issue cycles {cyc}  ({cyc/max(tot,1):.2f}/FFMA)")
