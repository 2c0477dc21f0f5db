#!/usr/bin/env python3
"""Dev tool: time the tiled correlation kernel with parts switched off (QPWC_ABLATE bits)."""
import os, subprocess, sys
lvl = sys.argv[1] if len(sys.argv) > 1 else "4"
op = sys.argv[2] if len(sys.argv) > 2 else "corr"
code = r'''
import os, sys, torch
sys.path.insert(0, ".")
from qpwcnet_b200 import ops
from qpwcnet_b200.pyramid import levels_for
lv = levels_for(436, 1024)[int(sys.argv[1])]
B, H, W, C = 8, lv.H, lv.W, lv.C
g = torch.Generator(device="cuda").manual_seed(0)
prv = torch.randn((B, H, W, C), device="cuda", generator=g); nxt = torch.randn((B, H, W, C), device="cuda", generator=g)
flo = torch.randn((B, H, W, 2), device="cuda", generator=g) * 2
out = torch.empty((B, H, W, 81), device="cuda")
f = (lambda: ops.cost_volume_into(out, prv, nxt, 4)) if sys.argv[2] == "corr" else (lambda: ops.warp_cost_volume_into(out, prv, nxt, flo, "tfa", 4))
for _ in range(3): f()
torch.cuda.synchronize()
ts = []
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); f(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
ts.sort(); print(f"{ts[len(ts)//2]*1e3:8.1f} us")
'''
names = {0: "full", 1: "no FFMA loop", 2: "no epilogue", 3: "no FFMA, no epilogue (loads only)", 4: "no loads", 5: "no loads, no FFMA (epilogue only)", 6: "no loads, no epilogue (FFMA only)", 7: "nothing"}
for a in range(8):
    env = dict(os.environ, QPWC_ABLATE=str(a))
    r = subprocess.run([sys.executable, "-c", code, lvl, op], env=env, capture_output=True, text=True)
    print(f"ablate={a} {names[a]:38s} {r.stdout.strip()} {r.stderr.strip()[-200:]}")
