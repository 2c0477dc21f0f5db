#!/usr/bin/env python3
"""Run one hot-path op a few times at one pyramid level (target for ncu).  Dev tool.
   python tools/prof_one.py --op corr|fused|warp|corr_bwd|warp_bwd --level 0..4 [--iters 3]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qpwcnet_b200 import ops  # noqa: E402
from qpwcnet_b200.pyramid import levels_for  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--op", default="corr")
ap.add_argument("--level", type=int, default=4)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--batch", type=int, default=8)
a = ap.parse_args()
lv = levels_for(436, 1024)[a.level]
B, H, W, C = a.batch, lv.H, lv.W, lv.C
g = torch.Generator(device="cuda").manual_seed(0)
prv = torch.randn((B, H, W, C), device="cuda", generator=g)
nxt = torch.randn((B, H, W, C), device="cuda", generator=g)
flo = torch.randn((B, H, W, 2), device="cuda", generator=g) * 2
out = torch.empty((B, H, W, 81), device="cuda")
for _ in range(a.iters):
    if a.op == "corr":
        ops.cost_volume_into(out, prv, nxt, 4)
    elif a.op == "fused":
        ops.warp_cost_volume_into(out, prv, nxt, flo, "tfa", 4)
    elif a.op == "warp":
        ops.warp(nxt, flo, "tfa")
    elif a.op == "corr_bwd":
        p, n = prv.clone().requires_grad_(), nxt.clone().requires_grad_()
        ops.cost_volume(p, n, 4).backward(torch.ones_like(out))
    elif a.op == "warp_bwd":
        n, f = nxt.clone().requires_grad_(), flo.clone().requires_grad_()
        ops.warp(n, f, "tfa").backward(torch.ones_like(nxt))
torch.cuda.synchronize()
print("done", a.op, (B, H, W, C))
