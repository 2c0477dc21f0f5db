#!/usr/bin/env python3
"""Kernel-time breakdown of the config-3 training-step hot path on one GPU (dev tool): 10 cost volumes
+ 8 fused warps + 5 half-flow warp pairs, forward and backward, per-GPU batch `per` (default 8)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qpwcnet_b200 import ops
per = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = "cuda"
levels = [(8, 14, 256), (16, 28, 256), (32, 56, 128), (64, 112, 64), (128, 224, 32)]
mk = lambda *s: torch.randn(s, device=dev).requires_grad_()
ten = [(mk(per, H, W, C), mk(per, H, W, C), mk(per, H, W, 2)) for (H, W, C) in levels]
img = [(mk(per, H, W, c), mk(per, H, W, c), mk(per, H, W, 2), mk(per, H, W, 2)) for (H, W, _), c in zip(levels, (3, 256, 128, 64, 32))]
def step():
    losses = []
    for _pass in range(2):
        for k, (p, n, f) in enumerate(ten):
            cv = ops.cost_volume(p, n, 4) if k == 0 else ops.warp_cost_volume(p, n, f, "tfa", 4)
            losses.append(cv.mean())
    for (pa, nb_, f01, f10) in img:
        losses.append(ops.half_flow_warps(pa, nb_, f01, f10, "tfa").mean())
    torch.stack(losses).sum().backward()
for _ in range(3): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): step()
e1.record(); torch.cuda.synchronize()
print(f"per-GPU batch {per}: {e0.elapsed_time(e1)/5:.2f} ms per step (device time)")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=70))
