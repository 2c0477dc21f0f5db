#!/usr/bin/env python3
"""Dev: per-kernel device time of the config-3 training-step hot path (torch profiler)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qpwcnet_b200.train_step import TrainHotPath
per = int(sys.argv[1]) if len(sys.argv) > 1 else 64
th = TrainHotPath(per, "cuda")
for _ in range(3):
    th.step(); th.zero_grad()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(5):
        th.step(); th.zero_grad()
    torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
tot = sum(e.device_time_total for e in rows)
print(f"total device time per step: {tot/5/1e3:.3f} ms")
for e in rows[:18]:
    print(f"{e.device_time_total/5/1e3:8.3f} ms {100*e.device_time_total/tot:5.1f}%  x{e.count//5:3d}  {e.key[:110]}")
