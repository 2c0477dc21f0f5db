#!/usr/bin/env python3
"""Dev: e2e step time of the config-2 pyramid through the host-buffer entry points (pinned tensors)."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qpwcnet_b200.pyramid import PyramidWorkload
wl = PyramidWorkload(436, 1024, 8, 4, device="cpu", seed=0)
for _ in range(3):
    wl.step()
torch.cuda.synchronize()
ts = []
for _ in range(10):
    t0 = time.perf_counter()
    wl.step()
    _ = float(wl.outputs[0][0, 0, 0, 0])
    ts.append(time.perf_counter() - t0)
print(" ".join(f"{t*1e3:.1f}" for t in ts))
ts.sort()
print(f"slice {os.environ.get('QPWC_HOST_SLICE_MB', '8')} MiB: median {ts[5]*1e3:.2f} ms  min {ts[0]*1e3:.2f} ms  -> {8/ts[5]:.0f} pairs/s")
