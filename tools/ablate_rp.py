#!/usr/bin/env python3
"""Ablation timing of the row-pair cost-volume kernel at the finest config-2 level (dev tool)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qpwcnet_b200 import ops
from tools.level_bench import timeit
os.environ["QPWC_CORR_VARIANT"] = "rowpair"
B, H, W, C = 8, 224, 512, int(sys.argv[1]) if len(sys.argv) > 1 else 32
g = torch.Generator(device="cuda").manual_seed(0)
prv = torch.randn((B, H, W, C), device="cuda", generator=g); nxt = torch.randn((B, H, W, C), device="cuda", generator=g)
out = torch.empty((B, H, W, 81), device="cuda")
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
for ab, name in [(0, "full"), (1, "no stores"), (2, "no repack"), (4, "no loads"), (8, "no epilogue"), (9, "no epilogue/stores"), (12, "no loads, no epilogue"), (14, "main loop only (no loads/repack/epilogue)")]:
    os.environ["QPWC_ABLATE_RP"] = str(ab)
    t = timeit(lambda: ops.cost_volume_into(out, prv, nxt, 4), 10, flush)
    print(f"ablate {ab:2d} {name:45s} {t*1e6:8.1f} us", flush=True)
