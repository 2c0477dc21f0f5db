#!/usr/bin/env python3
"""Dev: tensor-core cost volume at 224x512 B=8 for C = 8..32 (how much of the tile time is independent of C)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qpwcnet_b200 import ops
from tools.level_bench import timeit
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
ops.set_corr_engine("tc")
for C in (8, 16, 24, 32, 64):
    prv = torch.randn((8, 224, 512, C), device="cuda"); nxt = torch.randn((8, 224, 512, C), device="cuda")
    out = torch.empty((8, 224, 512, 81), device="cuda")
    t = timeit(lambda: ops.cost_volume_into(out, prv, nxt, 4), 10, flush)
    by = 4 * (2 * C + 81) * 8 * 224 * 512
    print(f"C={C:3d}: {t*1e6:7.1f} us   {7168/148*1e-6 and t*1e6/(7168/148):5.2f} us/tile   hbm bound {by/6551.4e9*1e6:6.1f} us")
