#!/usr/bin/env python3
"""Dev: run the tensor-core cost volume once at the finest level (for ncu)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qpwcnet_b200 import ops
B, H, W, C = 8, 224, 512, int(sys.argv[1]) if len(sys.argv) > 1 else 32
prv = torch.randn((B, H, W, C), device="cuda"); nxt = torch.randn((B, H, W, C), device="cuda")
out = torch.empty((B, H, W, 81), device="cuda")
ops.set_corr_engine("tc")
for _ in range(3):
    ops.cost_volume_into(out, prv, nxt, 4)
torch.cuda.synchronize()
