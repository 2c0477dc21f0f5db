#!/usr/bin/env python3
"""QPWC_ABLATE timing of the fused warp->correlation kernel at the finest config-2 level (dev tool)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qpwcnet_b200 import ops
from tools.level_bench import timeit
B, H, W, C = 8, 224, 512, 32
g = torch.Generator(device="cuda").manual_seed(0)
prv = torch.randn((B, H, W, C), device="cuda", generator=g); nxt = torch.randn((B, H, W, C), device="cuda", generator=g)
flo = torch.randn((B, H, W, 2), device="cuda", generator=g) * 2
out = torch.empty((B, H, W, 81), device="cuda"); flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
t = timeit(lambda: ops.warp_cost_volume_into(out, prv, nxt, flo, "tfa", 4), 12, flush)
print(f"QPWC_ABLATE={os.environ.get('QPWC_ABLATE', '0')}: fused {t*1e6:.1f} us")
