#!/usr/bin/env python3
"""channels_first cost volume at the config-2 level shapes (dev tool): native NCHW kernel vs the
NHWC kernel on pre-transposed data vs the full transposing route."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qpwcnet_b200 import ops
from qpwcnet_b200.pyramid import levels_for
from tools.level_bench import timeit
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
for lv in levels_for(436, 1024):
    B, H, W, C = 8, lv.H, lv.W, lv.C
    g = torch.Generator(device="cuda").manual_seed(0)
    p = torch.randn((B, C, H, W), device="cuda", generator=g); n = torch.randn((B, C, H, W), device="cuda", generator=g)
    ph, nh = p.permute(0, 2, 3, 1).contiguous(), n.permute(0, 2, 3, 1).contiguous()
    out = ops._corr_fwd_nchw(p, n, 4, 0.1); ref = ops.cost_volume(ph, nh, 4)
    err = float((out.permute(0, 2, 3, 1) - ref).abs().max() / ref.abs().max())
    t_n = timeit(lambda: ops._corr_fwd_nchw(p, n, 4, 0.1), 12, flush)
    t_h = timeit(lambda: ops.cost_volume(ph, nh, 4), 12, flush)
    t_t = timeit(lambda: ops.cost_volume(p.permute(0, 2, 3, 1).contiguous(), n.permute(0, 2, 3, 1).contiguous(), 4).permute(0, 3, 1, 2).contiguous(), 12, flush)
    print(f"{C}x{H}x{W} B={B}: native NCHW {t_n*1e6:7.1f} us | NHWC kernel {t_h*1e6:7.1f} us | transposing route {t_t*1e6:7.1f} us | rel diff {err:.1e}", flush=True)
