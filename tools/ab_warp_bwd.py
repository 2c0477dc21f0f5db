#!/usr/bin/env python3
"""Dev: warp backward, direct global atomics vs shared-memory pre-aggregated tiles (QPWC_OPT_WARP_BWD)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qpwcnet_b200 import ops, _cabi
from qpwcnet_b200.pyramid import levels_for
from tools.level_bench import timeit
L = _cabi.lib()
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
for sigma, label in ((2.0, "noise N(0,2^2)"), (0.0, "smooth")):
    print(label)
    for lv in levels_for(436, 1024)[1:]:
        B, H, W, C = 8, lv.H, lv.W, lv.C
        g = torch.Generator(device="cuda").manual_seed(0)
        img = torch.randn((B, H, W, C), device="cuda", generator=g)
        if sigma:
            flo = torch.randn((B, H, W, 2), device="cuda", generator=g) * sigma
        else:
            yy, xx = torch.meshgrid(torch.arange(H, device="cuda"), torch.arange(W, device="cuda"), indexing="ij")
            flo = torch.stack([3 * torch.sin(yy / 17.0) + 0.3, 2 * torch.cos(xx / 23.0) - 0.4], -1)[None].repeat(B, 1, 1, 1).contiguous().float()
        gw = torch.randn((B, H, W, C), device="cuda", generator=g)
        res = {}
        for v, name in ((1, "direct"), (2, "tiles")):
            L.qpwc_set_option(1, v)
            gi, gf = ops._warp_bwd(img, flo, gw, 1)
            torch.cuda.synchronize()
            res[name] = (gi, gf, timeit(lambda: ops._warp_bwd(img, flo, gw, 1), 10, flush))
        L.qpwc_set_option(1, 0)
        di = (res["direct"][0] - res["tiles"][0]).abs().max().item()
        df = (res["direct"][1] - res["tiles"][1]).abs().max().item()
        by = 4 * (3 * C + 4) * B * H * W
        print(f"  {H}x{W}x{C}: direct {res['direct'][2]*1e6:7.1f} us  tiles {res['tiles'][2]*1e6:7.1f} us  (bound {by/6551.4e9*1e6:5.1f} us)  max|dg_img| {di:.2e} max|dg_flow| {df:.2e}")
