#!/usr/bin/env python3
"""A/B of the plain cost-volume kernels at the config-2 level shapes (dev tool): QPWC_CORR_VARIANT =
tiled (channel-parity FFMA2, 4-row tiles) vs rowpair (row-pair FFMA2, 8-row tiles).  Checks that the
two agree and prints median device times with L2 flushed between iterations."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qpwcnet_b200 import ops  # noqa: E402
from qpwcnet_b200.pyramid import levels_for  # noqa: E402
from tools.level_bench import timeit  # noqa: E402


def main():
    dev = "cuda"
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
    shapes = [(8, lv.H, lv.W, lv.C) for lv in levels_for(436, 1024)] + [(8, 109, 256, 64), (8, 218, 512, 16), (3, 61, 190, 36)]
    for B, H, W, C in shapes:
        g = torch.Generator(device=dev).manual_seed(0)
        prv = torch.randn((B, H, W, C), device=dev, generator=g)
        nxt = torch.randn((B, H, W, C), device=dev, generator=g)
        res = {}
        for var in ("packed", "rowpair", "two", "default"):
            os.environ["QPWC_CORR_VARIANT"] = var
            out = torch.full((B, H, W, 81), float("nan"), device=dev)
            ops.cost_volume_into(out, prv, nxt, 4)
            torch.cuda.synchronize()
            t = timeit(lambda: ops.cost_volume_into(out, prv, nxt, 4), 15, flush)
            res[var] = (out.clone(), t)
        a, b = res["default"][0], res["rowpair"][0]
        err = (a - b).abs().max().item() / a.abs().max().item()
        nan = int(torch.isnan(b).sum().item())
        err2 = (a - res["two"][0]).abs().max().item() / a.abs().max().item()
        print(f"{H}x{W}x{C} B={B}: two-CTA {res['two'][1]*1e6:8.1f} us (diff {err2:.1e})  packed {res['packed'][1]*1e6:8.1f} us  rowpair {res['rowpair'][1]*1e6:8.1f} us  default(scalar) {res['default'][1]*1e6:8.1f} us  "
              f"rel diff {err:.2e}  nan {nan}", flush=True)


if __name__ == "__main__":
    main()
