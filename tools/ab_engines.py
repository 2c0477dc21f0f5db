#!/usr/bin/env python3
"""A/B of the two cost-volume forward engines (dev tool): tensor cores (3xTF32 split,
qpwc_corr_tc.cu) vs FFMA (qpwc_corr_tiled.cu) at the config-2 level shapes and a few ragged ones.
Prints median device times (L2 flushed between iterations), the difference between the engines and,
for small shapes, the error of each against the fp64 CPU oracle relative to mean|prv*nxt|."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qpwcnet_b200 import ops  # noqa: E402
from qpwcnet_b200.pyramid import levels_for  # noqa: E402
from tools.level_bench import timeit  # noqa: E402


def main():
    dev = "cuda"
    quick = "--quick" in sys.argv
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
    shapes = [(1, 8, 16, 8), (1, 16, 32, 32), (2, 13, 37, 24), (3, 61, 190, 40)]
    if not quick:
        shapes += [(8, lv.H, lv.W, lv.C) for lv in levels_for(436, 1024)] + [(8, 109, 256, 64), (8, 218, 512, 16)]
    for B, H, W, C in shapes:
        g = torch.Generator(device=dev).manual_seed(0)
        prv = torch.randn((B, H, W, C), device=dev, generator=g)
        nxt = torch.randn((B, H, W, C), device=dev, generator=g)
        res = {}
        for eng in ("ffma", "tc"):
            ops.set_corr_engine(eng)
            out = torch.full((B, H, W, 81), float("nan"), device=dev)
            ops.cost_volume_into(out, prv, nxt, 4)
            torch.cuda.synchronize()
            t = timeit(lambda: ops.cost_volume_into(out, prv, nxt, 4), 15, flush)
            res[eng] = (out.clone(), t)
        ops.set_corr_engine("auto")
        a, b = res["ffma"][0], res["tc"][0]
        nan = int(torch.isnan(b).sum().item())
        diff = (a - b).abs().max().item() / a.abs().max().item()
        msg = f"{H}x{W}x{C} B={B}: ffma {res['ffma'][1]*1e6:8.1f} us  tc {res['tc'][1]*1e6:8.1f} us  max|ffma-tc|/max {diff:.2e}  nan {nan}"
        if B * H * W * C <= 2_000_000:
            import oracle
            p64, n64 = prv.cpu().numpy().astype(np.float64), nxt.cpu().numpy().astype(np.float64)
            ref = oracle.cost_volume(p64, n64, 4)
            pad = np.zeros((B, H + 8, W + 8, C))
            pad[:, 4:4 + H, 4:4 + W] = np.abs(n64)
            cond = np.empty_like(ref)
            for i0 in range(9):
                for j0 in range(9):
                    cond[..., i0 * 9 + j0] = (np.abs(p64) * pad[:, i0:i0 + H, j0:j0 + W]).mean(-1)
            for eng in ("ffma", "tc"):
                err = np.abs(res[eng][0].cpu().numpy().astype(np.float64) - ref)
                msg += f"  {eng}: max|err| {err.max():.2e} worst err/cond {np.max(err / (cond + 1e-30)):.2e}"
        print(msg, flush=True)


if __name__ == "__main__":
    main()
