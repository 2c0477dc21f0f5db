#!/usr/bin/env python3
"""BASELINE.json config 4: fused warp->correlation sweep, C in {16,32,64,96,128,196} at the canonical
PWC-Net level sizes of 436x1024 (ceil-halving), B=8, d in {4,8}; parity spot-check against the oracle
on a crop + achieved fraction of the roofline (max of HBM and FP32 bound).  Measurement tool."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402  (checker only)
from qpwcnet_b200 import ops  # noqa: E402

HBM, FP32 = 6551.4e9, 74.45e12
LEVELS = [(218, 512, 16), (109, 256, 32), (55, 128, 64), (28, 64, 96), (14, 32, 128), (7, 16, 196)]
B = 8
dev = "cuda"
rows = []
for d in (4, 8):
    D = (2 * d + 1) ** 2
    for (H, W, C) in LEVELS:
        g = torch.Generator(device=dev).manual_seed(0)
        prv = torch.randn((B, H, W, C), device=dev, generator=g)
        nxt = torch.randn((B, H, W, C), device=dev, generator=g)
        flo = torch.randn((B, H, W, 2), device=dev, generator=g) * (d / 2)
        out = torch.empty((B, H, W, D), device=dev)
        scratch = torch.empty_like(nxt)
        res = {}
        for how in ("fused", "composed"):
            def run():
                if how == "fused":
                    ops.warp_cost_volume_into(out, prv, nxt, flo, "tfa", d)
                else:
                    ops.warp_into(scratch, nxt, flo, "tfa")
                    ops.cost_volume_into(out, prv, scratch, d)
            for _ in range(2):
                run()
            torch.cuda.synchronize()
            # n launches recorded into one CUDA graph: device time, not the host's launch latency (the coarse
            # levels take less time on the GPU than one Python -> C-ABI call takes on the host)
            n = 5 if d == 8 else 10
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.stream(side):
                with torch.cuda.graph(graph, stream=side):
                    for _ in range(n):
                        run()
            torch.cuda.current_stream().wait_stream(side)
            graph.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            graph.replay()
            e1.record()
            torch.cuda.synchronize()
            res[how] = e0.elapsed_time(e1) / n * 1e-3
        # parity spot check (fused) on a crop of batch item 0 against the fp64 oracle
        ops.warp_cost_volume_into(out, prv, nxt, flo, "tfa", d)
        hh, ww = min(H, 24), min(W, 32)
        pad = 2 * d + 8
        h1, w1 = min(H, hh + pad), min(W, ww + pad)
        # the crop's top-left corner coincides with the image's: border rule identical there
        ref = oracle.warp_cost_volume(prv[:1, :h1, :w1].cpu().numpy().astype(np.float64),
                                      nxt[:1, :h1, :w1].cpu().numpy().astype(np.float64),
                                      flo[:1, :h1, :w1].cpu().numpy().astype(np.float64), "tfa", d)
        safe_h, safe_w = max(1, h1 - pad), max(1, w1 - pad)
        if h1 == H and w1 == W:
            safe_h, safe_w = H, W
        got = out[:1, :safe_h, :safe_w].cpu().numpy()
        err = float(np.abs(got - ref[:, :safe_h, :safe_w]).max() / np.abs(ref).max())
        px = B * H * W
        by, fl = 4 * (2 * C + 2 + D) * px, 2 * D * C * px
        lb = max(by / HBM, fl / FP32)
        best = min(res.values())
        rows.append(dict(d=d, level=f"{H}x{W}x{C}", fused_us=res["fused"] * 1e6, composed_us=res["composed"] * 1e6,
                         lower_bound_us=lb * 1e6, bound="HBM" if by / HBM > fl / FP32 else "FP32",
                         frac_of_roofline=lb / best, rel_err_vs_oracle=err))
        print(rows[-1], flush=True)
json.dump(rows, open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/cfg4_sweep.json", "w"), indent=1)
