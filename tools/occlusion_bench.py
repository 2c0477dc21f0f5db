"""Times estimate_occlusion_map (three launches, the flow self-warp kept in registers) against the
composition the reference's graph performs (tf_warp kernel + index arithmetic + scatter + max, here
with our warp kernel and torch ops) at Sintel size.  GPU only."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from qpwcnet_b200 import ops


def composed(flow):
    B, H, W, _ = flow.shape
    i = torch.arange(H, device=flow.device, dtype=torch.float32)[None, :, None]
    j = torch.arange(W, device=flow.device, dtype=torch.float32)[None, None, :]
    i2, j2 = i + flow[..., 1], j + flow[..., 0]
    oob = ((i2 < 0) | (i2 >= H) | (j2 < 0) | (j2 >= W)).float()
    inv = -ops.warp(flow, flow, "tf")
    i3 = (i + inv[..., 1]).to(torch.int32).clamp(0, H - 1).long()
    j3 = (j + inv[..., 0]).to(torch.int32).clamp(0, W - 1).long()
    b = torch.arange(B, device=flow.device)[:, None, None].expand(B, H, W)
    map3 = torch.ones((B, H, W), device=flow.device)
    map3[b, i3, j3] = 0.0
    return torch.maximum(oob, map3)


def timed(fn, n=50):
    for _ in range(5):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def main():
    B, H, W = 8, 436, 1024
    g = torch.Generator(device="cuda").manual_seed(0)
    flow = torch.randn((B, H, W, 2), device="cuda", generator=g) * 6.0
    a, b = ops.occlusion_map(flow), composed(flow)
    res = {"shape": [B, H, W], "equal_to_composition": bool(torch.equal(a, b)),
           "occluded_fraction": float(a.mean()),
           "fused_us": timed(lambda: ops.occlusion_map(flow)), "composed_us": timed(lambda: composed(flow)),
           "algorithmic_bytes": B * H * W * (8 + 4)}
    res["fused_GBps"] = res["algorithmic_bytes"] / res["fused_us"] / 1e3
    json.dump(res, sys.stdout)
    print()


if __name__ == "__main__":
    main()
