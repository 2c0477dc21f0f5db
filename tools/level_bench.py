#!/usr/bin/env python3
"""Per-level timing of the hot-path kernels at the BASELINE.json config-2 shapes (dev tool).
Prints, for each pyramid level: plain corr, fused warp->corr, stand-alone warp, and the achieved
fraction of the HBM / FP32 rooflines.  Usage: python tools/level_bench.py [--iters N]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qpwcnet_b200 import ops  # noqa: E402
from qpwcnet_b200.pyramid import levels_for  # noqa: E402

HBM = 6551.4e9
FP32 = 74.45e12


def timeit(fn, iters, flush):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()                       # 256 MB write: evicts L2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2] * 1e-3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--json", default=None)
    ap.add_argument("--bwd", action="store_true")
    ap.add_argument("--up", action="store_true", help="also time the flow-upsampling fusion (SURVEY 8f rank 1)")
    a = ap.parse_args()
    dev = "cuda"
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
    rows = []
    for lv in levels_for(436, 1024):
        B, H, W, C = a.batch, lv.H, lv.W, lv.C
        g = torch.Generator(device=dev).manual_seed(0)
        prv = torch.randn((B, H, W, C), device=dev, generator=g)
        nxt = torch.randn((B, H, W, C), device=dev, generator=g)
        flo = torch.randn((B, H, W, 2), device=dev, generator=g) * 2
        out = torch.empty((B, H, W, 81), device=dev)
        wout = torch.empty_like(nxt)
        px = B * H * W
        t_corr = timeit(lambda: ops.cost_volume_into(out, prv, nxt, 4), a.iters, flush)
        t_fused = timeit(lambda: ops.warp_cost_volume_into(out, prv, nxt, flo, "tfa", 4), a.iters, flush)
        t_warp = timeit(lambda: ops.warp_into(wout, nxt, flo, "tfa"), a.iters, flush)
        if a.bwd:
            g81 = torch.randn((B, H, W, 81), device=dev, generator=g)
            gC = torch.randn((B, H, W, C), device=dev, generator=g)
            cvout = ops.cost_volume(prv, nxt, 4)
            t_cb = timeit(lambda: ops._corr_bwd(prv, nxt, cvout, g81, 4, 0.1), max(3, a.iters // 4), flush)
            t_wb = timeit(lambda: ops._warp_bwd(nxt, flo, gC, 1), max(3, a.iters // 4), flush)
            print(f"{'':>14}  corr_bwd {t_cb*1e6:9.1f} us ({4*81*C*px/t_cb/1e12:5.1f} TF)   warp_bwd {t_wb*1e6:8.1f} us ({4*(3*C+4)*px/t_wb/1e9:6.0f} GB/s)")
        if a.up and H % 2 == 0 and W % 2 == 0:
            fc = torch.randn((B, H // 2, W // 2, 2), device=dev, generator=g)
            with torch.no_grad():
                t_u = timeit(lambda: ops.upsample2x(fc, 2.0), a.iters, flush)
                t_wu = timeit(lambda: ops.warp_up(nxt, fc, "tfa"), a.iters, flush)
                t_fu = timeit(lambda: ops.warp_cost_volume_up(prv, nxt, fc, "tfa", 4), a.iters, flush)
            print(f"{'':>14}  upsample2x {t_u*1e6:6.1f} us + warp {t_warp*1e6:6.1f} us  vs  warp_up {t_wu*1e6:6.1f} us;"
                  f"  fused corr on coarse flow {t_fu*1e6:7.1f} us (on materialised flow {t_fused*1e6:7.1f} us)")
        flops = 2 * 81 * C * px
        by_corr, by_fused, by_warp = 4 * (2 * C + 81) * px, 4 * (2 * C + 2 + 81) * px, 4 * (2 * C + 2) * px
        lb = lambda by: max(by / HBM, flops / FP32)
        row = dict(level=f"{H}x{W}x{C}", corr_us=t_corr * 1e6, fused_us=t_fused * 1e6, warp_us=t_warp * 1e6,
                   corr_lb_us=lb(by_corr) * 1e6, fused_lb_us=lb(by_fused) * 1e6, warp_lb_us=by_warp / HBM * 1e6,
                   corr_frac=lb(by_corr) / t_corr, fused_frac=lb(by_fused) / t_fused,
                   warp_frac=by_warp / HBM / t_warp, corr_tflops=flops / t_corr / 1e12,
                   fused_tflops=flops / t_fused / 1e12)
        rows.append(row)
        print("{level:>14}  corr {corr_us:8.1f} us ({corr_frac:5.1%} of lb {corr_lb_us:6.1f}, {corr_tflops:5.1f} TF)  "
              "fused {fused_us:8.1f} us ({fused_frac:5.1%} of lb {fused_lb_us:6.1f}, {fused_tflops:5.1f} TF)  "
              "warp {warp_us:7.1f} us ({warp_frac:5.1%})".format(**row), flush=True)
    tot_f = sum(r["fused_us"] if i else r["corr_us"] for i, r in enumerate(rows))
    tot_u = sum((r["corr_us"] + r["warp_us"]) if i else r["corr_us"] for i, r in enumerate(rows))
    print(f"pyramid: fused path {tot_f:.1f} us -> {a.batch / tot_f * 1e6:.0f} pairs/s; unfused path {tot_u:.1f} us -> {a.batch / tot_u * 1e6:.0f} pairs/s")
    if a.json:
        json.dump(rows, open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()
