"""channels_first warp, forward + backward: native NCHW kernels against the transposing route
(permute -> NHWC kernel -> permute) at the finest pyramid level.  GPU only."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qpwcnet_b200 import ops  # noqa: E402


def timed(fn, n=30):
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
    out = []
    for (B, C, H, W) in ((8, 32, 224, 512), (8, 64, 112, 256), (8, 128, 56, 128)):
        g = torch.Generator(device="cuda").manual_seed(0)
        img = torch.randn((B, C, H, W), device="cuda", generator=g).requires_grad_()
        flo = (torch.randn((B, 2, H, W), device="cuda", generator=g) * 2).requires_grad_()
        go = torch.randn((B, C, H, W), device="cuda", generator=g)

        def native():
            o = ops.warp_nchw(img, flo, "tfa")
            torch.autograd.grad(o, (img, flo), go)

        def transposing():
            o = ops.warp(img.permute(0, 2, 3, 1).contiguous(), flo.permute(0, 2, 3, 1).contiguous(), "tfa")
            torch.autograd.grad(o.permute(0, 3, 1, 2).contiguous(), (img, flo), go)

        fwd = timed(lambda: ops.warp_nchw(img.detach(), flo.detach(), "tfa"))
        rec = {"shape": [B, C, H, W], "flow": "N(0,2^2) px per pixel", "native_fwd_us": fwd,
               "native_fwd_bwd_us": timed(native), "transposing_fwd_bwd_us": timed(transposing)}
        # a smooth flow (what a network predicts): neighbouring pixels sample neighbouring columns
        yy, xx = torch.meshgrid(torch.arange(H, device="cuda"), torch.arange(W, device="cuda"), indexing="ij")
        smooth = torch.stack([1.3 + 2.0 * torch.sin(yy / 37.0), -0.7 + 2.0 * torch.cos(xx / 53.0)])[None].expand(B, 2, H, W)
        flo = smooth.contiguous().requires_grad_()
        rec.update({"smooth_native_fwd_us": timed(lambda: ops.warp_nchw(img.detach(), flo.detach(), "tfa")),
                    "smooth_native_fwd_bwd_us": timed(native), "smooth_transposing_fwd_bwd_us": timed(transposing)})
        out.append(rec)
    json.dump(out, sys.stdout)
    print()


if __name__ == "__main__":
    main()
