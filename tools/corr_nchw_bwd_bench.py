"""channels_first cost-volume gradients: native NCHW kernels (qpwc_corr_bwd_nchw.cu) against the
transposing route (permute -> NHWC tiled gradient kernels -> permute), per pyramid level.  GPU only."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qpwcnet_b200 import ops  # noqa: E402
from qpwcnet_b200._cabi import check, lib  # noqa: E402


def timed(fn, n=20):
    for _ in range(3):
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
    res = []
    for (B, C, H, W) in ((8, 32, 224, 512), (8, 64, 112, 256), (8, 128, 56, 128), (8, 256, 28, 64)):
        g = torch.Generator(device="cuda").manual_seed(0)
        prv = torch.randn((B, C, H, W), device="cuda", generator=g)
        nxt = torch.randn((B, C, H, W), device="cuda", generator=g)
        out = ops.cost_volume_nchw(prv, nxt, 4)
        go = torch.randn_like(out)
        gp, gn = torch.empty_like(prv), torch.empty_like(nxt)
        s = torch.cuda.current_stream().cuda_stream

        def native():
            check(lib().qpwc_corr_bwd_nchw(prv.data_ptr(), nxt.data_ptr(), out.data_ptr(), go.data_ptr(),
                                           gp.data_ptr(), gn.data_ptr(), B, C, H, W, 4, 0.1, s))

        nh = lambda t: t.permute(0, 2, 3, 1).contiguous()

        def transposing():
            a, b = ops._corr_bwd(nh(prv), nh(nxt), nh(out), nh(go), 4, 0.1)
            return a.permute(0, 3, 1, 2).contiguous(), b.permute(0, 3, 1, 2).contiguous()

        native()
        a, b = transposing()
        err = max(float((gp - a).abs().max() / a.abs().max()), float((gn - b).abs().max() / b.abs().max()))
        flops = 2 * 2 * B * H * W * 81 * C
        t_n, t_t = timed(native), timed(transposing)
        res.append({"shape": [B, C, H, W], "native_us": t_n, "transposing_us": t_t, "native_TFLOPs": flops / t_n / 1e6,
                    "max_rel_diff": err})
    json.dump(res, sys.stdout)
    print()


if __name__ == "__main__":
    main()
