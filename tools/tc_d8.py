"""Dev: search range 8 on the tensor-core engine (four 9x9 windows) vs FFMA, at config 4's sizes."""
import sys, torch, numpy as np
sys.path.insert(0, ".")
from qpwcnet_b200 import ops
import oracle  # noqa: E402  (dev check only)

def t(fn, n=20):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

g = torch.Generator(device="cuda").manual_seed(0)
for (B, H, W, C) in [(1, 40, 56, 16), (8, 218, 512, 16), (8, 109, 256, 32), (8, 224, 512, 32), (8, 56, 128, 64)]:
    prv = torch.randn((B, H, W, C), device="cuda", generator=g)
    nxt = torch.randn((B, H, W, C), device="cuda", generator=g)
    out = torch.empty((B, H, W, 289), device="cuda")
    res = {}
    for eng in ("ffma", "tc"):
        ops.set_corr_engine(eng)
        us = t(lambda: ops.cost_volume_into(out, prv, nxt, 8))
        res[eng] = (us, out.clone())
    line = f"{B}x{H}x{W}x{C} d=8: ffma {res['ffma'][0]:8.1f} us   tc {res['tc'][0]:8.1f} us   max|tc-ffma| {(res['tc'][1]-res['ffma'][1]).abs().max().item():.2e}"
    if B * H * W <= 4000:
        ref = oracle.cost_volume(prv.cpu().numpy().astype(np.float64), nxt.cpu().numpy().astype(np.float64), 8)
        line += f"   max|tc-oracle| {np.abs(res['tc'][1].cpu().numpy() - ref).max():.2e}"
    by = 4 * (2 * C + 289) * B * H * W
    line += f"   alg {by / res['tc'][0] / 1e3:.0f} GB/s"
    print(line)
