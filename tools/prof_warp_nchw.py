"""Dev: channels_first warp forward/backward at the finest level for ncu (argv[1]: flow sigma in px, 0 = smooth ramp)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qpwcnet_b200 import ops
sig = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
B, C, H, W = 8, 32, 224, 512
img = torch.randn((B, C, H, W), device="cuda").requires_grad_()
if sig > 0:
    flo = (torch.randn((B, 2, H, W), device="cuda") * sig).requires_grad_()
else:
    yy, xx = torch.meshgrid(torch.arange(H, device="cuda"), torch.arange(W, device="cuda"), indexing="ij")
    flo = torch.stack([2.5 * torch.sin(xx / 40.0), 1.5 * torch.cos(yy / 30.0)])[None].repeat(B, 1, 1, 1).contiguous().requires_grad_()
go = torch.randn((B, C, H, W), device="cuda")
for _ in range(3):
    o = ops.warp_nchw(img, flo, "tfa")
    torch.autograd.grad(o, (img, flo), go)
torch.cuda.synchronize(); print("done")
