"""One finest-level launch pair of the native channels_first gradient kernels, for ncu."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qpwcnet_b200 import ops
from qpwcnet_b200._cabi import check, lib
B, C, H, W = 8, 32, 224, 512
p = torch.randn((B, C, H, W), device="cuda"); n = torch.randn((B, C, H, W), device="cuda")
out = ops.cost_volume_nchw(p, n, 4); go = torch.randn_like(out)
gp, gn = torch.empty_like(p), torch.empty_like(n)
for _ in range(3):
    check(lib().qpwc_corr_bwd_nchw(p.data_ptr(), n.data_ptr(), out.data_ptr(), go.data_ptr(), gp.data_ptr(), gn.data_ptr(),
                                   B, C, H, W, 4, 0.1, torch.cuda.current_stream().cuda_stream))
torch.cuda.synchronize(); print("done")
