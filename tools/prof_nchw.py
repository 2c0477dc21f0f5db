import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qpwcnet_b200 import ops
B, C, H, W = 8, 32, 224, 512
p = torch.randn((B, C, H, W), device="cuda"); n = torch.randn((B, C, H, W), device="cuda")
for _ in range(3): ops._corr_fwd_nchw(p, n, 4, 0.1)
torch.cuda.synchronize(); print("done")
