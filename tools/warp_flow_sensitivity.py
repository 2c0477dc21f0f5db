import os, sys, torch
sys.path.insert(0, os.getcwd())
from qpwcnet_b200 import ops
from tools.level_bench import timeit
B,H,W,C = 8,224,512,32
g = torch.Generator(device="cuda").manual_seed(0)
img = torch.randn((B,H,W,C), device="cuda", generator=g); out = torch.empty_like(img)
flush = torch.empty(64*1024*1024, dtype=torch.float32, device="cuda")
for name, flo in [("zero", torch.zeros((B,H,W,2), device="cuda")), ("const 1.3px", torch.full((B,H,W,2), 1.3, device="cuda")),
                  ("N(0,2^2)", torch.randn((B,H,W,2), device="cuda", generator=g)*2), ("N(0,8^2)", torch.randn((B,H,W,2), device="cuda", generator=g)*8)]:
    t = timeit(lambda: ops.warp_into(out, img, flo, "tfa"), 15, flush)
    print(f"flow {name:12s}: {t*1e6:6.1f} us")
t = timeit(lambda: out.copy_(img), 15, flush); print(f"torch copy       : {t*1e6:6.1f} us")
