import os, sys, torch
sys.path.insert(0, os.getcwd())
from qpwcnet_b200 import ops
from tools.level_bench import timeit
B,H,W,C = 8,224,512,32
g = torch.Generator(device="cuda").manual_seed(0)
prv = torch.randn((B,H,W,C), device="cuda", generator=g); nxt = torch.randn((B,H,W,C), device="cuda", generator=g)
out = torch.empty((B,H,W,81), device="cuda"); flush = torch.empty(64*1024*1024, dtype=torch.float32, device="cuda")
t = timeit(lambda: ops.cost_volume_into(out, prv, nxt, 4), 12, flush)
print(f"QPWC_ABLATE={os.environ.get('QPWC_ABLATE','0')}: {t*1e6:.1f} us")
