#!/usr/bin/env python3
"""Dev: time line of CTA 0 of the resident tensor-core kernel (clock64 stamps at the hand-over points)."""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qpwcnet_b200 import ops, _cabi
C = int(sys.argv[1]) if len(sys.argv) > 1 else 32
FUSED = len(sys.argv) > 2 and sys.argv[2] == "fused"
SIGMA = float(sys.argv[3]) if len(sys.argv) > 3 else 2.0
prv = torch.randn((8, 224, 512, C), device="cuda"); nxt = torch.randn((8, 224, 512, C), device="cuda")
out = torch.empty((8, 224, 512, 81), device="cuda")
flo = torch.randn((8, 224, 512, 2), device="cuda") * SIGMA
def run():
    if FUSED: ops.warp_cost_volume_into(out, prv, nxt, flo, "tfa", 4)
    else: ops.cost_volume_into(out, prv, nxt, 4)
ops.set_corr_engine("tc")
for _ in range(3):
    run()
buf = torch.zeros(40 * 24, dtype=torch.int64, device="cuda")
L = _cabi.lib()
L.qpwc_debug_tc_trace.argtypes = [ctypes.c_void_p]
assert L.qpwc_debug_tc_trace(buf.data_ptr()) == 0
run()
torch.cuda.synchronize()
L.qpwc_debug_tc_trace(None)
t = buf.cpu().view(40, 24)
names = ["tma:A issue", "tma:B issue", "spl:A landed", "spl:A->tmem", "spl:B landed", "spl:B lo done", "mma:h0 start", "mma:h0 issued",
         "mma:h1 start", "mma:h1 issued", "epi0:tfull0", "epi0:loaded", "epi0:sfree", "epi0:staged", "epi0:bar2", "epi1:tfull1", "epi1:loaded", "epi1:staged"]
t0 = int(t[8, 6])
print("clk relative to mma:h0 start of tile 8; rows = tiles 8..19")
print("tile " + " ".join(f"{n[:12]:>12s}" for n in names))
for k in range(8, 20):
    print(f"{k:4d} " + " ".join(f"{int(t[k, e]) - t0:12d}" for e in range(18)))
per = (int(t[30, 6]) - int(t[10, 6])) / 20
print(f"steady-state period: {per:.0f} clk per tile")
