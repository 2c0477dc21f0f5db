import os, sys, time, torch
sys.path.insert(0, ".")
os.environ["QPWC_ABLATE"] = "7"
from qpwcnet_b200 import ops, _cabi
B,H,W,C = 8,28,64,256
prv = torch.randn((B,H,W,C), device="cuda"); nxt = torch.randn((B,H,W,C), device="cuda"); out = torch.empty((B,H,W,81), device="cuda")
def bench(f, n=2000):
    for _ in range(50): f()
    torch.cuda.synchronize(); t=time.perf_counter()
    for _ in range(n): f()
    dt=(time.perf_counter()-t)/n; torch.cuda.synchronize(); return dt*1e6
print("cost_volume_into total      %.1f us" % bench(lambda: ops.cost_volume_into(out, prv, nxt, 4)))
print("dlview x3                   %.1f us" % bench(lambda: (_cabi.dlview(prv), _cabi.dlview(nxt), _cabi.dlview(out))))
print("to_dlpack only x3           %.1f us" % bench(lambda: (torch.utils.dlpack.to_dlpack(prv), torch.utils.dlpack.to_dlpack(nxt), torch.utils.dlpack.to_dlpack(out))))
print("detach x3                   %.1f us" % bench(lambda: (prv.detach(), nxt.detach(), out.detach())))
print("torch.cuda.device ctx       %.1f us" % bench(lambda: torch.cuda.device(prv.device).__enter__()))
print("current_stream.cuda_stream  %.1f us" % bench(lambda: torch.cuda.current_stream(prv.device).cuda_stream))
print("_prep x2 + checks           %.1f us" % bench(lambda: (ops._prep(prv,"p"), ops._prep(nxt,"n"), ops._check_out(out, prv))))
L=_cabi.lib(); vp,vn,vo=_cabi.dlview(prv),_cabi.dlview(nxt),_cabi.dlview(out); st=torch.cuda.current_stream().cuda_stream
print("raw ctypes qpwc_corr_fwd    %.1f us" % bench(lambda: L.qpwc_corr_fwd(vp.ptr, vn.ptr, vo.ptr, B,H,W,C,4,0.1,81,st)))
