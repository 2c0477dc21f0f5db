#!/usr/bin/env python3
"""Dev tool: time the tensor-core cost volume with parts switched off (QPWC_ABLATE, read once per process)."""
import os, subprocess, sys
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from qpwcnet_b200 import ops
    from tools.level_bench import timeit
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
    ops.set_corr_engine("tc")
    res = []
    for B, H, W, C in [(8, 224, 512, 32), (8, 112, 256, 64), (8, 28, 64, 256)]:
        prv = torch.randn((B, H, W, C), device="cuda"); nxt = torch.randn((B, H, W, C), device="cuda")
        out = torch.empty((B, H, W, 81), device="cuda")
        res.append(timeit(lambda: ops.cost_volume_into(out, prv, nxt, 4), 10, flush) * 1e6)
    print("  ".join(f"{t:8.1f}" for t in res))
else:
    names = {0: "full", 64: "per-row bulk copies instead of the tensor store", 1: "no drain", 16: "no copy-out", 17: "no epilogue", 2: "no split", 4: "no mma", 8: "no loads", 14: "epilogue only", 23: "only loads", 31: "nothing"}
    print(f"{'':28s} 224x512x32  112x256x64  28x64x256  (us)")
    for a, n in names.items():
        env = dict(os.environ, QPWC_ABLATE=str(a))
        r = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True, timeout=120)
        print(f"{n:28s} {r.stdout.strip()} {r.stderr.strip()[-200:] if r.returncode else ''}", flush=True)
