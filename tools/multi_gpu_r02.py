#!/usr/bin/env python3
"""Round-2 multi-GPU measurements (torchrun, NCCL over NVLink).  Dev/measurement tool.

  pcie     : concurrent pinned H2D / D2H bandwidth per rank with all ranks copying at once -- the
             ceiling the e2e leg of bench.py has to be graded against.
  config 5 : one 3840x2160 frame pair (2176 rows), pyramid levels row-sharded over the ranks with
             `sharded.ShardedLevel` (halo rows received in place, interior rows computed while the
             exchange is in flight).  Every rank checks its band against the unsharded op; timings
             separate the exchange from the compute; the unsharded single-GPU time is the yardstick.

  torchrun --nproc-per-node N --master-addr 127.0.0.1 tools/multi_gpu_r02.py [--skip-pcie]
"""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qpwcnet_b200 import ops, sharded  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
os.environ.setdefault("MASTER_PORT", "29533")
dist.init_process_group("nccl", device_id=dev)


def maxr(x):
    t = torch.tensor([x], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def minr(x):
    return -maxr(-x)


def timed(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return maxr(e0.elapsed_time(e1) / iters)


res = {"world": world}

# ------------------------------------------------------------------------------- PCIe ceiling
if "--skip-pcie" not in sys.argv:
    n = 256 << 20
    h_in, h_out = torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory()
    d_in, d_out = torch.empty(n, dtype=torch.uint8, device=dev), torch.empty(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def pcie(h2d, d2h, reps=6):
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            if h2d:
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        return n * reps / (time.perf_counter() - t0) / 1e9

    pcie(True, True, 2)
    a, b, c = pcie(True, False), pcie(False, True), pcie(True, True)
    res["pcie_gbs_per_rank_all_ranks_concurrent"] = {
        "h2d_alone": [minr(a), maxr(a)], "d2h_alone": [minr(b), maxr(b)], "both_each_direction": [minr(c), maxr(c)],
        "note": "[slowest rank, fastest rank]; every rank copies at the same time"}
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        res["cpu_affinity_words_rank0"] = [int(x) for x in pynvml.nvmlDeviceGetCpuAffinity(h, 4)] if rank == 0 else None
    except Exception:
        pass
    del h_in, h_out, d_in, d_out

# ---------------------------------------------------------------- config 5: 4K frame, row-sharded
g = torch.Generator(device="cpu").manual_seed(0)
c5 = []
for (H, W, C) in [(68, 120, 256), (136, 240, 256), (272, 480, 128), (544, 960, 64), (1088, 1920, 32)]:
    prv, nxt = torch.randn((1, H, W, C), generator=g), torch.randn((1, H, W, C), generator=g)
    hband = H // world
    reach = max(0, min(8, hband - 4 - 1))
    flo = (torch.randn((1, H, W, 2), generator=g) * 2).clamp_(-reach, reach)
    row = {"level": f"{H}x{W}x{C}", "band_rows": hband, "halo_bytes_per_neighbour": 4 * W * C * 4, "pair_flow_reach_rows": reach}
    full_p, full_n, full_f = prv.to(dev), nxt.to(dev), flo.to(dev)
    ref_cv = ops.cost_volume(full_p, full_n, 4)
    ref_pair = ops.warp_cost_volume(full_p, full_n, full_f, "tfa", 4)
    out1 = torch.empty_like(ref_cv)
    row["one_gpu_cv_ms"] = timed(lambda: ops.cost_volume_into(out1, full_p, full_n, 4))
    row["one_gpu_pair_ms"] = timed(lambda: ops.warp_cost_volume_into(out1, full_p, full_n, full_f, "tfa", 4))
    for pair, symm in ((False, False), (True, False), (False, True), (True, True)):
        key = ("pair" if pair else "cv") + ("_nvlink_push" if symm else "_nccl_p2p")
        if symm and "--no-symm" in sys.argv:
            continue
        try:
            lv = sharded.ShardedLevel(H, W, C, 4, reach=reach, pair=pair, group=None, symmetric=symm)
        except ValueError as e:
            row[key] = {"skipped": str(e)}
            continue
        lv.prv.copy_(full_p[:, lv.r0:lv.r1]); lv.nxt.copy_(full_n[:, lv.r0:lv.r1])
        if pair:
            lv.flow.copy_(full_f[:, lv.r0:lv.r1])
        out = lv.run()
        torch.cuda.synchronize()
        ref = (ref_pair if pair else ref_cv)[:, lv.r0:lv.r1]
        err = float((out - ref).abs().max() / ref.abs().max())
        ident = bool(torch.equal(out, ref))
        t_all = timed(lv.run)
        t_graph, ident_g = None, None
        if symm and "--graph" in sys.argv:
            lv.capture()
            ident_g = bool(minr(1.0 if torch.equal(lv.replay(), ref) else 0.0) == 1.0)
            t_graph = timed(lv.replay)
        if symm:
            t_x = timed(lv._exchange_symmetric)
        else:
            t_x = timed(lambda: [q.wait() for q in lv._exchange([lv.nxt_h] + ([lv.flow_h] if pair else []))])
        # compute alone: the same kernel sequence with the neighbours switched off
        up, down = lv.up, lv.down
        lv.up, lv.down = -1, lv.world
        t_c = timed(lv.run)
        lv.up, lv.down = up, down
        row[key] = {"sharded_ms": t_all, "cuda_graph_ms": t_graph, "cuda_graph_bit_identical": ident_g,  "exchange_alone_ms": t_x, "compute_alone_ms": t_c,
                    "rel_err_worst_rank": maxr(err), "bit_identical_on_every_rank": bool(minr(1.0 if ident else 0.0) == 1.0),
                    "speedup_vs_one_gpu": row["one_gpu_pair_ms" if pair else "one_gpu_cv_ms"] / (t_graph or t_all)}
    c5.append(row)
res["config5_rowsharded_4k"] = c5
if rank == 0:
    print(json.dumps(res, indent=1))
dist.destroy_process_group()
