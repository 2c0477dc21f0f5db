#!/usr/bin/env python3
"""BASELINE.json configs 3 and 5 on N GPUs (torchrun), NCCL over NVLink.  Dev/measurement tool.

  config 5: one 3840x2160 frame pair (padded to 2176 rows), pyramid level shapes, row-sharded over
            the ranks with a d-row halo exchange of `nxt` (qpwcnet_b200.sharded); checked against the
            unsharded op on rank 0 and timed (device time, max over ranks).
  config 3: hot path of one frame-interpolation training step at 256x448, global batch 64 split over
            the ranks: 10 cost volumes + 18 warps forward and backward, plus one NCCL all-reduce of
            3.1 M fp32 "gradients" (12.5 MB) -- the only collective of the step.

  torchrun --nproc-per-node N --master-addr 127.0.0.1 tools/multi_gpu_configs.py
"""
import json
import os
import sys

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


def timed(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


res = {"world": world}
# ---------------------------------------------------------------- config 5: 4K frame, row-sharded
g = torch.Generator(device="cpu").manual_seed(0)
c5 = []
for (H, W, C) in [(68, 120, 256), (136, 240, 256), (272, 480, 128), (544, 960, 64), (1088, 1920, 32)]:
    prv, nxt = torch.randn((1, H, W, C), generator=g), torch.randn((1, H, W, C), generator=g)
    flo = torch.randn((1, H, W, 2), generator=g) * 2
    r0, r1 = sharded.band(H, rank, world)
    pb, nb, fb = (t[:, r0:r1].contiguous().to(dev) for t in (prv, nxt, flo))
    out = sharded.cost_volume(pb, nb, 4)
    outf = sharded.warp_cost_volume(pb, nb, fb, "tfa", 4)
    ok = ok_f = None
    if rank == 0:
        ref = ops.cost_volume(prv.to(dev), nxt.to(dev), 4)[:, r0:r1]
        ok = bool(torch.equal(out, ref))
        reff = ops.warp_cost_volume(prv.to(dev), nxt.to(dev), flo.to(dev), "tfa", 4)[:, r0:r1]
        ok_f = float((outf - reff).abs().max() / reff.abs().max())
    t_cv = timed(lambda: sharded.cost_volume(pb, nb, 4))
    t_f = timed(lambda: sharded.warp_cost_volume(pb, nb, fb, "tfa", 4))
    c5.append({"level": f"{H}x{W}x{C}", "halo_bytes": 4 * W * C * 4, "sharded_cv_ms": t_cv, "sharded_fused_ms": t_f,
               "cv_bit_identical_rank0": ok, "fused_rel_err_rank0": ok_f})
res["config5_rowsharded_4k"] = c5

# ---------------------------------------------------------------- config 3: training-step hot path
per = 64 // world
levels = [(8, 14, 256), (16, 28, 256), (32, 56, 128), (64, 112, 64), (128, 224, 32)]
ten = []
for (H, W, C) in levels:
    mk = lambda *s: torch.randn(s, device=dev).requires_grad_()
    ten.append((mk(per, H, W, C), mk(per, H, W, C), mk(per, H, W, 2)))
img = [(torch.rand((per, H, W, c), device=dev).requires_grad_(), torch.rand((per, H, W, c), device=dev).requires_grad_(),
        torch.randn((per, H, W, 2), device=dev).requires_grad_(), torch.randn((per, H, W, 2), device=dev).requires_grad_())
       for (H, W, _), c in zip(levels, (3, 256, 128, 64, 32))]
grads = torch.zeros(3_100_000, device=dev)
comm = torch.cuda.Stream()


def train_hot_path():
    losses = []
    for _pass in range(2):                                  # Flower is applied twice (pwcnet.py:270-280)
        for k, (p, n, f) in enumerate(ten):
            cv = ops.cost_volume(p, n, 4) if k == 0 else ops.warp_cost_volume(p, n, f, "tfa", 4)
            losses.append(cv.mean())
    for (pa, nb_, f01, f10) in img:                          # FrameInterpolate: two half-flow warps/level,
        losses.append(ops.half_flow_warps(pa, nb_, f01, f10, "tfa").mean())   # one launch, 0.5 x fused
    for k in range(1, 5):                                    # the 8 UpFlow warps are inside the fused op
        pass
    torch.stack(losses).sum().backward()
    comm.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(comm):                            # the step's only collective
        dist.all_reduce(grads)
    torch.cuda.current_stream().wait_stream(comm)


t3 = timed(train_hot_path, iters=5, warm=2)
res["config3_train_hotpath"] = {"global_batch": 64, "per_gpu_batch": per, "ms_per_step": t3,
                                "triplets_per_s": 64 / (t3 * 1e-3), "allreduce_bytes": grads.numel() * 4}
if rank == 0:
    print(json.dumps(res, indent=1))
dist.destroy_process_group()
