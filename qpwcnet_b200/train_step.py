"""Hot path of one frame-interpolation pre-training step (BASELINE.json config 3), forward AND backward.

`TrainModel.train_step` (qpwcnet/app/frame_interpolation/pre_train.py:54-72) runs `build_interpolator`
(qpwcnet/core/pwcnet.py:247-287): the shared-weight `Flower` block is applied twice (nxt->prv and
prv->nxt, pwcnet.py:270-280) -- per application one plain cost volume at 1/32 scale and four UpFlow
warp -> cost-volume pairs (pwcnet.py:28-67) -- and `interpolator()` (pwcnet.py:101-122) adds five
`FrameInterpolate` blocks, each with two half-flow warps (non_layers.py:303-304): 10 cost volumes and
18 warps per step, and the same again in the backward pass.  Shapes: 256x448 triplets (BASELINE.json;
the reference script defaults to 256x512, pre_train.py:38), levels 8x14x256 ... 128x224x32, and the
five interpolation levels warp C = 3, 256, 128, 64, 32 channels.

The conv stacks between the calls are out of scope (cuDNN territory), so the harness feeds every call
synthetic features / flows of the right shape and drives the backward pass with synthetic upstream
gradients -- exactly what the hot path sees.  Data-parallel training: every rank runs its share of the
global batch; the only collective of the step is the all-reduce of the model's gradients (3.1 M fp32
parameters), issued here per pyramid level as soon as that level's backward calls have
run, on a side stream, so that it overlaps the rest of the backward pass.  The bucket VALUES are
synthetic (no conv stack produces them); their sizes follow the parameter split of the reference
model (encoder/decoder/flow blocks, pwcnet.py:145,179 and non_layers.py:213-273).
"""
from __future__ import annotations

import torch

from . import ops

LEVELS = ((8, 14, 256), (16, 28, 256), (32, 56, 128), (64, 112, 64), (128, 224, 32))   # coarse -> fine
# FrameInterpolate inputs, coarse -> fine: img_0 warps the 1/32-scale images (3 channels), img_1..4 the
# decoder features (pwcnet.py:101-121)
INTERP_CHANNELS = (3, 256, 128, 64, 32)
# gradient buckets, one per pyramid level, fine -> coarse = the order in which backward finishes them:
# the deep (coarse) blocks hold most of the 3.1 M parameters
BUCKET_ELEMS = (60_000, 190_000, 520_000, 1_030_000, 1_300_000)


class TrainHotPath:
    def __init__(self, per_gpu_batch: int, device, seed: int = 0, mode: str = "tfa", search_range: int = 4,
                 process_group=None, world: int = 1):
        self.B, self.d, self.mode, self.world, self.pg = per_gpu_batch, search_range, mode, world, process_group
        g = torch.Generator(device="cpu").manual_seed(seed)
        dev = torch.device(device)

        def leaf(*shape, scale=1.0, uniform=False):
            t = (torch.rand(shape, generator=g) if uniform else torch.randn(shape, generator=g) * scale)
            return t.to(dev).requires_grad_()

        B, D = per_gpu_batch, (2 * search_range + 1) ** 2
        self.flower = []          # two applications of Flower: per level (prv, nxt, flow, upstream gradient)
        for _ in range(2):
            lv = []
            for k, (H, W, C) in enumerate(LEVELS):
                lv.append((leaf(B, H, W, C), leaf(B, H, W, C), leaf(B, H, W, 2, scale=search_range / 2.0) if k else None,
                           torch.randn((B, H, W, D), generator=g).to(dev)))
            self.flower.append(lv)
        self.interp = []          # FrameInterpolate: (prv, nxt, flo_01, flo_10, upstream gradient of the concat)
        for (H, W, _), C in zip(LEVELS, INTERP_CHANNELS):
            self.interp.append((leaf(B, H, W, C, uniform=True), leaf(B, H, W, C, uniform=True),
                                leaf(B, H, W, 2, scale=2.0), leaf(B, H, W, 2, scale=2.0),
                                torch.randn((B, H, W, 2 * C), generator=g).to(dev)))
        self.buckets = [torch.zeros(n, device=dev) for n in BUCKET_ELEMS]
        self.comm = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None
        self.allreduce = world > 1
        self.counts = {"cost_volume": 10, "warp": 18}
        self.bucketing = "per_level"
        self._graph = None

    # ---- algorithmic work of one step (forward + backward), SURVEY 8(d)
    def algorithmic_bytes(self):
        D, tot = (2 * self.d + 1) ** 2, 0
        for k, (H, W, C) in enumerate(LEVELS):
            px = self.B * H * W
            fwd = 4 * (2 * C + D + (2 if k else 0)) * px
            bwd = 4 * (D + 2 * C + 2 * C + (4 if k else 0)) * px
            tot += 2 * (fwd + bwd)
        for (H, W, _), C in zip(LEVELS, INTERP_CHANNELS):
            px = self.B * H * W
            tot += 2 * (4 * (2 * C + 2) * px + 4 * (3 * C + 4) * px)
        return tot

    def allreduce_bytes(self):
        return 4 * sum(BUCKET_ELEMS)

    def _forward(self):
        outs, grads = [], []
        for lv in self.flower:
            for k, (p, n, f, g) in enumerate(lv):
                outs.append(ops.cost_volume(p, n, self.d) if k == 0 else ops.warp_cost_volume(p, n, f, self.mode, self.d))
                grads.append(g)
        for (pa, nb, f01, f10, g) in self.interp:
            outs.append(ops.half_flow_warps(pa, nb, f01, f10, self.mode))
            grads.append(g)
        return outs, grads

    def _backward_level(self, outs, grads, k):
        nl = len(LEVELS)
        idx = [k, nl + k, 2 * nl + k]
        torch.autograd.backward([outs[i] for i in idx], [grads[i] for i in idx])

    def _allreduce_bucket(self, j):
        self.comm.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.comm):
            torch.distributed.all_reduce(self.buckets[j], group=self.pg)

    def step(self):
        """Forward + backward of the 10 cost volumes and 18 warps, finest level first in the backward pass
        (the order autograd of the real network produces: the loss sits on the full-resolution outputs).
        bucketing = 'per_level': the gradient bucket of a level goes to the comm stream as soon as that
        level's backward calls are enqueued; 'single': one all-reduce of all buckets after the backward
        pass (fewer host-side NCCL launches: better when the step is launch-bound, i.e. small per-GPU
        batches).  With a captured graph (capture()) the compute is one cudaGraphLaunch."""
        nl = len(LEVELS)
        if self._graph is not None:
            self._graph.replay()
            if self.allreduce:
                for j in range(nl):
                    self._allreduce_bucket(j)
        else:
            outs, grads = self._forward()
            for j, k in enumerate(range(nl - 1, -1, -1)):
                self._backward_level(outs, grads, k)
                if self.allreduce and self.bucketing == "per_level":
                    self._allreduce_bucket(j)
            if self.allreduce and self.bucketing == "single":
                for j in range(nl):
                    self._allreduce_bucket(j)
        if self.allreduce:
            torch.cuda.current_stream().wait_stream(self.comm)

    def capture(self):
        """Record forward + backward into one CUDA graph (gradients are produced into graph-owned
        buffers; the NCCL all-reduces stay outside the graph and follow the replay)."""
        self.zero_grad()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                outs, grads = self._forward()
                for k in range(len(LEVELS) - 1, -1, -1):
                    self._backward_level(outs, grads, k)
                self.zero_grad()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            outs, grads = self._forward()
            for k in range(len(LEVELS) - 1, -1, -1):
                self._backward_level(outs, grads, k)
        self._graph = g
        return self

    def zero_grad(self):
        if self._graph is not None:
            return                     # graph-owned gradient buffers are overwritten by every replay
        for lv in self.flower:
            for (p, n, f, _) in lv:
                for t in (p, n, f):
                    if t is not None:
                        t.grad = None
        for (pa, nb, f01, f10, _) in self.interp:
            for t in (pa, nb, f01, f10):
                t.grad = None
