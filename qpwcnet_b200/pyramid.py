"""Benchmark harness: the cost-volume / warp call pattern of the reference's flow network at the
named shapes (BASELINE.json configs 2-5), with synthetic inputs.

``flower()`` (qpwcnet/core/pwcnet.py:28-67) runs, per forward pass, one plain cost volume at 1/32
scale (``Flow``, non_layers.py:332-338) and four warp -> cost-volume pairs at 1/16 .. 1/2 scale
(``UpFlow``, non_layers.py:366-387).  Feature widths follow the reference encoder/decoder
(pwcnet.py:145,179-207): 256 @1/32, then decoder outputs 256,128,64,32 @ 1/16,1/8,1/4,1/2.  The
network needs H, W divisible by 32, so 436x1024 is padded to 448x1024.
The conv stacks between the calls are out of scope (cuDNN territory); the harness feeds each level
synthetic features and flow of the right shape, which is exactly what the hot path sees.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

from . import ops

LEVEL_CHANNELS = (256, 256, 128, 64, 32)      # coarse -> fine
LEVEL_SCALES = (32, 16, 8, 4, 2)


@dataclass(frozen=True)
class Level:
    H: int
    W: int
    C: int
    fused: bool        # False: Flow (plain cost volume); True: UpFlow (warp -> cost volume)


def levels_for(height: int, width: int):
    Hp, Wp = -(-height // 32) * 32, -(-width // 32) * 32
    return tuple(Level(Hp // s, Wp // s, c, k > 0)
                 for k, (s, c) in enumerate(zip(LEVEL_SCALES, LEVEL_CHANNELS)))


def algorithmic_bytes(level: Level, B: int, d: int = 4) -> int:
    """4*(2C + D) per pixel (plain), 4*(2C + 2 + D) per pixel (fused)."""
    D = (2 * d + 1) ** 2
    return 4 * (2 * level.C + D + (2 if level.fused else 0)) * level.H * level.W * B


def algorithmic_flops(level: Level, B: int, d: int = 4) -> int:
    return 2 * (2 * d + 1) ** 2 * level.C * level.H * level.W * B


class PyramidWorkload:
    """Synthetic inputs + output buffers for one batch of frame pairs, on ``device`` ('cuda:N') or
    in pinned host memory (``device='cpu'``), and the 5-call hot path over them."""

    def __init__(self, height=436, width=1024, batch=8, search_range=4, device="cuda", seed=0,
                 flow_sigma=None, warp_mode="tfa", path="auto"):
        """path: how an UpFlow level (warp -> cost volume) is executed --
        'fused'    the library's fused entry point, one C-ABI call (tensor-core engine: warp kernel + cost-volume
                   kernel through a scratch buffer owned by the call; FFMA engine: one kernel);
        'composed' the warp call followed by the cost-volume call (scratch owned by the caller);
        'auto'     time both once per level on the device and keep the faster (device workloads)."""
        self.levels = levels_for(height, width)
        self.B, self.d, self.mode = batch, search_range, warp_mode
        self.D = (2 * search_range + 1) ** 2
        sigma = search_range / 2.0 if flow_sigma is None else flow_sigma
        g = torch.Generator(device="cpu").manual_seed(seed)
        on_host = torch.device(device).type == "cpu"
        self.inputs, self.outputs, self.scratch = [], [], []
        self.path = path
        for lv in self.levels:
            shp = (batch, lv.H, lv.W, lv.C)
            prv = torch.randn(shp, generator=g)
            nxt = torch.randn(shp, generator=g)
            flo = torch.randn((batch, lv.H, lv.W, 2), generator=g) * sigma if lv.fused else None
            out = torch.empty((batch, lv.H, lv.W, self.D))
            if on_host:
                ts = [t.pin_memory() if t is not None else None for t in (prv, nxt, flo)]
                out = out.pin_memory()
            else:
                ts = [t.to(device) if t is not None else None for t in (prv, nxt, flo)]
                out = out.to(device)
            self.inputs.append(tuple(ts))
            self.outputs.append(out)
            self.scratch.append(None)
        self.level_path = ["fused" if lv.fused else "plain" for lv in self.levels]
        if not on_host and path in ("composed", "auto"):
            for k, lv in enumerate(self.levels):
                if lv.fused:
                    self.scratch[k] = torch.empty_like(self.inputs[k][1])
                    self.level_path[k] = "composed"
            if path == "auto":
                self._autotune()

    # bytes per step
    def h2d_bytes(self):
        return sum(sum(t.numel() * 4 for t in ts if t is not None) for ts in self.inputs)

    def d2h_bytes(self):
        return sum(o.numel() * 4 for o in self.outputs)

    def algorithmic_bytes(self):
        return sum(algorithmic_bytes(lv, self.B, self.d) for lv in self.levels)

    def algorithmic_flops(self):
        return sum(algorithmic_flops(lv, self.B, self.d) for lv in self.levels)

    def run_level(self, k, how=None):
        prv, nxt, flo = self.inputs[k]
        how = how or self.level_path[k]
        if how == "fused":
            return ops.warp_cost_volume_into(self.outputs[k], prv, nxt, flo, self.mode, self.d)
        if how == "composed":
            ops.warp_into(self.scratch[k], nxt, flo, self.mode)
            return ops.cost_volume_into(self.outputs[k], prv, self.scratch[k], self.d)
        return ops.cost_volume_into(self.outputs[k], prv, nxt, self.d)

    def _autotune(self, reps=5):
        """Per UpFlow level: median device time of the library's fused entry point (one C-ABI call) vs the
        warp call followed by the cost-volume call."""
        self.autotune_ms = {}
        for k, lv in enumerate(self.levels):
            if not lv.fused:
                continue
            best = {}
            for how in ("fused", "composed"):
                for _ in range(2):
                    self.run_level(k, how)
                ts = []
                for _ in range(reps):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(4):                      # several launches per sample: hides host latency
                        self.run_level(k, how)
                    e1.record()
                    torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1) / 4)
                best[how] = sorted(ts)[len(ts) // 2]
            self.autotune_ms[f"{lv.H}x{lv.W}x{lv.C}"] = best
            self.level_path[k] = min(best, key=best.get)

    @property
    def launches_per_step(self):
        # kernels per level: warp + cost volume for an UpFlow level (the library's fused entry point runs the
        # same two kernels under the tensor-core engine; the FFMA engine's in-kernel fusion is one launch)
        two = ops.get_corr_engine() != "ffma"
        return sum(2 if (p == "composed" or (p == "fused" and two)) else 1 for p in self.level_path)

    def step(self):
        """One pass of the hot path over the batch: 1 cost volume + 4 fused warp->cost volumes
        (five C-ABI launches on the current stream)."""
        if not self.outputs[0].is_cuda:
            with ops.host_batch():      # host buffers: let the five levels' copies/kernels overlap
                # finest level first: the D2H engine can start after ONE slice of input has arrived, and the
                # slices that are still on the wire when the H2D engine falls idle are the small coarse ones
                for k in reversed(range(len(self.levels))):
                    self.run_level(k)
            return self.outputs
        for k in range(len(self.levels)):
            self.run_level(k)
        return self.outputs

    def step_concurrent(self, streams):
        """The same five levels, each on its own stream, forked from and joined back into the current
        stream: the synthetic levels carry no data dependence on each other, so the launch-bound
        coarse levels (28 and 112 CTAs) fill SMs the fine levels' tails leave idle.  Finest first."""
        cur = torch.cuda.current_stream()
        for k in reversed(range(len(self.levels))):
            s = streams[k]
            s.wait_stream(cur)
            with torch.cuda.stream(s):
                self.run_level(k)
        for s in streams:
            cur.wait_stream(s)
        return self.outputs

    def capture(self, concurrent=False):
        """Record step() into a CUDA graph (device-resident workloads only): the five launches --
        TMA descriptors included, they are plain kernel parameters -- replay with one
        cudaGraphLaunch, which removes the per-call host overhead from the launch-bound coarse
        levels.  Returns self; use replay() afterwards."""
        if not self.outputs[0].is_cuda:
            raise RuntimeError("CUDA graphs need device-resident buffers")
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):          # warm-up off the capture: smem attributes, entry points
            for _ in range(2):
                self.step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        if concurrent:
            self._streams = [torch.cuda.Stream() for _ in self.levels]
        with torch.cuda.graph(graph):
            if concurrent:
                self.step_concurrent(self._streams)
            else:
                self.step()
        if not hasattr(self, "_graphs"):
            self._graphs = {}
        self._graphs["branches" if concurrent else "serial"] = graph
        self._graph = graph
        return self

    def capture_levels(self):
        """One small CUDA graph per level (same kernels as step()): lets a caller time a single level between
        events without the host's launch latency sitting between its warp and cost-volume kernels."""
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for k in range(len(self.levels)):
                self.run_level(k)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self._level_graphs = []
        for k in range(len(self.levels)):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.run_level(k)
            self._level_graphs.append(g)
        return self._level_graphs

    def replay(self, which=None):
        """Replay the most recently captured graph, or the one named 'serial' / 'branches'."""
        (self._graph if which is None else self._graphs[which]).replay()
        return self.outputs

