"""Row-sharded cost volume / warp->cost volume for single frames too large for one pass per GPU
(BASELINE.json config 5: one 3840x2160 pair over 8 GPUs) -- SURVEY.md 8(e).

The image height is split into contiguous row bands, one per rank.  The correlation reads the second
frame up to `d` rows above/below a band, so each rank exchanges **d rows of `nxt`** with its two
neighbours (NCCL send/recv over NVLink; `d*W*C*4` bytes per message), computes its band with the
ordinary kernels on the halo-padded tensors and keeps the interior rows.  Image-border ranks have no
neighbour on one side: the kernels' zero padding (ZeroPadding2D semantics) applies there, while
interior cuts are filled by the halo, so the plain cost volume is bit-identical to the unsharded op.

The fused warp->cost volume additionally samples `nxt` at `row + flow_y`: its halo is
`d + ceil(max|flow_y|) + 1` rows, agreed on with one all-reduce(MAX) of that scalar.  Its result
matches the unsharded op to fp32 rounding of the sampling coordinate (the reference adds the flow to
an absolute row index; a band-local index rounds ~1 ulp differently) -- well inside the 1e-5 bound.

Only the exchange is communication; there is no collective on the per-pixel data path.
"""
from __future__ import annotations

import math

import torch
import torch.distributed as dist


def band(H: int, rank: int, world: int):
    """Rows [r0, r1) of a height-H image owned by `rank` (contiguous, near-equal bands)."""
    base, rem = divmod(H, world)
    r0 = rank * base + min(rank, rem)
    return r0, r0 + base + (1 if rank < rem else 0)


def exchange_halo(x: torch.Tensor, rows: int, group=None) -> torch.Tensor:
    """x: this rank's band (B, h, W, C).  Returns (B, top + h + bottom, W, C) where `top`/`bottom`
    are `rows` rows received from the previous/next rank (0 rows at the image border)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if rows <= 0 or world == 1:
        return x
    if x.shape[1] < rows:
        raise ValueError(f"band of {x.shape[1]} rows is thinner than the {rows}-row halo")
    up, down = rank - 1, rank + 1
    ops, bufs = [], {}
    if up >= 0:
        bufs["top"] = torch.empty_like(x[:, :rows])
        ops.append(dist.P2POp(dist.isend, x[:, :rows].contiguous(), up, group))
        ops.append(dist.P2POp(dist.irecv, bufs["top"], up, group))
    if down < world:
        bufs["bot"] = torch.empty_like(x[:, -rows:])
        ops.append(dist.P2POp(dist.isend, x[:, -rows:].contiguous(), down, group))
        ops.append(dist.P2POp(dist.irecv, bufs["bot"], down, group))
    for req in dist.batch_isend_irecv(ops):
        req.wait()
    parts = ([bufs["top"]] if "top" in bufs else []) + [x] + ([bufs["bot"]] if "bot" in bufs else [])
    return torch.cat(parts, dim=1)


def _crop(out, rank, world, rows):
    top = rows if rank > 0 else 0
    bot = rows if rank < world - 1 else 0
    return out[:, top:out.shape[1] - bot]


def cost_volume(prv_band, nxt_band, search_range=4, leaky_slope=0.1, group=None, op=None):
    """Cost volume of this rank's row band; identical to the band of the unsharded result."""
    from . import ops as _ops
    op = op or _ops.cost_volume
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    d = int(search_range)
    nxt_h = exchange_halo(nxt_band, d, group)
    # prv needs no neighbour data; pad it with zeros only to keep the two tensors the same shape
    # (rows of the halo produce outputs that are cropped away)
    top = d if rank > 0 else 0
    bot = d if rank < world - 1 else 0
    prv_h = torch.nn.functional.pad(prv_band, (0, 0, 0, 0, top, bot))
    return _crop(op(prv_h, nxt_h, d, leaky_slope), rank, world, d).contiguous()


def warp_cost_volume(prv_band, nxt_band, flow_band, mode="tfa", search_range=4, leaky_slope=0.1,
                     group=None, op=None, row_offset_fix=True):
    """Fused warp -> cost volume of this rank's band.  The warp samples absolute image rows, so the
    padded band is processed with the flow expressed in band coordinates (unchanged: the flow is a
    displacement) and a halo deep enough for the largest vertical displacement."""
    from . import ops as _ops
    op = op or _ops.warp_cost_volume
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    d = int(search_range)
    m = flow_band[..., 1].abs().max().reshape(1).to(torch.float32)
    dist.all_reduce(m, op=dist.ReduceOp.MAX, group=group)
    rows = d + int(math.ceil(float(m.item()))) + 1
    nxt_h = exchange_halo(nxt_band, rows, group)
    flow_h = exchange_halo(flow_band, rows, group)
    top = rows if rank > 0 else 0
    bot = rows if rank < world - 1 else 0
    prv_h = torch.nn.functional.pad(prv_band, (0, 0, 0, 0, top, bot))
    return _crop(op(prv_h, nxt_h, flow_h, mode, d, leaky_slope), rank, world, rows).contiguous()
