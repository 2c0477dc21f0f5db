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

When a band is thinner than the halo (coarse pyramid levels of a frame cut into 8 bands), the
neighbour exchange is replaced by an all-gather of the bands.  Only the exchange is communication;
there is no collective on the per-pixel data path.
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


def _min_band_rows(x: torch.Tensor, group=None) -> int:
    m = torch.tensor([x.shape[1]], dtype=torch.int64, device=x.device)
    dist.all_reduce(m, op=dist.ReduceOp.MIN, group=group)
    return int(m.item())


def _gather_halo(x: torch.Tensor, rows: int, group=None):
    """Overflow path: some band is thinner than the halo, so a neighbour alone cannot supply it.  The
    bands are all-gathered (padded to the tallest band) and every rank cuts its own window
    [r0 - rows, r1 + rows) out of the reassembled frame, clipped at the image border -- the same
    rows a (multi-hop) neighbour exchange would deliver.  Returns (tensor, top, bottom)."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    hs = torch.zeros(world, dtype=torch.int64, device=x.device)
    hs[rank] = x.shape[1]
    dist.all_reduce(hs, group=group)
    hs = [int(v) for v in hs.tolist()]
    hmax = max(hs)
    padded = torch.nn.functional.pad(x, (0, 0, 0, 0, 0, hmax - x.shape[1])).contiguous()
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    full = torch.cat([p[:, :h] for p, h in zip(parts, hs)], dim=1)
    r0 = sum(hs[:rank])
    r1 = r0 + hs[rank]
    top, bot = min(rows, r0), min(rows, full.shape[1] - r1)
    return full[:, r0 - top:r1 + bot].contiguous(), top, bot


def exchange_halo_ex(x: torch.Tensor, rows: int, group=None, min_band_rows: int | None = None):
    """x: this rank's band (B, h, W, C).  Returns (y, top, bottom): y = (B, top + h + bottom, W, C)
    with `top`/`bottom` rows of the neighbouring bands (`rows` each, fewer at the image border).
    Bands at least `rows` tall exchange with their two neighbours (send/recv); when the thinnest band
    (`min_band_rows`; agreed on with one all-reduce if not given) is thinner than the halo, the bands
    are all-gathered instead."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if rows <= 0 or world == 1:
        return x, 0, 0
    hmin = _min_band_rows(x, group) if min_band_rows is None else min_band_rows
    if hmin < rows:
        return _gather_halo(x, rows, group)
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
    return torch.cat(parts, dim=1), (rows if "top" in bufs else 0), (rows if "bot" in bufs else 0)


def exchange_halo(x: torch.Tensor, rows: int, group=None) -> torch.Tensor:
    """`exchange_halo_ex` without the halo sizes."""
    return exchange_halo_ex(x, rows, group)[0]


def cost_volume(prv_band, nxt_band, search_range=4, leaky_slope=0.1, group=None, op=None):
    """Cost volume of this rank's row band; identical to the band of the unsharded result."""
    from . import ops as _ops
    op = op or _ops.cost_volume
    d = int(search_range)
    nxt_h, top, bot = exchange_halo_ex(nxt_band, d, group)
    # prv needs no neighbour data; pad it with zeros only to keep the two tensors the same shape
    # (rows of the halo produce outputs that are cropped away)
    prv_h = torch.nn.functional.pad(prv_band, (0, 0, 0, 0, top, bot))
    out = op(prv_h, nxt_h, d, leaky_slope)
    return out[:, top:out.shape[1] - bot].contiguous()


def warp_cost_volume(prv_band, nxt_band, flow_band, mode="tfa", search_range=4, leaky_slope=0.1,
                     group=None, op=None):
    """Fused warp -> cost volume of this rank's band.  The warp samples absolute image rows, so the
    padded band is processed with the flow expressed in band coordinates (unchanged: the flow is a
    displacement) and a halo deep enough for the largest vertical displacement."""
    from . import ops as _ops
    op = op or _ops.warp_cost_volume
    d = int(search_range)
    # one collective for both scalars: the largest vertical displacement and the thinnest band
    # (an empty band -- H < world size -- contributes 0; non-finite flow is clamped: the warp itself
    # saturates such samples at the border, so a halo of the whole image is always enough)
    fy = flow_band[..., 1].abs()
    fmax = torch.nan_to_num(fy, nan=0.0, posinf=3.0e38).max().to(torch.float32) if fy.numel() else \
        torch.zeros((), device=flow_band.device)
    m = torch.stack([fmax, torch.tensor(-float(flow_band.shape[1]), device=flow_band.device)])
    dist.all_reduce(m, op=dist.ReduceOp.MAX, group=group)
    h_total = torch.tensor(float(flow_band.shape[1]), device=flow_band.device)
    dist.all_reduce(h_total, op=dist.ReduceOp.SUM, group=group)
    rows = d + int(math.ceil(min(float(m[0].item()), float(h_total.item())))) + 1
    hmin = int(-m[1].item())
    nxt_h, top, bot = exchange_halo_ex(nxt_band, rows, group, hmin)
    flow_h, _, _ = exchange_halo_ex(flow_band, rows, group, hmin)
    prv_h = torch.nn.functional.pad(prv_band, (0, 0, 0, 0, top, bot))
    out = op(prv_h, nxt_h, flow_h, mode, d, leaky_slope)
    return out[:, top:out.shape[1] - bot].contiguous()
