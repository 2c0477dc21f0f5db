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


# ---------------------------------------------------------------------------------------------
# Overlapped, allocation-free variant ("interior rows compute while halos are in flight")
# ---------------------------------------------------------------------------------------------
class ShardedLevel:
    """One pyramid level of a row-sharded frame pair on this rank, with every buffer allocated once.

    Layout: the band's rows live in the middle of halo-padded buffers (`R` rows above and below), so
    the neighbours' rows are received straight into their final place -- no cat / pad / crop.  `R` =
    `d` for the plain cost volume; `d + reach + 1` for the UpFlow pair, `reach` being a fixed budget for
    the vertical flow (the generic functions above handle larger flows by agreeing on the
    halo size with a collective).  Bands must be at least `R` rows tall.

    A call enqueues, without any host synchronisation:
      side stream : send the first / last R band rows to the row neighbours, receive theirs (NCCL p2p)
      main stream : the rows that need no neighbour data -- cost volume: band rows [d, h-d);
                    pair: the warp of band rows [reach+1, h-reach-1)
      main stream, after the exchange: the border strips (cost volume: rows [0,d) and [h-d,h) through a
                    3d-row scratch; pair: the two warp strips, then ONE cost-volume call over the
                    halo-padded warped band, whose band rows are all valid).
    Image-border ranks have no neighbour on one side: their halo rows stay zero (cost volume:
    ZeroPadding2D semantics).  The warp samples absolute image rows, so the pair hands the warp a view
    whose row 0 is the image's row 0 only on the first rank; on the others the flow is a displacement
    and band-local indices differ from absolute ones by a constant, which changes the fp32 rounding
    of `row + flow` by at most an ulp of the row index (same note as `warp_cost_volume` above).
    """

    def __init__(self, H, W, C, search_range=4, reach=0, pair=False, mode="tfa", group=None, device=None,
                 ops_module=None, symmetric=False):
        self.group, self.d, self.pair, self.mode = group, int(search_range), bool(pair), mode
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.r0, self.r1 = band(H, self.rank, self.world)
        self.h, self.H, self.W, self.C = self.r1 - self.r0, H, W, C
        self.reach = int(reach)
        self.R = self.d + (self.reach + 1 if pair else 0)
        if self.h < self.R and self.world > 1:
            raise ValueError(f"band of {self.h} rows is thinner than the halo ({self.R} rows): use the generic functions")
        from . import ops as _ops
        self.ops = ops_module or _ops
        dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.dev = torch.device(dev)
        R, h, D = self.R, self.h, (2 * self.d + 1) ** 2
        z = lambda *s: torch.zeros(s, device=self.dev)  # noqa: E731
        self.nxt_h = z(1, h + 2 * R, W, C)
        self.prv_h = z(1, h + 2 * self.d, W, C)        # halo rows of prv are never read for valid outputs
        self.out_h = z(1, h + 2 * self.d, W, D)
        if pair:
            self.flow_h = z(1, h + 2 * R, W, 2)
            self.nxtw_h = z(1, h + 2 * self.d, W, C)    # warped band with the d halo rows the cost volume reads
        else:
            self.strip = z(1, 3 * self.d, W, D)
        self.up, self.down = self.rank - 1, self.rank + 1
        self.side = torch.cuda.Stream(device=self.dev) if self.dev.type == "cuda" else None
        self._wtmp = None
        self._symm = None
        if symmetric and self.side is not None and self.world > 1:
            self._init_symmetric()

    def _init_symmetric(self):
        """Halo exchange as direct NVLink stores into the neighbours' buffers (torch symmetric memory: the
        haloed tensors are allocated from a peer-mapped pool, every rank sees its neighbours' copies) with
        device-side barriers -- no NCCL call, ~10 us instead of ~60 us of host-launched send/recv.  All
        ranks allocate the same (tallest-band) shape; the halo slots sit where the RECEIVER expects them."""
        import torch.distributed._symmetric_memory as symm
        grp = self.group if self.group is not None else dist.group.WORLD
        hmax = max(band(self.H, r, self.world)[1] - band(self.H, r, self.world)[0] for r in range(self.world))
        R = self.R

        def make(last):
            t = symm.empty((1, hmax + 2 * R, self.W, last), dtype=torch.float32, device=self.dev)
            t.zero_()
            return t, symm.rendezvous(t, grp)

        self._symm = {}
        full_n, hdl_n = make(self.C)
        self.nxt_h = full_n[:, :self.h + 2 * R]
        self._symm["nxt"] = (full_n, hdl_n)
        if self.pair:
            full_f, hdl_f = make(2)
            self.flow_h = full_f[:, :self.h + 2 * R]
            self._symm["flow"] = (full_f, hdl_f)
        self._peer = {}
        for name, (full, hdl) in self._symm.items():
            for nb in (self.up, self.down):
                if 0 <= nb < self.world:
                    self._peer[(name, nb)] = hdl.get_buffer(nb, tuple(full.shape), torch.float32)
        torch.cuda.synchronize()
        dist.barrier(group=self.group)

    def _exchange_symmetric(self):
        """Push my border rows into the neighbours' halo slots; two device-side barriers bracket the pushes
        (1: nobody still reads the halos of the previous pass; 2: every push has landed)."""
        R, h = self.R, self.h
        hdl0 = self._symm["nxt"][1]
        hdl0.barrier(channel=0)
        for name, (full, _) in self._symm.items():
            mine = full
            if self.up >= 0:
                h_up = band(self.H, self.up, self.world)[1] - band(self.H, self.up, self.world)[0]
                self._peer[(name, self.up)][:, R + h_up:R + h_up + R].copy_(mine[:, R:2 * R])
            if self.down < self.world:
                self._peer[(name, self.down)][:, 0:R].copy_(mine[:, h:h + R])
        hdl0.barrier(channel=1)

    # views
    @property
    def nxt(self):
        return self.nxt_h[:, self.R:self.R + self.h]

    @property
    def prv(self):
        return self.prv_h[:, self.d:self.d + self.h]

    @property
    def flow(self):
        return self.flow_h[:, self.R:self.R + self.h]

    @property
    def out(self):
        return self.out_h[:, self.d:self.d + self.h]

    def _exchange(self, bufs):
        """Post the halo sends / receives of every (haloed tensor) in `bufs`; returns the requests."""
        R, h, p2p = self.R, self.h, []
        for t in bufs:
            if self.up >= 0:
                p2p.append(dist.P2POp(dist.isend, t[:, R:2 * R], self.up, self.group))
                p2p.append(dist.P2POp(dist.irecv, t[:, 0:R], self.up, self.group))
            if self.down < self.world:
                p2p.append(dist.P2POp(dist.isend, t[:, h:h + R], self.down, self.group))
                p2p.append(dist.P2POp(dist.irecv, t[:, h + R:h + 2 * R], self.down, self.group))
        return dist.batch_isend_irecv(p2p) if p2p else []

    def run(self):
        """Enqueue one pass over the level; returns the band's cost volume (a view of `out_h`)."""
        o, d, R, h = self.ops, self.d, self.R, self.h
        cuda = self.side is not None
        bufs = [self.nxt_h] + ([self.flow_h] if self.pair else [])
        if cuda:
            self.side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.side):
                if self._symm is not None:
                    self._exchange_symmetric()
                else:
                    reqs = self._exchange(bufs)
                    for q in reqs:
                        q.wait()                  # stream-level wait on the NCCL stream, no host block
        else:
            reqs = self._exchange(bufs)
        if not self.pair:
            # interior: band-only data is enough for rows [d, h-d) -- the call also writes rows [0,d) and
            # [h-d,h) (computed against zero padding), which the strips below overwrite
            o.cost_volume_into(self.out, self.prv, self.nxt, d)
            if cuda:
                torch.cuda.current_stream().wait_stream(self.side)
            else:
                for q in reqs:
                    q.wait()
            if self.world > 1:
                for top in (True, False):
                    if (top and self.up < 0) or (not top and self.down >= self.world):
                        continue              # image border: zero padding is the right answer already
                    a = 0 if top else h - d   # band rows [a, a+d) from the 3d rows around them
                    o.cost_volume_into(self.strip, self.prv_h[:, a:a + 3 * d], self.nxt_h[:, a + R - d:a + R + 2 * d], d)
                    self.out[:, a:a + d].copy_(self.strip[:, d:2 * d])
            return self.out
        # ---- UpFlow pair: warp rows [-d, h+d) of the band into nxtw_h, then one cost volume over it
        k = self.reach + 1
        lo, hi = (k if self.up >= 0 else 0), (h - k if self.down < self.world else h)
        hi = max(hi, lo)
        if hi > lo:   # interior rows: their taps stay inside the band (|flow_y| <= reach)
            self._warp_rows(lo, hi)
        if cuda:
            torch.cuda.current_stream().wait_stream(self.side)
        else:
            for q in reqs:
                q.wait()
        top_lo = -d if self.up >= 0 else 0
        bot_hi = h + d if self.down < self.world else h
        if lo > top_lo:
            self._warp_rows(top_lo, lo)
        if bot_hi > hi:
            self._warp_rows(hi, bot_hi)
        o.cost_volume_into(self.out_h, self.prv_h, self.nxtw_h, d)
        return self.out

    def _warp_rows(self, a, b):
        """Warped band rows [a, b) (band coordinates, may reach d rows into the halo) -> nxtw_h.  The warp
        runs on the window of haloed rows that holds every tap of those rows (`reach + 1` rows around
        them) with ABSOLUTE row arithmetic (`warp_rows_into`: bit-identical to the unsharded warp), and
        rows [a, b) of its result are kept."""
        d, R, k = self.d, self.R, self.reach + 1
        s0, s1 = max(a + R - k, 0), min(b + R + k, self.h + 2 * R)          # source window in haloed coordinates
        if self.up < 0:
            s0 = max(s0, R)          # first rank: no rows above the image
        if self.down >= self.world:
            s1 = min(s1, R + self.h)
        img, flo = self.nxt_h[:, s0:s1], self.flow_h[:, s0:s1]
        if hasattr(self.ops, "warp_rows_into"):
            if self._wtmp is None or self._wtmp.shape[1] < s1 - s0:
                self._wtmp = torch.empty((1, self.h + 2 * R, self.W, self.C), device=self.dev)
            tmp = self._wtmp[:, :s1 - s0]
            self.ops.warp_rows_into(tmp, img, flo, self.mode, self.r0 - R + s0, self.H)
        else:                        # CPU stand-in (tests): view-local row arithmetic, equal within fp32 rounding
            tmp = self.ops.warp(img, flo, self.mode)
        self.nxtw_h[:, d + a:d + b].copy_(tmp[:, a + R - s0:b + R - s0])

    # ---- CUDA graph of one pass.  Only with the symmetric-memory exchange (plain kernels, copies and
    # device-side barriers); the NCCL p2p variant must not be captured (it deadlocked under capture).
    def capture(self):
        if self._symm is None:
            raise RuntimeError("capture() needs the symmetric-memory exchange (symmetric=True)")
        for _ in range(2):
            self.run()
        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self.run()
        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        return self

    def replay(self):
        self._graph.replay()
        return self.out
