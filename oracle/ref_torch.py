"""Composition-faithful PyTorch-CPU transcription of the reference op graphs -- TEST INFRASTRUCTURE.

Second, independent restatement used to pin ``qpwc_oracle.c`` (tests/test_oracle.py) and to obtain
gradients by autograd exactly the way TF autodiff differentiates the same compositions.  It keeps
the reference's *op structure* (pad + 81 slices + mean + concat; 4 gathers + weighted add; ...)
rather than an optimised loop, so it is also what ``bench.py`` can time as "the reference's op
graph on host cores".  Never imported by the product package.

Every function takes/returns NHWC tensors.  Citations are relative to /root/reference/.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def cost_volume(prv: torch.Tensor, nxt: torch.Tensor, search_range: int = 4, slope: float = 0.1):
    """qpwcnet/core/layers.py:72-100 (CostVolume.call), channels_last branch."""
    r = search_range
    q = 2 * r + 1
    H, W = prv.shape[1], prv.shape[2]
    pad_nxt = F.pad(nxt, (0, 0, r, r, r, r))                       # ZeroPadding2D(r)   layers.py:77
    cost_vol = []
    for i0 in range(q):                                            # layers.py:80
        for j0 in range(q):                                        # layers.py:81
            roi = pad_nxt[:, i0:i0 + H, j0:j0 + W, :]              # tf.slice           layers.py:86-88
            cost_vol.append((prv * roi).mean(dim=3, keepdim=True))  # reduce_mean        layers.py:94
    cost_vol = torch.cat(cost_vol, dim=3)                          # layers.py:96
    # tf.nn.leaky_relu: grad is alpha*g for features <= 0 (features > 0 ? g : alpha*g)
    return torch.where(cost_vol > 0, cost_vol, slope * cost_vol)   # layers.py:99


def warp_tf(img: torch.Tensor, flow: torch.Tensor):
    """qpwcnet/core/warp.py:63-153 (tf_warp) + 8-47 (get_pixel_value), channels_last branch."""
    B, H, W, C = img.shape
    xs, ys = torch.meshgrid(torch.arange(W), torch.arange(H), indexing="xy")   # warp.py:88
    x = xs.to(flow.dtype)[None] + flow[..., 0]                                 # warp.py:100-111
    y = ys.to(flow.dtype)[None] + flow[..., 1]
    x0 = x.detach().to(torch.int32)          # tf.cast -> truncation toward zero   warp.py:115
    x1 = x0 + 1
    y0 = y.detach().to(torch.int32)
    y1 = y0 + 1
    x0 = x0.clamp(0, W - 1); x1 = x1.clamp(0, W - 1)                           # warp.py:121-124
    y0 = y0.clamp(0, H - 1); y1 = y1.clamp(0, H - 1)
    bidx = torch.arange(B)[:, None, None]

    def pix(yy, xx):                                                            # gather_nd  warp.py:41
        return img[bidx, yy.long(), xx.long()]

    Ia, Ib, Ic, Id = pix(y0, x0), pix(y1, x0), pix(y0, x1), pix(y1, x1)        # warp.py:127-130
    x0f, x1f, y0f, y1f = (t.to(flow.dtype) for t in (x0, x1, y0, y1))          # warp.py:133-136
    wa = (x1f - x) * (y1f - y)                                                 # warp.py:139-142
    wb = (x1f - x) * (y - y0f)
    wc = (x - x0f) * (y1f - y)
    wd = (x - x0f) * (y - y0f)
    wa, wb, wc, wd = (w[..., None] for w in (wa, wb, wc, wd))
    return ((wa * Ia + wb * Ib) + wc * Ic) + wd * Id                           # add_n      warp.py:151


def _tf_clamp01(a: torch.Tensor):
    """min(max(0, a), 1) with TF's gradient tie rule: the gradient reaches `a` iff 0 < a <= 1
    (Maximum routes ties to its first argument -- the constant 0; Minimum routes ties to its
    first argument -- `a`)."""
    one = torch.ones_like(a)
    zero = torch.zeros_like(a)
    return torch.where(a > 0, torch.where(a <= 1, a, one), zero)


def warp_tfa(img: torch.Tensor, flow: torch.Tensor):
    """WarpV2: tfa.image.dense_image_warp(img, -flo[..., ::-1]) -- qpwcnet/core/layers.py:177-186,
    with dense_image_warp/interpolate_bilinear as in the in-repo copy warp.py:156-211 (sign as tfa:
    query = grid - flow_arg)."""
    B, H, W, C = img.shape
    if H < 2 or W < 2:
        raise ValueError("Grid must be at least 2x2")
    arg = -flow.flip(-1)                                            # -flo[..., ::-1]   layers.py:185
    gx, gy = torch.meshgrid(torch.arange(W), torch.arange(H), indexing="xy")
    grid = torch.stack([gy, gx], dim=2).to(flow.dtype)[None]        # (y, x)            warp.py:198-200
    query = (grid - arg).reshape(B, H * W, 2)                       # tfa: grid - flow
    alphas, floors, ceils = [], [], []
    for dim, size in ((0, H), (1, W)):                              # indexing='ij'
        qd = query[..., dim]
        fl = torch.minimum(torch.maximum(torch.zeros_like(qd), torch.floor(qd.detach())),
                           torch.full_like(qd, float(size - 2)))
        ifl = fl.to(torch.int32)
        floors.append(ifl)
        ceils.append(ifl + 1)
        alphas.append(_tf_clamp01(qd - fl)[..., None])
    flat = img.reshape(B * H * W, C)
    boff = (torch.arange(B) * H * W)[:, None]

    def gather(yc, xc):
        return flat[(boff + yc.long() * W + xc.long())]

    tl = gather(floors[0], floors[1]); tr = gather(floors[0], ceils[1])
    bl = gather(ceils[0], floors[1]); br = gather(ceils[0], ceils[1])
    top = alphas[1] * (tr - tl) + tl
    bot = alphas[1] * (br - bl) + bl
    out = alphas[0] * (bot - top) + top
    return out.reshape(B, H, W, C)


def warp(img, flow, mode: str = "tfa"):
    return warp_tf(img, flow) if mode == "tf" else warp_tfa(img, flow)


def warp_cost_volume(prv, nxt, flow, mode: str = "tfa", search_range: int = 4, slope: float = 0.1):
    """qpwcnet/core/non_layers.py:377-380 (UpFlow): cost_volume((prv, warp((nxt, flo))))."""
    return cost_volume(prv, warp(nxt, flow, mode), search_range, slope)
