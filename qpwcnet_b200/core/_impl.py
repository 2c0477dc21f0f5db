"""Layout plumbing shared by the layer classes and the functors."""
from __future__ import annotations

import torch

from .. import ops
from ..backend import get_axis, image_data_format


def resolve_format(data_format):
    fmt = image_data_format() if data_format is None else data_format
    get_axis(fmt)  # raises ValueError('Unsupported data format : ...') like the reference
    return fmt


def to_nhwc(x: torch.Tensor, fmt: str) -> torch.Tensor:
    return x.permute(0, 2, 3, 1) if fmt == "channels_first" else x


def from_nhwc(x: torch.Tensor, fmt: str) -> torch.Tensor:
    return x.permute(0, 3, 1, 2).contiguous() if fmt == "channels_first" else x


def cost_volume(inputs, search_range, fmt, leaky_slope=0.1):
    prv, nxt = inputs
    if fmt == "channels_first" and prv.is_cuda and prv.dim() == 4:
        return ops.cost_volume_nchw(prv, nxt, search_range, leaky_slope)   # native NCHW kernel
    out = ops.cost_volume(to_nhwc(prv, fmt), to_nhwc(nxt, fmt), search_range, leaky_slope)
    return from_nhwc(out, fmt)


def warp(inputs, mode, fmt):
    img, flo = inputs
    if img.dim() != 4:
        # tf_warp only defines `is_batch` for rank-4 input (qpwcnet/core/warp.py:75-80)
        raise ValueError("warp expects batched rank-4 input, got shape {}".format(tuple(img.shape)))
    if fmt == "channels_first" and img.is_cuda:
        return ops.warp_nchw(img, flo, mode)                                # native NCHW kernel
    out = ops.warp(to_nhwc(img, fmt), to_nhwc(flo, fmt), mode)
    return from_nhwc(out, fmt)


def warp_cost_volume(inputs, mode, search_range, fmt, leaky_slope=0.1):
    prv, nxt, flo = inputs
    if fmt == "channels_first" and prv.is_cuda and prv.dim() == 4:
        # native NCHW kernels end to end (warp, cost volume and both gradients): no transposes
        return ops.cost_volume_nchw(prv, ops.warp_nchw(nxt, flo, mode), search_range, leaky_slope)
    out = ops.warp_cost_volume(to_nhwc(prv, fmt), to_nhwc(nxt, fmt), to_nhwc(flo, fmt), mode,
                               search_range, leaky_slope)
    return from_nhwc(out, fmt)


def half_flow_warps(inputs, mode, fmt, flow_scale=0.5):
    prv, nxt, flo_01, flo_10 = inputs
    if fmt == "channels_first" and prv.is_cuda and prv.dim() == 4:
        # native NCHW warps with the flow scale fused; the concat along C is a plane-wise copy, no transposes
        return torch.cat([ops.warp_nchw(prv, flo_10, mode, flow_scale), ops.warp_nchw(nxt, flo_01, mode, flow_scale)], dim=1)
    out = ops.half_flow_warps(to_nhwc(prv, fmt), to_nhwc(nxt, fmt), to_nhwc(flo_01, fmt),
                              to_nhwc(flo_10, fmt), mode, flow_scale)
    return from_nhwc(out, fmt)


def upsample(x, scale, fmt):
    if fmt == "channels_first" and x.dim() == 4:
        # every (batch, channel) plane is a one-channel NHWC image: same kernel, same arithmetic, no transposes
        B, C, H, W = x.shape
        return ops.upsample2x(x.contiguous().view(B * C, H, W, 1), scale).view(B, C, 2 * H, 2 * W)
    return from_nhwc(ops.upsample2x(to_nhwc(x, fmt), scale), fmt)
