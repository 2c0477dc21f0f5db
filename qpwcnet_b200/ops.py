"""Functional ops over libqpwc: cost volume, warp, fused warp->cost volume (NHWC, fp32).

``torch.autograd.Function``s whose forward/backward hand DLPack-exported pointers to the C ABI on
torch's current CUDA stream.  CPU (host) tensors are routed to the library's host-buffer entry
points (H2D -> kernels -> D2H, pipelined inside the library); they are inference-only.  Nothing
here computes on the CPU.

Reference semantics (citations relative to the reference checkout):
  cost_volume       qpwcnet/core/layers.py:72-100, 117-132
  warp mode 'tf'    qpwcnet/core/warp.py:63-153        mode 'tfa'  qpwcnet/core/layers.py:171-186
  warp_cost_volume  qpwcnet/core/non_layers.py:377-380
"""
from __future__ import annotations

import torch

from . import _cabi
from ._cabi import WARP_MODES, check, dlptr, dlview, kDLCUDA, kDLCUDAManaged, lib


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


class _on_device:
    """`with _on_device(dev):` -- make `dev` current for the C call; a no-op (no context push) in
    the common case where it already is."""
    __slots__ = ("idx", "prev")

    def __init__(self, device):
        self.idx = device.index if device.index is not None else torch.cuda.current_device()

    def __enter__(self):
        self.prev = torch.cuda.current_device()
        if self.prev != self.idx:
            torch.cuda.set_device(self.idx)

    def __exit__(self, *exc):
        if self.prev != self.idx:
            torch.cuda.set_device(self.prev)
        return False


def _prep(t: torch.Tensor, name: str, last: int | None = None) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name}: expected a torch.Tensor, got {type(t).__name__}")
    if t.dim() != 4:
        raise ValueError(f"{name}: expected a rank-4 NHWC tensor, got shape {tuple(t.shape)}")
    if t.dtype != torch.float32:
        raise TypeError(f"{name}: expected float32, got {t.dtype}")
    if last is not None and t.shape[-1] != last:
        raise ValueError(f"{name}: last dimension must be {last}, got shape {tuple(t.shape)}")
    return t.contiguous()


def _same(a: torch.Tensor, b: torch.Tensor, what: str):
    if a.shape != b.shape:
        raise ValueError(f"{what}: shape mismatch {tuple(a.shape)} vs {tuple(b.shape)}")
    if a.device != b.device:
        raise ValueError(f"{what}: device mismatch {a.device} vs {b.device}")


class _V:
    __slots__ = ("ptr", "on_cuda")

    def __init__(self, t):
        self.ptr, devt = dlptr(t)
        self.on_cuda = devt == kDLCUDA or devt == kDLCUDAManaged


def _views(*ts):
    # every tensor reaching this point went through _prep()/torch.empty(): dense by construction;
    # pointer and device come out of the tensor's DLPack export
    return [_V(t) for t in ts]


def _mode(mode) -> int:
    if mode in WARP_MODES:
        return WARP_MODES[mode]
    if mode in (0, 1):
        return int(mode)
    raise ValueError(f"warp mode must be 'tf' or 'tfa', got {mode!r}")


# ------------------------------------------------------------------------------- raw launchers
def _corr_fwd(prv, nxt, d, slope, out=None, out_stride=None):
    B, H, W, C = prv.shape
    D = (2 * d + 1) ** 2
    ops = D if out_stride is None else int(out_stride)
    if out is None:
        out = torch.empty((B, H, W, ops), dtype=torch.float32, device=prv.device)
    vp, vn, vo = _views(prv, nxt, out)
    if vp.on_cuda:
        with _on_device(prv.device):
            check(lib().qpwc_corr_fwd(vp.ptr, vn.ptr, vo.ptr, B, H, W, C, d, slope, ops,
                                      _stream_ptr(prv.device)))
    else:
        if ops != D:
            raise ValueError("strided output is only available on the device path")
        check(lib().qpwc_corr_fwd_host(vp.ptr, vn.ptr, vo.ptr, B, H, W, C, d, slope,
                                       torch.cuda.current_device()))
    return out


def _corr_bwd(prv, nxt, out, g_out, d, slope):
    B, H, W, C = prv.shape
    ops = out.shape[-1]
    g_prv = torch.empty_like(prv)
    g_nxt = torch.empty_like(nxt)
    vp, vn, vo, vg, vgp, vgn = _views(prv, nxt, out, g_out, g_prv, g_nxt)
    with _on_device(prv.device):
        check(lib().qpwc_corr_bwd(vp.ptr, vn.ptr, vo.ptr, vg.ptr, vgp.ptr, vgn.ptr, B, H, W, C, d,
                                  slope, ops, _stream_ptr(prv.device)))
    return g_prv, g_nxt


def _warp_fwd(img, flow, mode, out=None):
    B, H, W, C = img.shape
    if out is None:
        out = torch.empty_like(img)
    vi, vf, vo = _views(img, flow, out)
    if vi.on_cuda:
        with _on_device(img.device):
            check(lib().qpwc_warp_fwd(vi.ptr, vf.ptr, vo.ptr, B, H, W, C, mode,
                                      _stream_ptr(img.device)))
    else:
        check(lib().qpwc_warp_fwd_host(vi.ptr, vf.ptr, vo.ptr, B, H, W, C, mode,
                                       torch.cuda.current_device()))
    return out


def _warp_bwd(img, flow, g_out, mode):
    B, H, W, C = img.shape
    g_img = torch.empty_like(img)
    g_flow = torch.empty_like(flow)
    vi, vf, vg, vgi, vgf = _views(img, flow, g_out, g_img, g_flow)
    with _on_device(img.device):
        check(lib().qpwc_warp_bwd(vi.ptr, vf.ptr, vg.ptr, vgi.ptr, vgf.ptr, B, H, W, C, mode,
                                  _stream_ptr(img.device)))
    return g_img, g_flow


def _warp_fwd_scaled(img, flow, mode, scale, out=None, out_view=None):
    """warp(img, scale * flow) -> out (dense, or `out_view`: a channel slice [.., o:o+C] of a wider
    contiguous NHWC buffer, written in place)."""
    B, H, W, C = img.shape
    if out_view is None:
        out = torch.empty_like(img) if out is None else out
        target, stride = out, C
    else:
        target, stride = out_view, out_view.stride(2)
    vi, vf = _views(img, flow)
    with _on_device(img.device):
        check(lib().qpwc_warp_fwd_ex(vi.ptr, vf.ptr, target.data_ptr(), B, H, W, C, mode, float(scale),
                                     stride, _stream_ptr(img.device)))
    return target


def _warp_bwd_scaled(img, flow, g_out, mode, scale):
    """g_out may be a channel slice of a wider contiguous NHWC gradient buffer."""
    B, H, W, C = img.shape
    g_img = torch.empty_like(img)
    g_flow = torch.empty_like(flow)
    vi, vf, vgi, vgf = _views(img, flow, g_img, g_flow)
    with _on_device(img.device):
        check(lib().qpwc_warp_bwd_ex(vi.ptr, vf.ptr, g_out.data_ptr(), vgi.ptr, vgf.ptr, B, H, W, C, mode,
                                     float(scale), g_out.stride(2), _stream_ptr(img.device)))
    return g_img, g_flow


def _slice_ok(v, C):
    """v: (B,H,W,C) view whose pixels are C contiguous floats at a constant pixel stride."""
    return (v.stride(3) == 1 and v.stride(1) == v.stride(2) * v.shape[2]
            and v.stride(0) == v.stride(1) * v.shape[1] and v.stride(2) >= C)


def _warp_corr_fwd(prv, nxt, flow, mode, d, slope, out=None, out_stride=None):
    B, H, W, C = prv.shape
    D = (2 * d + 1) ** 2
    ops = D if out_stride is None else int(out_stride)
    if out is None:
        out = torch.empty((B, H, W, ops), dtype=torch.float32, device=prv.device)
    vp, vn, vf, vo = _views(prv, nxt, flow, out)
    if vp.on_cuda:
        with _on_device(prv.device):
            check(lib().qpwc_warp_corr_fwd(vp.ptr, vn.ptr, vf.ptr, vo.ptr, B, H, W, C, d, slope,
                                           mode, ops, _stream_ptr(prv.device)))
    else:
        if ops != D:
            raise ValueError("strided output is only available on the device path")
        check(lib().qpwc_warp_corr_fwd_host(vp.ptr, vn.ptr, vf.ptr, vo.ptr, B, H, W, C, d, slope,
                                            mode, torch.cuda.current_device()))
    return out


def _warp_corr_bwd(prv, nxt, flow, out, g_out, mode, d, slope):
    B, H, W, C = prv.shape
    ops = out.shape[-1]
    g_prv, g_nxt, g_flow = torch.empty_like(prv), torch.empty_like(nxt), torch.empty_like(flow)
    # no workspace (library >= 0.2): the call owns an L2-resident scratch for the warped frame and its gradient
    vp, vn, vf, vo, vg, vgp, vgn, vgf = _views(prv, nxt, flow, out, g_out, g_prv, g_nxt, g_flow)
    with _on_device(prv.device):
        check(lib().qpwc_warp_corr_bwd(vp.ptr, vn.ptr, vf.ptr, vo.ptr, vg.ptr, vgp.ptr, vgn.ptr,
                                       vgf.ptr, None, 0, B, H, W, C, d, slope, mode,
                                       ops, _stream_ptr(prv.device)))
    return g_prv, g_nxt, g_flow


def _no_host_grad(*ts):
    if any(t.requires_grad and not t.is_cuda for t in ts) and torch.is_grad_enabled():
        raise RuntimeError("host (CPU) tensors take the inference-only staged path of libqpwc; "
                           "move them to the GPU to differentiate")


# -------------------------------------------------------------------------------------- autograd
class _CostVolume(torch.autograd.Function):
    @staticmethod
    def forward(ctx, prv, nxt, d, slope):
        out = _corr_fwd(prv, nxt, d, slope)
        ctx.save_for_backward(prv, nxt, out)
        ctx.cfg = (d, slope)
        return out

    @staticmethod
    def backward(ctx, g_out):
        prv, nxt, out = ctx.saved_tensors
        d, slope = ctx.cfg
        g_prv, g_nxt = _corr_bwd(prv, nxt, out, g_out.contiguous(), d, slope)
        return g_prv, g_nxt, None, None


def _corr_fwd_nchw(prv, nxt, d, slope):
    """Native channels_first forward (tiled kernel for d == 4, W % 4 == 0; shape-generic NCHW kernel elsewhere)."""
    B, C, H, W = prv.shape
    out = torch.empty((B, (2 * d + 1) ** 2, H, W), dtype=torch.float32, device=prv.device)
    if out.numel() == 0:
        return out
    vp, vn = _views(prv, nxt)
    with _on_device(prv.device):
        check(lib().qpwc_corr_fwd_nchw(vp.ptr, vn.ptr, out.data_ptr(), B, C, H, W, d, slope, _stream_ptr(prv.device)))
    return out


class _CostVolumeNCHW(torch.autograd.Function):
    """channels_first cost volume: native NCHW forward and gradient kernels at every shape -- no layout
    transposes (qpwcnet/core/layers.py:83-85 transposes around a channels_last op instead)."""

    @staticmethod
    def forward(ctx, prv, nxt, d, slope):
        out = _corr_fwd_nchw(prv, nxt, d, slope)
        ctx.save_for_backward(prv, nxt, out)
        ctx.cfg = (d, slope)
        return out

    @staticmethod
    def backward(ctx, g_out):
        prv, nxt, out = ctx.saved_tensors
        d, slope = ctx.cfg
        B, C, H, W = prv.shape
        g_out = g_out.contiguous()
        g_prv, g_nxt = torch.empty_like(prv), torch.empty_like(nxt)
        if prv.numel():
            with _on_device(prv.device):
                check(lib().qpwc_corr_bwd_nchw(prv.data_ptr(), nxt.data_ptr(), out.data_ptr(), g_out.data_ptr(),
                                               g_prv.data_ptr(), g_nxt.data_ptr(), B, C, H, W, d, slope,
                                               _stream_ptr(prv.device)))
        return g_prv, g_nxt, None, None


class _WarpNCHW(torch.autograd.Function):
    """channels_first warp: native NCHW forward and backward kernels."""

    @staticmethod
    def forward(ctx, img, flow, mode, scale=1.0):
        B, C, H, W = img.shape
        out = torch.empty_like(img)
        vi, vf = _views(img, flow)
        with _on_device(img.device):
            check(lib().qpwc_warp_fwd_nchw(vi.ptr, vf.ptr, out.data_ptr(), B, C, H, W, mode, float(scale),
                                           _stream_ptr(img.device)))
        ctx.save_for_backward(img, flow)
        ctx.mode, ctx.scale = mode, float(scale)
        return out

    @staticmethod
    def backward(ctx, g_out):
        img, flow = ctx.saved_tensors
        B, C, H, W = img.shape
        g_out = g_out.contiguous()
        g_img, g_flow = torch.empty_like(img), torch.empty_like(flow)
        with _on_device(img.device):
            check(lib().qpwc_warp_bwd_nchw(img.data_ptr(), flow.data_ptr(), g_out.data_ptr(), g_img.data_ptr(),
                                           g_flow.data_ptr(), B, C, H, W, ctx.mode, ctx.scale, _stream_ptr(img.device)))
        return g_img, g_flow, None, None


class _Warp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img, flow, mode):
        out = _warp_fwd(img, flow, mode)
        ctx.save_for_backward(img, flow)
        ctx.mode = mode
        return out

    @staticmethod
    def backward(ctx, g_out):
        img, flow = ctx.saved_tensors
        g_img, g_flow = _warp_bwd(img, flow, g_out.contiguous(), ctx.mode)
        return g_img, g_flow, None


class _WarpScaled(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img, flow, mode, scale):
        out = _warp_fwd_scaled(img, flow, mode, scale)
        ctx.save_for_backward(img, flow)
        ctx.cfg = (mode, scale)
        return out

    @staticmethod
    def backward(ctx, g_out):
        img, flow = ctx.saved_tensors
        mode, scale = ctx.cfg
        if not _slice_ok(g_out, img.shape[3]):
            g_out = g_out.contiguous()
        g_img, g_flow = _warp_bwd_scaled(img, flow, g_out, mode, scale)
        return g_img, g_flow, None, None


class _WarpPair(torch.autograd.Function):
    """FrameInterpolate's two half-flow warps as one forward launch into one (B,H,W,2C) buffer."""

    @staticmethod
    def forward(ctx, img_a, flow_a, img_b, flow_b, mode, scale):
        B, H, W, C = img_a.shape
        out = torch.empty((B, H, W, 2 * C), dtype=torch.float32, device=img_a.device)
        _warp_pair_fwd(out, img_a, flow_a, img_b, flow_b, mode, scale)
        ctx.save_for_backward(img_a, flow_a, img_b, flow_b)
        ctx.cfg = (mode, scale)
        return out

    @staticmethod
    def backward(ctx, g_out):
        img_a, flow_a, img_b, flow_b = ctx.saved_tensors
        mode, scale = ctx.cfg
        C = img_a.shape[3]
        g_out = g_out.contiguous()
        ga, gfa = _warp_bwd_scaled(img_a, flow_a, g_out[..., :C], mode, scale)   # slices, no copy
        gb, gfb = _warp_bwd_scaled(img_b, flow_b, g_out[..., C:], mode, scale)
        return ga, gfa, gb, gfb, None, None


def _warp_pair_fwd(out, img_a, flow_a, img_b, flow_b, mode, scale):
    B, H, W, C = img_a.shape
    va, vfa, vb, vfb = _views(img_a, flow_a, img_b, flow_b)
    with _on_device(img_a.device):
        check(lib().qpwc_warp_pair_fwd(va.ptr, vfa.ptr, vb.ptr, vfb.ptr, out.data_ptr(), B, H, W, C, mode,
                                       float(scale), out.stride(2), _stream_ptr(img_a.device)))
    return out


def _upsample2x_fwd(x, scale):
    B, H, W, C = x.shape
    out = torch.empty((B, 2 * H, 2 * W, C), dtype=torch.float32, device=x.device)
    (vx,) = _views(x)
    with _on_device(x.device):
        check(lib().qpwc_upsample2x_fwd(vx.ptr, out.data_ptr(), B, H, W, C, float(scale), _stream_ptr(x.device)))
    return out


def _upsample2x_bwd(g_out, scale):
    B, H2, W2, C = g_out.shape
    g = torch.empty((B, H2 // 2, W2 // 2, C), dtype=torch.float32, device=g_out.device)
    (vg,) = _views(g_out)
    with _on_device(g_out.device):
        check(lib().qpwc_upsample2x_bwd(vg.ptr, g.data_ptr(), B, H2 // 2, W2 // 2, C, float(scale),
                                        _stream_ptr(g_out.device)))
    return g


class _Upsample2x(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, scale):
        ctx.scale = scale
        return _upsample2x_fwd(x, scale)

    @staticmethod
    def backward(ctx, g_out):
        return _upsample2x_bwd(g_out.contiguous(), ctx.scale), None


class _WarpUp(torch.autograd.Function):
    """warp(img, up_scale * bilinear_x2(flow_coarse)) with the upsampling done inside the warp kernel."""

    @staticmethod
    def forward(ctx, img, flow_c, mode, up_scale):
        B, H, W, C = img.shape
        out = torch.empty_like(img)
        vi, vf = _views(img, flow_c)
        with _on_device(img.device):
            check(lib().qpwc_warp_fwd_up(vi.ptr, vf.ptr, out.data_ptr(), B, H, W, C, mode, float(up_scale),
                                         _stream_ptr(img.device)))
        ctx.save_for_backward(img, flow_c)
        ctx.cfg = (mode, up_scale)
        return out

    @staticmethod
    def backward(ctx, g_out):
        img, flow_c = ctx.saved_tensors
        mode, up_scale = ctx.cfg
        flow = _upsample2x_fwd(flow_c, up_scale)                 # re-materialised for the backward only
        g_img, g_flow = _warp_bwd(img, flow, g_out.contiguous(), mode)
        return g_img, _upsample2x_bwd(g_flow, up_scale), None, None


class _WarpCostVolumeUp(torch.autograd.Function):
    """Fused UpFlow pair on the coarse flow: cost_volume(prv, warp(nxt, up_scale * bilinear_x2(flow_c)))."""

    @staticmethod
    def forward(ctx, prv, nxt, flow_c, mode, d, slope, up_scale):
        B, H, W, C = prv.shape
        D = (2 * d + 1) ** 2
        out = torch.empty((B, H, W, D), dtype=torch.float32, device=prv.device)
        vp, vn, vf = _views(prv, nxt, flow_c)
        with _on_device(prv.device):
            check(lib().qpwc_warp_corr_fwd_up(vp.ptr, vn.ptr, vf.ptr, out.data_ptr(), B, H, W, C, d, slope, mode,
                                              D, float(up_scale), _stream_ptr(prv.device)))
        ctx.save_for_backward(prv, nxt, flow_c, out)
        ctx.cfg = (mode, d, slope, up_scale)
        return out

    @staticmethod
    def backward(ctx, g_out):
        prv, nxt, flow_c, out = ctx.saved_tensors
        mode, d, slope, up_scale = ctx.cfg
        flow = _upsample2x_fwd(flow_c, up_scale)
        g = _warp_corr_bwd(prv, nxt, flow, out, g_out.contiguous(), mode, d, slope)
        return g[0], g[1], _upsample2x_bwd(g[2], up_scale), None, None, None, None


class _WarpCostVolume(torch.autograd.Function):
    @staticmethod
    def forward(ctx, prv, nxt, flow, mode, d, slope):
        out = _warp_corr_fwd(prv, nxt, flow, mode, d, slope)
        ctx.save_for_backward(prv, nxt, flow, out)
        ctx.cfg = (mode, d, slope)
        return out

    @staticmethod
    def backward(ctx, g_out):
        prv, nxt, flow, out = ctx.saved_tensors
        mode, d, slope = ctx.cfg
        g = _warp_corr_bwd(prv, nxt, flow, out, g_out.contiguous(), mode, d, slope)
        return g[0], g[1], g[2], None, None, None


# ------------------------------------------------------------------------------------ public API
def cost_volume(prv, nxt, search_range: int = 4, leaky_slope: float = 0.1):
    """``leaky_relu(mean_c(prv * shift(nxt, di, dj)))`` for all |di|,|dj| <= search_range; NHWC in,
    ``(B,H,W,(2d+1)^2)`` out, channel = (di+d)*(2d+1)+(dj+d)."""
    prv, nxt = _prep(prv, "prv"), _prep(nxt, "nxt")
    _same(prv, nxt, "cost_volume(prv, nxt)")
    _no_host_grad(prv, nxt)
    if not prv.is_cuda:
        return _corr_fwd(prv, nxt, int(search_range), float(leaky_slope))
    return _CostVolume.apply(prv, nxt, int(search_range), float(leaky_slope))


def cost_volume_nchw(prv, nxt, search_range: int = 4, leaky_slope: float = 0.1):
    """``cost_volume`` for channels_first tensors: (B,C,H,W) x2 -> (B,(2d+1)^2,H,W), through the
    native NCHW kernel (qpwc_corr_nchw.cu) when the shape allows it (d == 4, W % 4 == 0)."""
    if prv.dim() != 4 or prv.shape != nxt.shape or prv.device != nxt.device:
        raise ValueError(f"cost_volume_nchw(prv, nxt): {tuple(prv.shape)}@{prv.device} vs {tuple(nxt.shape)}@{nxt.device}")
    if prv.dtype != torch.float32 or nxt.dtype != torch.float32:
        raise TypeError("cost_volume_nchw expects float32 tensors")
    if not prv.is_cuda:
        raise ValueError("cost_volume_nchw needs CUDA tensors")
    return _CostVolumeNCHW.apply(prv.contiguous(), nxt.contiguous(), int(search_range), float(leaky_slope))


def warp_nchw(img, flow, mode="tfa", flow_scale: float = 1.0):
    """``warp`` for channels_first tensors: img (B,C,H,W), flow (B,2,H,W) (plane 0 = x) -> (B,C,H,W),
    through the native NCHW kernel; ``flow_scale`` fuses FrameInterpolate's ``0.5 * flo``."""
    if img.dim() != 4 or flow.dim() != 4 or flow.shape[1] != 2 or img.shape[0] != flow.shape[0] \
            or img.shape[2:] != flow.shape[2:] or img.device != flow.device:
        raise ValueError(f"warp_nchw: img {tuple(img.shape)}@{img.device} vs flow {tuple(flow.shape)}@{flow.device}")
    if img.dtype != torch.float32 or flow.dtype != torch.float32:
        raise TypeError("warp_nchw expects float32 tensors")
    if not img.is_cuda:
        raise ValueError("warp_nchw needs CUDA tensors")
    m = _mode(mode)
    if m == 1 and (img.shape[2] < 2 or img.shape[3] < 2):
        raise ValueError("Grid must be at least 2x2 (tfa interpolate_bilinear)")
    return _WarpNCHW.apply(img.contiguous(), flow.contiguous(), m, float(flow_scale))


def warp(img, flow, mode="tfa", flow_scale: float = 1.0):
    """Bilinear backward warp: ``out[b,i,j] = img[b, i+s*flow[...,1], j+s*flow[...,0]]``; ``mode``
    picks the reference's border rule ('tf' = Warp/tf_warp, 'tfa' = WarpV2).  ``flow_scale`` s != 1
    fuses the reference's ``warp((img, 0.5 * flo))`` (FrameInterpolate, non_layers.py:303-304)."""
    img = _prep(img, "img")
    flow = _prep(flow, "flow", last=2)
    m = _mode(mode)
    _check_flow("warp", img, flow, m)
    _no_host_grad(img, flow)
    if float(flow_scale) != 1.0:
        if not img.is_cuda:
            raise ValueError("warp(flow_scale != 1) needs CUDA tensors")
        return _WarpScaled.apply(img, flow, m, float(flow_scale))
    if not img.is_cuda:
        return _warp_fwd(img, flow, m)
    return _Warp.apply(img, flow, m)


def upsample2x(x, scale: float = 1.0):
    """``Upsample(scale)`` of the reference (non_layers.py:183-193): ``scale *
    UpSampling2D(interpolation='bilinear')(x)`` -- tf.image.resize bilinear with half-pixel
    centres, x2.  NHWC (B,H,W,C) -> (B,2H,2W,C).  Differentiable."""
    x = _prep(x, "x")
    if not x.is_cuda:
        raise ValueError("upsample2x needs a CUDA tensor")
    return _Upsample2x.apply(x, float(scale))


def occlusion_map(flow, data_format: str = "channels_last"):
    """``estimate_occlusion_map(flow, data_format)`` (qpwcnet/core/occlusion.py:27-118): (B,H,W)
    float map, 1 where the next frame has no value -- the flow target leaves the image, or no
    pixel's naive inverse flow ``-tf_warp(flow, flow)`` lands there.  ``flow``: (B,H,W,2), or
    (B,2,H,W) for 'channels_first'; channel 0 = dx, 1 = dy.  Not differentiable (integer scatter)."""
    if not isinstance(flow, torch.Tensor) or flow.dim() != 4:
        raise ValueError("occlusion_map expects a batched rank-4 flow")
    if flow.dtype != torch.float32:
        raise TypeError("occlusion_map expects a float32 flow")
    if not flow.is_cuda:
        raise ValueError("occlusion_map needs a CUDA tensor")
    cf = data_format == "channels_first"
    if flow.shape[1 if cf else 3] != 2:
        raise ValueError(f"occlusion_map: flow {tuple(flow.shape)} needs 2 channels ({data_format})")
    flow = flow.detach().contiguous()
    B = flow.shape[0]
    H, W = (flow.shape[2], flow.shape[3]) if cf else (flow.shape[1], flow.shape[2])
    out = torch.empty((B, H, W), dtype=torch.float32, device=flow.device)
    with _on_device(flow.device):
        check(lib().qpwc_occlusion_map(flow.data_ptr(), out.data_ptr(), B, H, W, int(cf), _stream_ptr(flow.device)))
    return out


def _check_coarse(img, flow_c, what):
    B, H, W, _ = img.shape
    if H % 2 or W % 2 or tuple(flow_c.shape) != (B, H // 2, W // 2, 2) or flow_c.device != img.device:
        raise ValueError(f"{what}: features {tuple(img.shape)}@{img.device} need a coarse flow "
                         f"({B}, {H // 2}, {W // 2}, 2) on the same device (H, W even), got {tuple(flow_c.shape)}@{flow_c.device}")
    if not img.is_cuda:
        raise ValueError(f"{what} needs CUDA tensors")


def warp_up(img, flow_coarse, mode="tfa", up_scale: float = 2.0):
    """``warp((img, Upsample(up_scale)(flow_coarse)))`` with the x2 bilinear flow upsampling
    (pwcnet.py:49-56) interpolated inside the warp kernel: the upsampled flow is never read back
    from HBM.  ``flow_coarse``: (B, H/2, W/2, 2).  Differentiable."""
    img = _prep(img, "img")
    flow_coarse = _prep(flow_coarse, "flow_coarse", last=2)
    _check_coarse(img, flow_coarse, "warp_up")
    m = _mode(mode)
    if m == 1 and (img.shape[1] < 2 or img.shape[2] < 2):
        raise ValueError("Grid must be at least 2x2 (tfa interpolate_bilinear)")
    return _WarpUp.apply(img, flow_coarse, m, float(up_scale))


def warp_cost_volume_up(prv, nxt, flow_coarse, mode="tfa", search_range: int = 4, leaky_slope: float = 0.1,
                        up_scale: float = 2.0):
    """Fused ``cost_volume(prv, warp(nxt, Upsample(up_scale)(flow_coarse)))``: UpFlow on the coarse
    flow, one kernel; neither the upsampled flow nor the warped frame is read back from HBM."""
    prv, nxt = _prep(prv, "prv"), _prep(nxt, "nxt")
    flow_coarse = _prep(flow_coarse, "flow_coarse", last=2)
    _same(prv, nxt, "warp_cost_volume_up(prv, nxt)")
    _check_coarse(prv, flow_coarse, "warp_cost_volume_up")
    m = _mode(mode)
    if m == 1 and (prv.shape[1] < 2 or prv.shape[2] < 2):
        raise ValueError("Grid must be at least 2x2 (tfa interpolate_bilinear)")
    return _WarpCostVolumeUp.apply(prv, nxt, flow_coarse, m, int(search_range), float(leaky_slope), float(up_scale))


def _pair_args(prv, nxt, flo_01, flo_10, mode):
    prv, nxt = _prep(prv, "prv"), _prep(nxt, "nxt")
    flo_01, flo_10 = _prep(flo_01, "flo_01", last=2), _prep(flo_10, "flo_10", last=2)
    _same(prv, nxt, "half_flow_warps(prv, nxt)")
    for f in (flo_01, flo_10):
        if f.shape[:3] != prv.shape[:3] or f.device != prv.device:
            raise ValueError(f"half_flow_warps: flow {tuple(f.shape)}@{f.device} vs features {tuple(prv.shape)}@{prv.device}")
    if not prv.is_cuda:
        raise ValueError("half_flow_warps needs CUDA tensors")
    m = _mode(mode)
    if m == 1 and (prv.shape[1] < 2 or prv.shape[2] < 2):
        raise ValueError("Grid must be at least 2x2 (tfa interpolate_bilinear)")
    return prv, nxt, flo_01, flo_10, m


def half_flow_warps(prv, nxt, flo_01, flo_10, mode="tfa", flow_scale: float = 0.5):
    """FrameInterpolate's warp pair (qpwcnet/core/non_layers.py:303-311) in one launch: returns the
    (B,H,W,2C) tensor ``concat([warp(prv, s*flo_10), warp(nxt, s*flo_01)], -1)`` -- the reference's
    ``[prv_w, nxt_w]`` -- without materialising ``s*flo``.  Differentiable."""
    prv, nxt, flo_01, flo_10, m = _pair_args(prv, nxt, flo_01, flo_10, mode)
    return _WarpPair.apply(prv, flo_10, nxt, flo_01, m, float(flow_scale))


def half_flow_warps_into(out, prv, nxt, flo_01, flo_10, mode="tfa", flow_scale: float = 0.5):
    """Inference-only: the pair lands in channels [0, 2C) of every pixel of a caller-owned
    contiguous (B,H,W,S >= 2C) buffer (the concat input of FrameInterpolate's convolutions)."""
    prv, nxt, flo_01, flo_10, m = _pair_args(prv, nxt, flo_01, flo_10, mode)
    C = prv.shape[3]
    if (out.dim() != 4 or out.shape[:3] != prv.shape[:3] or out.shape[3] < 2 * C or out.dtype != torch.float32
            or out.device != prv.device or not out.is_contiguous()):
        raise ValueError(f"half_flow_warps_into: out must be a contiguous float32 (B,H,W,S>=2C) tensor on {prv.device}")
    return _warp_pair_fwd(out, prv, flo_10, nxt, flo_01, m, float(flow_scale))


def _check_flow(what, ref, flow, m):
    """Shape / device / mode checks shared by the autograd entry points and their `_into` twins: the
    C ABI only sees B, H, W of the image, so a smaller or other-device flow would be read out of
    bounds."""
    if ref.shape[:3] != flow.shape[:3] or ref.device != flow.device:
        raise ValueError(f"{what}: image {tuple(ref.shape)} @ {ref.device} vs flow {tuple(flow.shape)} @ {flow.device}")
    if m == 1 and (ref.shape[1] < 2 or ref.shape[2] < 2):
        raise ValueError("Grid must be at least 2x2 (tfa interpolate_bilinear)")


def _no_grad_into(what, *ts):
    """The `_into` entry points are inference-only: refuse silently dropping a graph."""
    if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in ts):
        raise RuntimeError(f"{what} is inference-only (no autograd graph is recorded); call it under "
                           "torch.no_grad() / with detached inputs, or use the differentiable twin")


def warp_cost_volume(prv, nxt, flow, mode="tfa", search_range: int = 4, leaky_slope: float = 0.1):
    """Fused ``cost_volume(prv, warp(nxt, flow))`` (UpFlow): one kernel, the warped frame never
    reaches HBM."""
    prv, nxt = _prep(prv, "prv"), _prep(nxt, "nxt")
    flow = _prep(flow, "flow", last=2)
    _same(prv, nxt, "warp_cost_volume(prv, nxt)")
    m = _mode(mode)
    _check_flow("warp_cost_volume", prv, flow, m)
    _no_host_grad(prv, nxt, flow)
    if not prv.is_cuda:
        return _warp_corr_fwd(prv, nxt, flow, m, int(search_range), float(leaky_slope))
    return _WarpCostVolume.apply(prv, nxt, flow, m, int(search_range), float(leaky_slope))


def cost_volume_into(out, prv, nxt, search_range: int = 4, leaky_slope: float = 0.1):
    """Inference-only ``cost_volume`` into a caller-owned ``out`` of shape (B,H,W,S), S >= (2d+1)^2:
    the cost volume lands in channels [0, (2d+1)^2) of every pixel (concat-buffer epilogue,
    qpwcnet/core/non_layers.py:335-336).  Host tensors take the staged path (S must be dense)."""
    prv, nxt = _prep(prv, "prv"), _prep(nxt, "nxt")
    _same(prv, nxt, "cost_volume_into(prv, nxt)")
    _check_out(out, prv)
    _no_grad_into("cost_volume_into", prv, nxt)
    return _corr_fwd(prv, nxt, int(search_range), float(leaky_slope), out=out, out_stride=out.shape[-1])


def warp_cost_volume_into(out, prv, nxt, flow, mode="tfa", search_range: int = 4,
                          leaky_slope: float = 0.1):
    """Inference-only fused ``warp_cost_volume`` into a caller-owned (possibly wider) ``out``."""
    prv, nxt = _prep(prv, "prv"), _prep(nxt, "nxt")
    flow = _prep(flow, "flow", last=2)
    _same(prv, nxt, "warp_cost_volume_into(prv, nxt)")
    _check_out(out, prv)
    _check_flow("warp_cost_volume_into", prv, flow, _mode(mode))
    _no_grad_into("warp_cost_volume_into", prv, nxt, flow)
    return _warp_corr_fwd(prv, nxt, flow, _mode(mode), int(search_range), float(leaky_slope),
                          out=out, out_stride=out.shape[-1])


def warp_into(out, img, flow, mode="tfa"):
    """Inference-only ``warp`` into a caller-owned ``out`` (same shape as ``img``)."""
    img = _prep(img, "img")
    flow = _prep(flow, "flow", last=2)
    if out.shape != img.shape or out.dtype != torch.float32 or out.device != img.device or not out.is_contiguous():
        raise ValueError("out must be a dense float32 tensor shaped like img on the same device")
    _check_flow("warp_into", img, flow, _mode(mode))
    _no_grad_into("warp_into", img, flow)
    return _warp_fwd(img, flow, _mode(mode), out=out)


def warp_rows_into(out, img, flow, mode, row_offset: int, full_height: int):
    """Inference-only warp of rows [row_offset, row_offset + H) of a `full_height`-row frame (row-sharded
    frames): absolute-row arithmetic, bit-identical to the unsharded warp for rows whose taps lie inside
    the view."""
    img = _prep(img, "img")
    flow = _prep(flow, "flow", last=2)
    if out.shape != img.shape or out.dtype != torch.float32 or out.device != img.device or not out.is_contiguous():
        raise ValueError("out must be a dense float32 tensor shaped like img on the same device")
    if not img.is_cuda:
        raise ValueError("warp_rows_into needs CUDA tensors")
    m = _mode(mode)
    if img.shape[:3] != flow.shape[:3] or img.device != flow.device:
        raise ValueError(f"warp_rows_into: image {tuple(img.shape)} vs flow {tuple(flow.shape)}")
    _no_grad_into("warp_rows_into", img, flow)
    B, H, W, C = img.shape
    vi, vf, vo = _views(img, flow, out)
    with _on_device(img.device):
        check(lib().qpwc_warp_fwd_rows(vi.ptr, vf.ptr, vo.ptr, B, H, W, C, m, int(row_offset), int(full_height),
                                       _stream_ptr(img.device)))
    return out


def _check_out(out, prv):
    if out.dtype != torch.float32 or out.dim() != 4 or out.shape[:3] != prv.shape[:3] \
            or out.device != prv.device or not out.is_contiguous():
        raise ValueError(f"out must be a dense float32 (B,H,W,S) tensor on {prv.device}, "
                         f"got {tuple(out.shape)} {out.dtype} @ {out.device}")


class host_batch:
    """``with ops.host_batch():`` -- host-buffer (CPU tensor) calls inside the block only enqueue
    their H2D / kernels / D2H on the library's streams and overlap with each other; the block exit
    waits for all of them.  Inputs and outputs must stay alive and untouched until then."""

    def __enter__(self):
        check(lib().qpwc_host_set_deferred(1))
        return self

    def __exit__(self, *exc):
        check(lib().qpwc_host_set_deferred(0))
        check(lib().qpwc_host_sync(-1))      # every device that received deferred work
        return False


_ENGINES = {"auto": 0, "ffma": 1, "tc": 2}


def set_corr_engine(name: str) -> None:
    """Cost-volume forward arithmetic: 'auto' (tensor cores, 3xTF32 split, where the shape allows),
    'ffma' (plain fp32 FFMA kernels only) or 'tc' (tensor cores or an error).  Process-wide."""
    check(lib().qpwc_set_option(0, _ENGINES[name]))


def get_corr_engine() -> str:
    v = int(lib().qpwc_get_option(0))
    return {v_: k for k, v_ in _ENGINES.items()}[v]


def library_version() -> int:
    return int(lib().qpwc_version())


def bind_host_thread_near(device=None):
    """Pin the calling thread to the CPUs NVML reports as local to ``device`` (its NUMA node), so that
    pinned host buffers allocated afterwards are first-touched next to the GPU's PCIe root.  Host-buffer
    throughput (``*_host`` entry points) drops 2-3x when the staging pages land on the other socket.
    Best effort: returns the previous affinity set (pass it to ``os.sched_setaffinity(0, ...)`` to undo),
    or None when NVML / the affinity call is unavailable."""
    import os
    try:
        import pynvml
        idx = torch.cuda.current_device() if device is None else torch.device(device).index or 0
        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(idx).uuid)
        try:
            h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        old = os.sched_getaffinity(0)
        cpus &= old
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return old
    except Exception:
        return None
