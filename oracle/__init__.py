"""CPU oracle for the qpwcnet cost-volume / warp hot path  --  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import this package.  The product package ``qpwcnet_b200`` never does, and has
no CPU path: it fails loudly when its CUDA library is missing.

numpy front-end over ``oracle/qpwc_oracle.c`` (plain-C restatement of the reference, fp32 + fp64).
Pinning status, reference citations and the "parity unpinned" note for the tfa-semantics warp are
in the header of ``qpwc_oracle.c``.

All arrays are NHWC (``channels_last``), C-contiguous, float32 or float64 (the dtype of the first
argument selects the arithmetic).  Warp modes: ``"tf"`` = ``Warp``/``tf_warp``
(qpwcnet/core/warp.py:63-153), ``"tfa"`` = ``WarpV2`` (qpwcnet/core/layers.py:171-186).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libqpwc_oracle.so")
_lib = None

MODES = ("tf", "tfa")


def build(force: bool = False) -> str:
    """Compile the C restatement with the recipe in oracle/Makefile; returns the .so path."""
    srcs = [os.path.join(_HERE, f) for f in ("qpwc_oracle.c", "qpwc_oracle_body.inc", "Makefile")]
    stale = (not os.path.exists(_SO)) or any(
        os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-s"] + (["-B"] if force else []), check=True)
    return _SO


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.qo_num_threads.restype = ctypes.c_int
    return _lib


def num_threads() -> int:
    return int(_load().qo_num_threads())


def set_num_threads(n: int) -> None:
    _load().qo_set_num_threads(ctypes.c_int(int(n)))


def _suffix(a: np.ndarray) -> str:
    if a.dtype == np.float32:
        return "f32"
    if a.dtype == np.float64:
        return "f64"
    raise TypeError(f"oracle supports float32/float64, got {a.dtype}")


def _c(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


def _p(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


def _real(a: np.ndarray, v: float):
    return ctypes.c_float(v) if a.dtype == np.float32 else ctypes.c_double(v)


def cost_volume(prv, nxt, search_range: int = 4, slope: float = 0.1, out_stride: int | None = None):
    """CostVolume((prv, nxt)) -- layers.py:72-100.  Returns (B,H,W,(2d+1)^2) (or stride-padded)."""
    prv = _c(prv, prv.dtype)
    nxt = _c(nxt, prv.dtype)
    B, H, W, C = prv.shape
    assert nxt.shape == prv.shape
    d = int(search_range)
    D = (2 * d + 1) ** 2
    ops = D if out_stride is None else int(out_stride)
    assert ops >= D
    out = np.zeros((B, H, W, ops), dtype=prv.dtype)
    fn = getattr(_load(), "qo_cost_volume_" + _suffix(prv))
    fn(_p(prv), _p(nxt), _p(out), B, H, W, C, d, _real(prv, slope), ctypes.c_longlong(ops))
    return out


def cost_volume_bwd(prv, nxt, out, g_out, search_range: int = 4, slope: float = 0.1):
    """Gradients (g_prv, g_nxt) of cost_volume; ``out`` is the forward result (leaky mask)."""
    prv = _c(prv, prv.dtype)
    nxt = _c(nxt, prv.dtype)
    out = _c(out, prv.dtype)
    g_out = _c(g_out, prv.dtype)
    B, H, W, C = prv.shape
    d = int(search_range)
    ops = out.shape[-1]
    assert g_out.shape == out.shape
    g_prv = np.empty_like(prv)
    g_nxt = np.empty_like(prv)
    fn = getattr(_load(), "qo_cost_volume_bwd_" + _suffix(prv))
    fn(_p(prv), _p(nxt), _p(out), _p(g_out), _p(g_prv), _p(g_nxt), B, H, W, C, d,
       _real(prv, slope), ctypes.c_longlong(ops))
    return g_prv, g_nxt


def warp(img, flow, mode: str = "tfa"):
    """Warp((img, flow)) [mode 'tf', warp.py:63-153] / WarpV2((img, flow)) [mode 'tfa']."""
    assert mode in MODES
    img = _c(img, img.dtype)
    flow = _c(flow, img.dtype)
    B, H, W, C = img.shape
    assert flow.shape == (B, H, W, 2)
    if mode == "tfa" and (H < 2 or W < 2):
        raise ValueError("Grid must be at least 2x2 (tfa interpolate_bilinear)")
    out = np.empty_like(img)
    fn = getattr(_load(), f"qo_warp_{mode}_" + _suffix(img))
    fn(_p(img), _p(flow), _p(out), B, H, W, C)
    return out


def warp_bwd(img, flow, g_out, mode: str = "tfa"):
    """Gradients (g_img, g_flow) of warp."""
    assert mode in MODES
    img = _c(img, img.dtype)
    flow = _c(flow, img.dtype)
    g_out = _c(g_out, img.dtype)
    B, H, W, C = img.shape
    g_img = np.empty_like(img)
    g_flow = np.empty_like(flow)
    fn = getattr(_load(), f"qo_warp_{mode}_bwd_" + _suffix(img))
    fn(_p(img), _p(flow), _p(g_out), _p(g_img), _p(g_flow), B, H, W, C)
    return g_img, g_flow


def warp_cost_volume(prv, nxt, flow, mode: str = "tfa", search_range: int = 4, slope: float = 0.1,
                     out_stride: int | None = None):
    """UpFlow's ``CostVolumeV2((prv, WarpV2((nxt, flo))))`` -- non_layers.py:377-380."""
    return cost_volume(prv, warp(nxt, flow, mode), search_range, slope, out_stride)


def warp_cost_volume_bwd(prv, nxt, flow, g_out, mode: str = "tfa", search_range: int = 4,
                         slope: float = 0.1):
    """Gradients (g_prv, g_nxt, g_flow) of warp_cost_volume (chain rule over the two oracles)."""
    nxt_w = warp(nxt, flow, mode)
    out = cost_volume(prv, nxt_w, search_range, slope, g_out.shape[-1])
    g_prv, g_nxt_w = cost_volume_bwd(prv, nxt_w, out, g_out, search_range, slope)
    g_nxt, g_flow = warp_bwd(nxt, flow, g_nxt_w, mode)
    return g_prv, g_nxt, g_flow


# ------------------------------------------------------------------ x2 bilinear upsampling (numpy)
def occlusion_map(flow):
    """estimate_occlusion_map (qpwcnet/core/occlusion.py:27-118) on an NHWC flow (B,H,W,2), fp32.
    Parity unpinned (no TF here): a restatement on top of the pinned `warp(mode='tf')`."""
    flow = _c(flow, np.float32)
    B, H, W, _ = flow.shape
    i, j = np.meshgrid(np.arange(H, dtype=np.float32), np.arange(W, dtype=np.float32), indexing="ij")
    dj, di = flow[..., 0], flow[..., 1]                               # occlusion.py:59
    i2, j2 = i + di, j + dj
    oob = ((i2 < 0) | (i2 >= H) | (j2 < 0) | (j2 >= W)).astype(np.float32)   # occlusion.py:74
    inv = -warp(flow, flow, "tf")                                     # occlusion.py:83
    i3 = np.clip(np.trunc(i + inv[..., 1]), 0, H - 1).astype(np.int64)       # occlusion.py:85-90
    j3 = np.clip(np.trunc(j + inv[..., 0]), 0, W - 1).astype(np.int64)
    map3 = np.ones((B, H, W), np.float32)
    b = np.broadcast_to(np.arange(B)[:, None, None], (B, H, W))
    map3[b, i3, j3] = 0.0                                             # scatter-min of zeros, occlusion.py:92-93
    return np.maximum(oob, map3)                                      # occlusion.py:96


def _up2_coords(n_out, n_in, dtype):
    """tf.image.resize bilinear, half-pixel centres (ResizeBilinear kernel, restated from its
    published algorithm; TF is not installable here => parity for this op is UNPINNED):
    in = (o + 0.5) * (n_in / n_out) - 0.5; lo = max(floor(in), 0); hi = min(ceil(in), n_in - 1);
    lerp = in - floor(in)."""
    o = np.arange(n_out, dtype=dtype)
    src = (o + dtype(0.5)) * dtype(n_in / n_out) - dtype(0.5)
    fl = np.floor(src)
    lo = np.maximum(fl.astype(np.int64), 0)
    hi = np.minimum(np.ceil(src).astype(np.int64), n_in - 1)
    return lo, hi, (src - fl).astype(dtype)


def upsample2x(x, scale: float = 1.0):
    """``Upsample(scale)`` -- qpwcnet/core/non_layers.py:183-193: ``scale *
    UpSampling2D(interpolation='bilinear')(x)``; NHWC (B,H,W,C) -> (B,2H,2W,C), arithmetic in x.dtype
    with every product/sum rounded on its own: top = tl + (tr-tl)*xl; bot = bl + (br-bl)*xl;
    out = top + (bot-top)*yl."""
    x = np.ascontiguousarray(x)
    dt = x.dtype.type
    B, H, W, C = x.shape
    ylo, yhi, yl = _up2_coords(2 * H, H, dt)
    xlo, xhi, xl = _up2_coords(2 * W, W, dt)
    tl, tr = x[:, ylo][:, :, xlo], x[:, ylo][:, :, xhi]
    bl, br = x[:, yhi][:, :, xlo], x[:, yhi][:, :, xhi]
    xl_, yl_ = xl[None, None, :, None], yl[None, :, None, None]
    top = tl + (tr - tl) * xl_
    bot = bl + (br - bl) * xl_
    return (dt(scale) * (top + (bot - top) * yl_)).astype(x.dtype)


def upsample2x_bwd(g_out, scale: float = 1.0):
    """Adjoint of ``upsample2x`` (what TF autodiff / ResizeBilinearGrad computes up to summation
    order): g_in[lo/hi] += weight * g_out."""
    g_out = np.ascontiguousarray(g_out)
    dt = g_out.dtype.type
    B, H2, W2, C = g_out.shape
    H, W = H2 // 2, W2 // 2
    ylo, yhi, yl = _up2_coords(H2, H, dt)
    xlo, xhi, xl = _up2_coords(W2, W, dt)
    tmp = np.zeros((B, H, W2, C), dtype=g_out.dtype)          # rows first
    np.add.at(tmp, (slice(None), ylo), g_out * (dt(1) - yl)[None, :, None, None])
    np.add.at(tmp, (slice(None), yhi), g_out * yl[None, :, None, None])
    g = np.zeros((B, H, W, C), dtype=g_out.dtype)
    np.add.at(g, (slice(None), slice(None), xlo), tmp * (dt(1) - xl)[None, None, :, None])
    np.add.at(g, (slice(None), slice(None), xhi), tmp * xl[None, None, :, None])
    return (dt(scale) * g).astype(g_out.dtype)
