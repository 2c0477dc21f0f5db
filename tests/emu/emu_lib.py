"""numpy front-end over tests/emu/_build/libqpwc_emu.so (the kernels compiled for the CPU emulation
harness).  Test infrastructure only -- see tests/emu/README.md."""
import ctypes
import importlib.util
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("build_emu", os.path.join(_HERE, "build_emu.py"))
_build = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_build)

_lib = None
MODES = {"tf": 0, "tfa": 1}


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(_build.build())
        _lib.qpwc_last_error.restype = ctypes.c_char_p
        _lib.qpwc_warp_corr_bwd_workspace.restype = ctypes.c_size_t
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _ck(rc):
    if rc != 0:
        raise RuntimeError(f"emu libqpwc error {rc}: {lib().qpwc_last_error().decode()}")


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def corr_fwd(prv, nxt, d=4, slope=0.1, ops=None):
    prv, nxt = _f(prv), _f(nxt)
    B, H, W, C = prv.shape
    D = (2 * d + 1) ** 2
    ops = D if ops is None else ops
    out = np.full((B, H, W, ops), np.nan, np.float32)
    _ck(lib().qpwc_corr_fwd(_p(prv), _p(nxt), _p(out), B, H, W, C, d, ctypes.c_float(slope),
                            ctypes.c_longlong(ops), None))
    return out


def corr_fwd_nchw(prv, nxt, d=4, slope=0.1):
    """prv, nxt: (B,C,H,W) -> (B,(2d+1)^2,H,W) through the native channels_first kernel."""
    prv, nxt = _f(prv), _f(nxt)
    B, C, H, W = prv.shape
    out = np.full((B, (2 * d + 1) ** 2, H, W), np.nan, np.float32)
    _ck(lib().qpwc_corr_fwd_nchw(_p(prv), _p(nxt), _p(out), B, C, H, W, d, ctypes.c_float(slope), None))
    return out


def occlusion_map(flow, channels_first=False):
    flow = _f(flow)
    B = flow.shape[0]
    H, W = (flow.shape[2], flow.shape[3]) if channels_first else (flow.shape[1], flow.shape[2])
    out = np.full((B, H, W), np.nan, np.float32)
    _ck(lib().qpwc_occlusion_map(_p(flow), _p(out), B, H, W, int(channels_first), None))
    return out


def corr_bwd_nchw(prv, nxt, out, g_out, d=4, slope=0.1):
    prv, nxt, out, g_out = _f(prv), _f(nxt), _f(out), _f(g_out)
    B, C, H, W = prv.shape
    gp, gn = np.full_like(prv, np.nan), np.full_like(prv, np.nan)
    _ck(lib().qpwc_corr_bwd_nchw(_p(prv), _p(nxt), _p(out), _p(g_out), _p(gp), _p(gn), B, C, H, W, d,
                                 ctypes.c_float(slope), None))
    return gp, gn


def corr_bwd(prv, nxt, out, g_out, d=4, slope=0.1):
    prv, nxt, out, g_out = _f(prv), _f(nxt), _f(out), _f(g_out)
    B, H, W, C = prv.shape
    gp, gn = np.full_like(prv, np.nan), np.full_like(prv, np.nan)
    _ck(lib().qpwc_corr_bwd(_p(prv), _p(nxt), _p(out), _p(g_out), _p(gp), _p(gn), B, H, W, C, d,
                            ctypes.c_float(slope), ctypes.c_longlong(out.shape[-1]), None))
    return gp, gn


def warp_fwd(img, flow, mode):
    img, flow = _f(img), _f(flow)
    B, H, W, C = img.shape
    out = np.full_like(img, np.nan)
    _ck(lib().qpwc_warp_fwd(_p(img), _p(flow), _p(out), B, H, W, C, MODES[mode], None))
    return out


def warp_bwd(img, flow, g_out, mode):
    img, flow, g_out = _f(img), _f(flow), _f(g_out)
    B, H, W, C = img.shape
    gi, gf = np.full_like(img, np.nan), np.full_like(flow, np.nan)
    _ck(lib().qpwc_warp_bwd(_p(img), _p(flow), _p(g_out), _p(gi), _p(gf), B, H, W, C, MODES[mode], None))
    return gi, gf


def warp_fwd_nchw(img, flow, mode):
    img, flow = _f(img), _f(flow)
    B, C, H, W = img.shape
    out = np.full_like(img, np.nan)
    _ck(lib().qpwc_warp_fwd_nchw(_p(img), _p(flow), _p(out), B, C, H, W, MODES[mode], ctypes.c_float(1.0), None))
    return out


def warp_bwd_nchw(img, flow, g_out, mode):
    img, flow, g_out = _f(img), _f(flow), _f(g_out)
    B, C, H, W = img.shape
    gi, gf = np.full_like(img, np.nan), np.full_like(flow, np.nan)
    _ck(lib().qpwc_warp_bwd_nchw(_p(img), _p(flow), _p(g_out), _p(gi), _p(gf), B, C, H, W, MODES[mode],
                                 ctypes.c_float(1.0), None))
    return gi, gf


def warp_corr_fwd(prv, nxt, flow, mode, d=4, slope=0.1, ops=None):
    prv, nxt, flow = _f(prv), _f(nxt), _f(flow)
    B, H, W, C = prv.shape
    D = (2 * d + 1) ** 2
    ops = D if ops is None else ops
    out = np.full((B, H, W, ops), np.nan, np.float32)
    _ck(lib().qpwc_warp_corr_fwd(_p(prv), _p(nxt), _p(flow), _p(out), B, H, W, C, d,
                                 ctypes.c_float(slope), MODES[mode], ctypes.c_longlong(ops), None))
    return out


def warp_corr_bwd(prv, nxt, flow, out, g_out, mode, d=4, slope=0.1):
    prv, nxt, flow, out, g_out = _f(prv), _f(nxt), _f(flow), _f(out), _f(g_out)
    B, H, W, C = prv.shape
    gp, gn, gf = np.full_like(prv, np.nan), np.full_like(prv, np.nan), np.full_like(flow, np.nan)
    n = lib().qpwc_warp_corr_bwd_workspace(B, H, W, C)
    ws = np.zeros(n // 4 + 4, np.float32)
    _ck(lib().qpwc_warp_corr_bwd(_p(prv), _p(nxt), _p(flow), _p(out), _p(g_out), _p(gp), _p(gn),
                                 _p(gf), _p(ws), ctypes.c_size_t(ws.nbytes), B, H, W, C, d,
                                 ctypes.c_float(slope), MODES[mode],
                                 ctypes.c_longlong(out.shape[-1]), None))
    return gp, gn, gf


def corr_fwd_host(prv, nxt, d=4, slope=0.1):
    prv, nxt = _f(prv), _f(nxt)
    B, H, W, C = prv.shape
    out = np.full((B, H, W, (2 * d + 1) ** 2), np.nan, np.float32)
    _ck(lib().qpwc_corr_fwd_host(_p(prv), _p(nxt), _p(out), B, H, W, C, d, ctypes.c_float(slope), 0))
    return out
