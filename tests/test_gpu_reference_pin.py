"""CUDA library (through the drop-in layer classes, i.e. ctypes -> C ABI) against the fixtures minted
by EXECUTING the reference's own source (oracle/pin_to_reference.py), both data formats.

Tolerances are BASELINE.json's: cost volume <= 1e-5 relative, graded against the reference's `exact`
run both as max-norm and per element condition-aware (|delta| <= 1e-5 * (1/C) sum_c |prv*nxt|, SURVEY
8c); warp forward bit-identical to the reference's fp32 output (strong form of <= 1e-6 absolute);
warp gradients <= 1e-6 absolute against exact arithmetic, or as close as the reference's own fp32
run, and the achieved max-abs errors are printed (`-s`) and collected in ACHIEVED."""
import os

import numpy as np
import pytest
import torch

from qpwcnet.core import layers, non_layers
from qpwcnet.core.warp import tf_warp
from qpwcnet_b200 import ops

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PIN = np.load(os.path.join(GOLD, "ref_pin.npz"))
CFG1 = np.load(os.path.join(GOLD, "ref_cfg1.npz"))
DEV = "cuda:0"
FORMATS = ("channels_last", "channels_first")
ACHIEVED = {}


def names(prefix):
    return sorted({k.split("/")[0] for k in PIN.files if k.startswith(prefix)})


def dev(a, fmt="channels_last"):
    t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(DEV)
    return t.permute(0, 3, 1, 2).contiguous() if fmt == "channels_first" else t


def host(t, fmt="channels_last"):
    t = t.detach()
    if fmt == "channels_first" and t.dim() == 4:
        t = t.permute(0, 2, 3, 1)
    return t.cpu().numpy()


def note(key, got, exact):
    e = float(np.abs(got.astype(np.float64) - exact).max()) if got.size else 0.0
    ACHIEVED[key] = max(ACHIEVED.get(key, 0.0), e)
    print(f"[achieved] {key}: max|gpu - exact| = {e:.3e}")
    return e


def cv_condition(prv, nxt, d):
    """(1/C) sum_c |prv[p,c] * nxt[p+disp,c]| per output element (fp64)."""
    B, H, W, C = prv.shape
    pad = np.zeros((B, H + 2 * d, W + 2 * d, C))
    pad[:, d:d + H, d:d + W] = np.abs(nxt)
    out = np.empty((B, H, W, (2 * d + 1) ** 2))
    ap = np.abs(prv).astype(np.float64)
    for i0 in range(2 * d + 1):
        for j0 in range(2 * d + 1):
            out[..., i0 * (2 * d + 1) + j0] = (ap * pad[:, i0:i0 + H, j0:j0 + W]).mean(-1)
    return out


def assert_cv(got, exact, prv, nxt, d, key):
    err = np.abs(got.astype(np.float64) - exact)
    note(key, got, exact)
    assert err.max() <= 1e-5 * np.abs(exact).max()
    cond = cv_condition(prv, nxt, d)
    bad = err > 1e-5 * cond + 1e-30
    assert not bad.any(), f"{bad.sum()} elements exceed 1e-5 * mean|prv*nxt| (worst ratio {np.max(err / (cond + 1e-30)):.2e})"


def assert_as_accurate(got, ref32, exact, key, tol=1e-6, factor=4.0):
    e_got = note(key, got, exact)
    e_ref = float(np.abs(ref32.astype(np.float64) - exact).max())
    assert e_got <= max(tol, factor * e_ref), f"gpu off exact by {e_got:.3e}; reference's own fp32 run by {e_ref:.3e}"


@pytest.mark.parametrize("fmt", FORMATS)
@pytest.mark.parametrize("name", names("cv_"))
def test_cost_volume_vs_reference_run(name, fmt):
    prv, nxt, g, d = (PIN[f"{name}/{k}"] for k in ("prv", "nxt", "g_out", "d"))
    d = int(d)
    tp, tn = dev(prv, fmt).requires_grad_(), dev(nxt, fmt).requires_grad_()
    for cls in (layers.CostVolume, layers.CostVolumeV2, non_layers.CostVolume):
        out = cls(search_range=d, data_format=fmt)((tp, tn))
        assert_cv(host(out, fmt), PIN[f"{name}/out_exact"], prv, nxt, d, f"cv_fwd/{fmt}")
    gp, gn = torch.autograd.grad(out, (tp, tn), dev(g, fmt))
    for got, k in ((gp, "g_prv"), (gn, "g_nxt")):
        e = note(f"cv_bwd/{fmt}", host(got, fmt), PIN[f"{name}/{k}_exact"])
        assert e <= 1e-5 * np.abs(PIN[f"{name}/{k}_exact"]).max()


@pytest.mark.parametrize("fmt", FORMATS)
@pytest.mark.parametrize("name", names("warp_") + names("half_"))
def test_warp_vs_reference_run(name, fmt):
    img, flow, g = (PIN[f"{name}/{k}"] for k in ("img", "flow", "g_out"))
    ti, tf_ = dev(img, fmt).requires_grad_(), dev(flow, fmt).requires_grad_()
    if name.startswith("half_"):        # FrameInterpolate: warp((img, 0.5 * flo)), non_layers.py:303-304
        out = non_layers.Warp(data_format=fmt)((ti, 0.5 * tf_))
    else:
        out = layers.Warp(data_format=fmt)((ti, tf_))
        np.testing.assert_array_equal(host(tf_warp(ti, tf_, fmt), fmt), PIN[f"{name}/out"])
    np.testing.assert_array_equal(host(out, fmt), PIN[f"{name}/out"])           # bit-identical
    note(f"warp_fwd/{fmt}", host(out, fmt), PIN[f"{name}/out_exact"])
    gi, gf = torch.autograd.grad(out, (ti, tf_), dev(g, fmt))
    assert_as_accurate(host(gi, fmt), PIN[f"{name}/g_img"], PIN[f"{name}/g_img_exact"], f"warp_bwd_img/{fmt}")
    assert_as_accurate(host(gf, fmt), PIN[f"{name}/g_flow"], PIN[f"{name}/g_flow_exact"], f"warp_bwd_flow/{fmt}")
    if f"{name}/tfa/out" in PIN.files:
        out2 = layers.WarpV2(data_format=fmt)((ti, tf_))
        np.testing.assert_array_equal(host(out2, fmt), PIN[f"{name}/tfa/out"])


@pytest.mark.parametrize("fmt", FORMATS)
@pytest.mark.parametrize("name", names("fused_"))
def test_upflow_pair_vs_reference_run(name, fmt):
    prv, nxt, flow, g, d = (PIN[f"{name}/{k}"] for k in ("prv", "nxt", "flow", "g_out", "d"))
    d = int(d)
    tp, tn, tf_ = (dev(a, fmt).requires_grad_() for a in (prv, nxt, flow))
    out = layers.WarpCostVolume(search_range=d, warp_mode="tf", data_format=fmt)((tp, tn, tf_))
    e = note(f"fused_fwd/{fmt}", host(out, fmt), PIN[f"{name}/out_exact"])
    assert e <= 1e-5 * np.abs(PIN[f"{name}/out_exact"]).max()
    gp, gn, gf = torch.autograd.grad(out, (tp, tn, tf_), dev(g, fmt))
    for got, k in ((gp, "g_prv"), (gn, "g_nxt")):
        e = note(f"fused_bwd/{fmt}", host(got, fmt), PIN[f"{name}/{k}_exact"])
        assert e <= 1e-5 * np.abs(PIN[f"{name}/{k}_exact"]).max()
    assert_as_accurate(host(gf, fmt), PIN[f"{name}/g_flow"], PIN[f"{name}/g_flow_exact"], f"fused_bwd_flow/{fmt}")


@pytest.mark.parametrize("fmt", FORMATS)
@pytest.mark.parametrize("name", names("occ_"))
def test_occlusion_map_vs_reference_run(name, fmt):
    got = ops.occlusion_map(dev(PIN[f"{name}/flow"], fmt), data_format=fmt)
    np.testing.assert_array_equal(host(got), PIN[f"{name}/map"])


@pytest.mark.parametrize("name", names("kat_"))
def test_known_answer_cases(name):
    img, flow = dev(PIN[f"{name}/img"]), dev(PIN[f"{name}/flow"])
    np.testing.assert_array_equal(host(layers.Warp()((img, flow))), PIN[f"{name}/tf"])
    np.testing.assert_array_equal(host(layers.WarpV2()((img, flow))), PIN[f"{name}/tfa"])


@pytest.mark.parametrize("fmt", FORMATS)
def test_config1_reference_shapes(fmt):
    """test/test_cost_volume.py:20-21, test/test_warp.py:24-25: (4,32,64,3), d = 4."""
    r1 = np.random.default_rng(int(CFG1["seed"]))
    f32 = lambda a: np.asarray(a, dtype=np.float32)  # noqa: E731
    prv, nxt = f32(r1.standard_normal((4, 32, 64, 3))), f32(r1.standard_normal((4, 32, 64, 3)))
    img, flo = f32(r1.random((4, 32, 64, 3))), f32(r1.standard_normal((4, 32, 64, 2)))
    g_cv, g_w = f32(r1.standard_normal((4, 32, 64, 81))), f32(r1.standard_normal((4, 32, 64, 3)))
    samp = (slice(None), slice(None, None, 5), slice(None, None, 7))
    tp, tn = dev(prv, fmt).requires_grad_(), dev(nxt, fmt).requires_grad_()
    c1 = layers.CostVolume(search_range=4, data_format=fmt)((tp, tn))
    c2 = layers.CostVolumeV2(search_range=4, data_format=fmt)((tp, tn))
    assert float((c1 - c2).sum()) == 0.0                       # what test_cost_volume.py prints
    cv = host(c1, fmt)
    np.testing.assert_allclose(cv[samp], CFG1["cv/sample"], rtol=0, atol=1e-5 * np.abs(CFG1["cv/sample"]).max())
    assert abs(cv.sum(dtype=np.float64) - float(CFG1["cv/sum"])) <= 1e-5 * np.abs(cv).sum(dtype=np.float64)
    gp, gn = torch.autograd.grad(c1, (tp, tn), dev(g_cv, fmt))
    np.testing.assert_allclose(host(gp, fmt)[samp], CFG1["cv/g_prv"], rtol=0, atol=1e-5 * np.abs(CFG1["cv/g_prv"]).max())
    np.testing.assert_allclose(host(gn, fmt)[samp], CFG1["cv/g_nxt"], rtol=0, atol=1e-5 * np.abs(CFG1["cv/g_nxt"]).max())
    ti, tf_ = dev(img, fmt).requires_grad_(), dev(flo, fmt).requires_grad_()
    for mode, cls in (("tf", layers.Warp), ("tfa", layers.WarpV2)):
        w = cls(data_format=fmt)((ti, tf_))
        np.testing.assert_array_equal(host(w, fmt)[samp], CFG1[f"warp/{mode}/sample"])
        gi, gf = torch.autograd.grad(w, (ti, tf_), dev(g_w, fmt))
        np.testing.assert_allclose(host(gi, fmt)[samp], CFG1[f"warp/{mode}/g_img"], rtol=0, atol=2e-6)
        np.testing.assert_allclose(host(gf, fmt)[samp], CFG1[f"warp/{mode}/g_flow"], rtol=0, atol=4e-6)
