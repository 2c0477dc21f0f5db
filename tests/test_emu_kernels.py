"""Runs the CUDA kernel bodies on the CPU emulation harness (tests/emu) and checks them against the
oracle.  This validates index arithmetic / tiling / epilogues in the GPU-less container; the real
parity tests (tests/test_gpu_parity.py, -m gpu) run the sm_100a build through the C ABI."""
import os
import sys

import numpy as np
import pytest

import oracle

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
import emu_lib  # noqa: E402


def rng(s):
    return np.random.default_rng(s)


def cv_tol(ref64, prv, nxt):
    return 1e-5 * np.abs(ref64).max()


@pytest.mark.parametrize("B,H,W,C,d", [(2, 6, 7, 3, 4), (1, 5, 9, 8, 2), (1, 4, 5, 5, 1)])
def test_emu_corr_fwd_direct(B, H, W, C, d):
    r = rng(0)
    prv = r.standard_normal((B, H, W, C)).astype(np.float32)
    nxt = r.standard_normal((B, H, W, C)).astype(np.float32)
    ref = oracle.cost_volume(prv.astype(np.float64), nxt.astype(np.float64), d)
    got = emu_lib.corr_fwd(prv, nxt, d)
    assert np.abs(got - ref).max() <= 1e-5 * np.abs(ref).max()
    got_s = emu_lib.corr_fwd(prv, nxt, d, ops=(2 * d + 1) ** 2 + 5)
    np.testing.assert_array_equal(got_s[..., :(2 * d + 1) ** 2], got)
    assert np.isnan(got_s[..., (2 * d + 1) ** 2:]).all()      # padding lanes untouched


@pytest.mark.parametrize("B,H,W,C,d", [(2, 6, 7, 3, 4), (1, 5, 9, 8, 2)])
def test_emu_corr_bwd_direct(B, H, W, C, d):
    r = rng(1)
    prv = r.standard_normal((B, H, W, C)).astype(np.float32)
    nxt = r.standard_normal((B, H, W, C)).astype(np.float32)
    out = oracle.cost_volume(prv, nxt, d)
    g = r.standard_normal(out.shape).astype(np.float32)
    gp, gn = oracle.cost_volume_bwd(*(a.astype(np.float64) for a in (prv, nxt, out, g)), d)
    gp2, gn2 = emu_lib.corr_bwd(prv, nxt, out, g, d)
    assert np.abs(gp2 - gp).max() <= 1e-5 * np.abs(gp).max()
    assert np.abs(gn2 - gn).max() <= 1e-5 * np.abs(gn).max()


@pytest.mark.parametrize("mode", ["tf", "tfa"])
@pytest.mark.parametrize("B,H,W,C", [(2, 6, 7, 3), (1, 5, 9, 8), (1, 4, 6, 2), (1, 3, 5, 12)])
def test_emu_warp_fwd_bwd(mode, B, H, W, C):
    r = rng(2)
    img = r.random((B, H, W, C)).astype(np.float32)
    flow = (r.standard_normal((B, H, W, 2)) * 2.5).astype(np.float32)
    np.testing.assert_array_equal(emu_lib.warp_fwd(img, flow, mode), oracle.warp(img, flow, mode))
    g = r.standard_normal(img.shape).astype(np.float32)
    gi, gf = oracle.warp_bwd(*(a.astype(np.float64) for a in (img, flow, g)), mode)
    gi2, gf2 = emu_lib.warp_bwd(img, flow, g, mode)
    np.testing.assert_allclose(gi2, gi, rtol=0, atol=2e-6)
    np.testing.assert_allclose(gf2, gf, rtol=0, atol=1e-5)


@pytest.mark.parametrize("mode", ["tf", "tfa"])
def test_emu_fused_fwd_bwd_direct(mode):
    r = rng(3)
    B, H, W, C, d = 1, 6, 7, 3, 4
    prv = r.standard_normal((B, H, W, C)).astype(np.float32)
    nxt = r.standard_normal((B, H, W, C)).astype(np.float32)
    flo = (r.standard_normal((B, H, W, 2)) * 2).astype(np.float32)
    a64 = [a.astype(np.float64) for a in (prv, nxt, flo)]
    ref = oracle.warp_cost_volume(*a64, mode, d)
    got = emu_lib.warp_corr_fwd(prv, nxt, flo, mode, d)
    assert np.abs(got - ref).max() <= 1e-5 * np.abs(ref).max()
    g = r.standard_normal(ref.shape).astype(np.float32)
    gp, gn, gf = oracle.warp_cost_volume_bwd(*a64, g.astype(np.float64), mode, d)
    gp2, gn2, gf2 = emu_lib.warp_corr_bwd(prv, nxt, flo, got, g, mode, d)
    assert np.abs(gp2 - gp).max() <= 1e-5 * np.abs(gp).max()
    assert np.abs(gn2 - gn).max() <= 1e-5 * max(np.abs(gn).max(), 1.0)
    assert np.abs(gf2 - gf).max() <= 1e-5 * max(np.abs(gf).max(), 1.0)


def test_emu_host_entry_and_errors():
    r = rng(4)
    prv = r.standard_normal((5, 4, 5, 3)).astype(np.float32)
    nxt = r.standard_normal((5, 4, 5, 3)).astype(np.float32)
    np.testing.assert_array_equal(emu_lib.corr_fwd_host(prv, nxt, 2), emu_lib.corr_fwd(prv, nxt, 2))
    with pytest.raises(RuntimeError, match="search_range"):
        emu_lib.corr_fwd(prv, nxt, 0)
    with pytest.raises(RuntimeError, match="out_pixel_stride"):
        emu_lib.corr_fwd(prv, nxt, 4, ops=80)
    with pytest.raises(RuntimeError, match="2x2"):
        emu_lib.warp_fwd(np.zeros((1, 1, 4, 2), np.float32), np.zeros((1, 1, 4, 2), np.float32), "tfa")


# ------------------------------------------------------------------ register-tiled kernels (d=4, C%4==0)
TILED = [(1, 7, 60, 8), (2, 13, 70, 12), (1, 9, 57, 20), (1, 16, 16, 4),
         (1, 4, 16, 8), (1, 3, 40, 4)]      # last two: single 4-row tile => the 2-row-tile variant


@pytest.mark.parametrize("B,H,W,C", TILED)
def test_emu_corr_fwd_tiled(B, H, W, C):
    r = rng(10)
    prv = r.standard_normal((B, H, W, C)).astype(np.float32)
    nxt = r.standard_normal((B, H, W, C)).astype(np.float32)
    ref = oracle.cost_volume(prv.astype(np.float64), nxt.astype(np.float64), 4)
    got = emu_lib.corr_fwd(prv, nxt, 4)
    assert np.abs(got - ref).max() <= 1e-5 * np.abs(ref).max()
    got_s = emu_lib.corr_fwd(prv, nxt, 4, ops=81 + 3)
    np.testing.assert_array_equal(got_s[..., :81], got)
    assert np.isnan(got_s[..., 81:]).all()


@pytest.mark.parametrize("B,H,W,C", [(1, 7, 60, 8), (2, 13, 70, 12), (1, 9, 57, 20), (1, 16, 16, 4), (1, 3, 40, 4)])
def test_emu_corr_fwd_ragged(B, H, W, C):
    """Tiled FFMA kernel on a second set of shapes: odd heights, ragged widths, channel tails, strided
    output."""
    r = rng(12)
    prv = r.standard_normal((B, H, W, C)).astype(np.float32)
    nxt = r.standard_normal((B, H, W, C)).astype(np.float32)
    ref = oracle.cost_volume(prv.astype(np.float64), nxt.astype(np.float64), 4)
    got = emu_lib.corr_fwd(prv, nxt, 4)
    assert not np.isnan(got).any()
    assert np.abs(got - ref).max() <= 1e-5 * np.abs(ref).max()
    got_s = emu_lib.corr_fwd(prv, nxt, 4, ops=81 + 3)
    np.testing.assert_array_equal(got_s[..., :81], got)
    assert np.isnan(got_s[..., 81:]).all()


@pytest.mark.parametrize("B,C,H,W", [(1, 8, 7, 124), (2, 12, 9, 60), (1, 3, 5, 252), (1, 20, 16, 16), (1, 8, 3, 16)])   # last: 2-row tiles
def test_emu_corr_fwd_nchw(B, C, H, W):
    """Native channels_first kernel (qpwc_corr_nchw.cu): planar TMA tiles, pixel-pair FFMA2, direct
    NCHW stores; ragged tiles (W not a multiple of 120), odd heights, channel tails."""
    r = rng(14)
    prv = r.standard_normal((B, C, H, W)).astype(np.float32)
    nxt = r.standard_normal((B, C, H, W)).astype(np.float32)
    ref = oracle.cost_volume(prv.transpose(0, 2, 3, 1).astype(np.float64), nxt.transpose(0, 2, 3, 1).astype(np.float64), 4)
    got = emu_lib.corr_fwd_nchw(prv, nxt, 4)
    assert not np.isnan(got).any()
    assert np.abs(got.transpose(0, 2, 3, 1) - ref).max() <= 1e-5 * np.abs(ref).max()


def test_emu_corr_fwd_search_range_8_small():
    r = rng(13)
    prv = r.standard_normal((1, 10, 60, 8)).astype(np.float32)
    nxt = r.standard_normal((1, 10, 60, 8)).astype(np.float32)
    ref = oracle.cost_volume(prv.astype(np.float64), nxt.astype(np.float64), 8)
    got = emu_lib.corr_fwd(prv, nxt, 8)
    assert not np.isnan(got).any()
    assert np.abs(got - ref).max() <= 1e-5 * np.abs(ref).max()


@pytest.mark.parametrize("mode", ["tf", "tfa"])
@pytest.mark.parametrize("B,H,W,C", TILED[:3])
def test_emu_fused_fwd_tiled(mode, B, H, W, C):
    r = rng(11)
    prv = r.standard_normal((B, H, W, C)).astype(np.float32)
    nxt = r.standard_normal((B, H, W, C)).astype(np.float32)
    flo = (r.standard_normal((B, H, W, 2)) * 3).astype(np.float32)
    ref = oracle.warp_cost_volume(*(a.astype(np.float64) for a in (prv, nxt, flo)), mode, 4)
    got = emu_lib.warp_corr_fwd(prv, nxt, flo, mode, 4)
    assert np.abs(got - ref).max() <= 1e-5 * np.abs(ref).max()
    # the fused kernel warps with exactly the stand-alone warp kernel's arithmetic
    comp = emu_lib.corr_fwd(prv, emu_lib.warp_fwd(nxt, flo, mode), 4)
    np.testing.assert_allclose(got, comp, rtol=0, atol=1e-6 * np.abs(ref).max())


@pytest.mark.parametrize("B,H,W,C", [(1, 10, 60, 8), (2, 7, 21, 4)])
def test_emu_tiled_search_range_8(B, H, W, C):
    """d = 8 runs the tiled kernel as four 9x9 windows of the 17x17 range."""
    r = rng(31)
    prv = r.standard_normal((B, H, W, C)).astype(np.float32)
    nxt = r.standard_normal((B, H, W, C)).astype(np.float32)
    flo = (r.standard_normal((B, H, W, 2)) * 3).astype(np.float32)
    a64 = [a.astype(np.float64) for a in (prv, nxt, flo)]
    ref = oracle.cost_volume(a64[0], a64[1], 8)
    got = emu_lib.corr_fwd(prv, nxt, 8)
    assert got.shape[-1] == 289 and not np.isnan(got).any()
    assert np.abs(got - ref).max() <= 1e-5 * np.abs(ref).max()
    got_s = emu_lib.corr_fwd(prv, nxt, 8, ops=289 + 7)
    np.testing.assert_array_equal(got_s[..., :289], got)
    assert np.isnan(got_s[..., 289:]).all()
    for mode in ("tf", "tfa"):
        reff = oracle.warp_cost_volume(*a64, mode, 8)
        gotf = emu_lib.warp_corr_fwd(prv, nxt, flo, mode, 8)
        assert np.abs(gotf - reff).max() <= 1e-5 * np.abs(reff).max()


@pytest.mark.parametrize("C", [16, 32])
def test_emu_fused_rolling_rows_long_segments(C, monkeypatch):
    """Fused variant with all channels resident (C <= 32): vertically consecutive tiles keep their
    shared halo rows in shared memory; run in a subprocess with a forced segment length of 8 tiles so
    that the ring rotation cycles through all its phases."""
    import subprocess
    code = r'''
import sys, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests/emu")
import oracle, emu_lib
r = np.random.default_rng(41)
B, H, W, C = 1, 30, 70, int(sys.argv[1])
prv = r.standard_normal((B, H, W, C)).astype(np.float32)
nxt = r.standard_normal((B, H, W, C)).astype(np.float32)
flo = (r.standard_normal((B, H, W, 2)) * 3).astype(np.float32)
for mode in ("tf", "tfa"):
    ref = oracle.warp_cost_volume(*(a.astype(np.float64) for a in (prv, nxt, flo)), mode, 4)
    got = emu_lib.warp_corr_fwd(prv, nxt, flo, mode, 4)
    assert np.abs(got - ref).max() <= 1e-5 * np.abs(ref).max(), mode
print("ok")
'''
    env = dict(os.environ, QPWC_SEG="8")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, "-c", code, str(C)], cwd=root, env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and "ok" in res.stdout, res.stderr[-2000:]


@pytest.mark.parametrize("B,H,W,sigma", [(2, 9, 13, 2.0), (1, 16, 24, 6.0), (1, 1, 5, 1.0), (1, 7, 1, 1.0)])
def test_emu_occlusion_map(B, H, W, sigma):
    """estimate_occlusion_map (occlusion.py:27-118): same map as the oracle, both layouts."""
    flow = (rng(70 + H).standard_normal((B, H, W, 2)) * sigma).astype(np.float32)
    ref = oracle.occlusion_map(flow)
    np.testing.assert_array_equal(emu_lib.occlusion_map(flow), ref)
    np.testing.assert_array_equal(emu_lib.occlusion_map(np.ascontiguousarray(flow.transpose(0, 3, 1, 2)), True), ref)


@pytest.mark.parametrize("mode", ["tf", "tfa"])
@pytest.mark.parametrize("B,C,H,W", [(2, 3, 6, 7), (1, 8, 5, 9), (1, 5, 2, 130)])
def test_emu_warp_nchw_fwd_bwd(mode, B, C, H, W):
    """channels_first warp (warp.py:36-40, layers.py:179-183): native kernels, forward + gradients."""
    r = rng(80 + C)
    img = r.random((B, C, H, W)).astype(np.float32)
    flow = (r.standard_normal((B, 2, H, W)) * 2.5).astype(np.float32)
    nhwc = lambda a: np.ascontiguousarray(a.transpose(0, 2, 3, 1))
    np.testing.assert_array_equal(nhwc(emu_lib.warp_fwd_nchw(img, flow, mode)), oracle.warp(nhwc(img), nhwc(flow), mode))
    g = r.standard_normal(img.shape).astype(np.float32)
    gi, gf = oracle.warp_bwd(*(nhwc(a).astype(np.float64) for a in (img, flow, g)), mode)
    gi2, gf2 = emu_lib.warp_bwd_nchw(img, flow, g, mode)
    # far out-of-image samples extrapolate with large weights (mode tf): fp32 itself is then the
    # limit, so the bound is the fp32 oracle's own distance from the fp64 one
    gi32, gf32 = oracle.warp_bwd(*(nhwc(a) for a in (img, flow, g)), mode)
    assert np.abs(nhwc(gi2) - gi).max() <= max(2e-6, 2 * np.abs(gi32 - gi).max())
    assert np.abs(nhwc(gf2) - gf).max() <= max(1e-5, 2 * np.abs(gf32 - gf).max())


@pytest.mark.parametrize("B,C,H,W", [(1, 5, 6, 8), (2, 9, 5, 132), (1, 17, 9, 260), (1, 3, 1, 4)])
def test_emu_corr_bwd_nchw(B, C, H, W):
    """Gradients of the channels_first cost volume: native kernel (coefficients in registers, planar
    TMA tiles), multi-tile in x and y, ragged channel chunks, against the fp64 oracle."""
    r = rng(90 + C)
    prv = r.standard_normal((B, C, H, W)).astype(np.float32)
    nxt = r.standard_normal((B, C, H, W)).astype(np.float32)
    nhwc = lambda a: np.ascontiguousarray(a.transpose(0, 2, 3, 1))
    out = oracle.cost_volume(nhwc(prv), nhwc(nxt), 4)
    g = r.standard_normal(out.shape).astype(np.float32)
    gp, gn = oracle.cost_volume_bwd(*(a.astype(np.float64) for a in (nhwc(prv), nhwc(nxt), out, g)), 4)
    nchw = lambda a: np.ascontiguousarray(a.transpose(0, 3, 1, 2))
    gp2, gn2 = emu_lib.corr_bwd_nchw(prv, nxt, nchw(out), nchw(g))
    assert not np.isnan(gp2).any() and not np.isnan(gn2).any()      # every element written
    assert np.abs(nhwc(gp2) - gp).max() <= 1e-5 * np.abs(gp).max()
    assert np.abs(nhwc(gn2) - gn).max() <= 1e-5 * np.abs(gn).max()


@pytest.mark.parametrize("B,C,H,W,d", [(2, 7, 8, 14, 4), (1, 5, 6, 7, 2), (1, 3, 9, 10, 8), (2, 4, 5, 9, 1)])
def test_emu_corr_nchw_generic_shapes(B, C, H, W, d):
    """Shape-generic channels_first kernels (W % 4 != 0, search ranges other than 4): forward and both
    gradients against the fp64 oracle -- no shape is left to a transposing route."""
    r = rng(95 + C + W)
    prv = r.standard_normal((B, C, H, W)).astype(np.float32)
    nxt = r.standard_normal((B, C, H, W)).astype(np.float32)
    nhwc = lambda a: np.ascontiguousarray(a.transpose(0, 2, 3, 1))
    nchw = lambda a: np.ascontiguousarray(a.transpose(0, 3, 1, 2))
    ref = oracle.cost_volume(nhwc(prv).astype(np.float64), nhwc(nxt).astype(np.float64), d)
    got = emu_lib.corr_fwd_nchw(prv, nxt, d)
    assert not np.isnan(got).any()
    assert np.abs(nhwc(got) - ref).max() <= 1e-5 * np.abs(ref).max()
    g = r.standard_normal(ref.shape).astype(np.float32)
    gp, gn = oracle.cost_volume_bwd(nhwc(prv).astype(np.float64), nhwc(nxt).astype(np.float64), nhwc(got).astype(np.float64),
                                    g.astype(np.float64), d)
    gp2, gn2 = emu_lib.corr_bwd_nchw(prv, nxt, got, nchw(g), d)
    assert not np.isnan(gp2).any() and not np.isnan(gn2).any()
    assert np.abs(nhwc(gp2) - gp).max() <= 1e-5 * np.abs(gp).max()
    assert np.abs(nhwc(gn2) - gn).max() <= 1e-5 * np.abs(gn).max()
