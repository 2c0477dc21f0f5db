"""Pins the C oracle (oracle/qpwc_oracle.c): against the independent op-by-op torch transcription
of the same reference lines (oracle/ref_torch.py, gradients by autograd), against hand-derived
known-answer cases, and against the committed golden fixtures (tests/golden/)."""
import numpy as np
import pytest
import torch

import oracle
from oracle import ref_torch

torch.set_grad_enabled(True)


def rng(seed):
    return np.random.default_rng(seed)


CV_CASES = [(2, 7, 9, 3, 4), (1, 5, 6, 8, 2), (1, 12, 10, 5, 4), (1, 3, 4, 4, 1), (1, 20, 21, 2, 8)]


@pytest.mark.parametrize("B,H,W,C,d", CV_CASES)
def test_cost_volume_matches_composition(B, H, W, C, d):
    r = rng(1)
    prv = r.standard_normal((B, H, W, C))
    nxt = r.standard_normal((B, H, W, C))
    ref = ref_torch.cost_volume(torch.from_numpy(prv), torch.from_numpy(nxt), d).numpy()
    got64 = oracle.cost_volume(prv, nxt, d)
    np.testing.assert_allclose(got64, ref, rtol=0, atol=1e-14)
    got32 = oracle.cost_volume(prv.astype(np.float32), nxt.astype(np.float32), d)
    assert got32.dtype == np.float32
    np.testing.assert_allclose(got32, ref, rtol=0, atol=2e-6)


@pytest.mark.parametrize("B,H,W,C,d", CV_CASES[:4])
def test_cost_volume_bwd_matches_autograd(B, H, W, C, d):
    r = rng(2)
    prv = torch.from_numpy(r.standard_normal((B, H, W, C))).requires_grad_()
    nxt = torch.from_numpy(r.standard_normal((B, H, W, C))).requires_grad_()
    out = ref_torch.cost_volume(prv, nxt, d)
    g = torch.from_numpy(r.standard_normal(tuple(out.shape)))
    gp, gn = torch.autograd.grad(out, (prv, nxt), g)
    o = oracle.cost_volume(prv.detach().numpy(), nxt.detach().numpy(), d)
    gp2, gn2 = oracle.cost_volume_bwd(prv.detach().numpy(), nxt.detach().numpy(), o, g.numpy(), d)
    np.testing.assert_allclose(gp2, gp.numpy(), rtol=0, atol=1e-13)
    np.testing.assert_allclose(gn2, gn.numpy(), rtol=0, atol=1e-13)


def test_cost_volume_strided_output():
    r = rng(3)
    prv = r.standard_normal((1, 5, 6, 4)).astype(np.float32)
    nxt = r.standard_normal((1, 5, 6, 4)).astype(np.float32)
    dense = oracle.cost_volume(prv, nxt, 4)
    strided = oracle.cost_volume(prv, nxt, 4, out_stride=81 + 4 + 2)
    np.testing.assert_array_equal(strided[..., :81], dense)
    assert np.all(strided[..., 81:] == 0)


def test_cost_volume_integer_shift_kat():
    """nxt = prv shifted by (di,dj) => channel (di+d)*(2d+1)+(dj+d) holds mean(prv^2) > all others
    (channel order: row displacement outer, column inner; layers.py:80-81, vis.py:22-34)."""
    d, q = 4, 9
    r = rng(4)
    prv = r.standard_normal((1, 16, 18, 64)).astype(np.float32)   # C=64: self-match dominates
    for di, dj in [(0, 0), (-4, 4), (3, -2), (4, 4)]:
        nxt = np.zeros_like(prv)
        # nxt[i+di, j+dj] = prv[i, j]
        src = prv[:, max(0, -di):16 - max(0, di), max(0, -dj):18 - max(0, dj)]
        nxt[:, max(0, di):16 + min(0, di), max(0, dj):18 + min(0, dj)] = src
        out = oracle.cost_volume(prv, nxt, d)
        k = (di + d) * q + (dj + d)
        inner = out[0, 5:11, 5:13]
        assert np.all(inner.argmax(-1) == k)
        np.testing.assert_allclose(inner[..., k], (prv[0, 5:11, 5:13] ** 2).mean(-1), rtol=1e-6)


def test_cost_volume_leaky_slope():
    prv = np.ones((1, 2, 2, 1), np.float32)
    nxt = -np.ones((1, 2, 2, 1), np.float32)
    out = oracle.cost_volume(prv, nxt, 1)
    assert out[0, 0, 0, 4] == np.float32(-0.1)          # centre: -1 * 0.1
    assert out[0, 0, 0, 0] == 0.0                       # (-1,-1) is zero padding


WARP_CASES = [(2, 6, 7, 3), (1, 9, 5, 8), (1, 2, 2, 1), (1, 4, 11, 2)]


@pytest.mark.parametrize("mode", ["tf", "tfa"])
@pytest.mark.parametrize("B,H,W,C", WARP_CASES)
def test_warp_matches_composition(mode, B, H, W, C):
    r = rng(5)
    img = r.random((B, H, W, C))
    flow = r.standard_normal((B, H, W, 2)) * 3.0      # plenty of out-of-bounds samples
    ref = ref_torch.warp(torch.from_numpy(img), torch.from_numpy(flow), mode).numpy()
    np.testing.assert_allclose(oracle.warp(img, flow, mode), ref, rtol=0, atol=1e-14)
    i32, f32 = img.astype(np.float32), flow.astype(np.float32)
    ref32 = ref_torch.warp(torch.from_numpy(i32), torch.from_numpy(f32), mode).numpy()
    # same op order, no fused multiply-add on either side: bit-exact in fp32
    np.testing.assert_array_equal(oracle.warp(i32, f32, mode), ref32)


@pytest.mark.parametrize("mode", ["tf", "tfa"])
@pytest.mark.parametrize("B,H,W,C", WARP_CASES)
def test_warp_bwd_matches_autograd(mode, B, H, W, C):
    r = rng(6)
    img = torch.from_numpy(r.random((B, H, W, C))).requires_grad_()
    flow = torch.from_numpy(r.standard_normal((B, H, W, 2)) * 2.0).requires_grad_()
    out = ref_torch.warp(img, flow, mode)
    g = torch.from_numpy(r.standard_normal(tuple(out.shape)))
    gi, gf = torch.autograd.grad(out, (img, flow), g)
    gi2, gf2 = oracle.warp_bwd(img.detach().numpy(), flow.detach().numpy(), g.numpy(), mode)
    np.testing.assert_allclose(gi2, gi.numpy(), rtol=0, atol=1e-13)
    np.testing.assert_allclose(gf2, gf.numpy(), rtol=0, atol=1e-12)


def test_warp_tfa_zero_flow_gradient_tie_rule():
    """q - floor == 0 exactly => TF routes the Maximum tie to the constant: zero flow gradient
    (except where the floor is clamped to size-2, there q - floor == 1 passes)."""
    img = torch.from_numpy(rng(7).random((1, 4, 5, 2))).requires_grad_()
    flow = torch.zeros((1, 4, 5, 2), dtype=torch.float64, requires_grad=True)
    out = ref_torch.warp_tfa(img, flow)
    g = torch.ones_like(out)
    gi, gf = torch.autograd.grad(out, (img, flow), g)
    gi2, gf2 = oracle.warp_bwd(img.detach().numpy(), flow.detach().numpy(), g.numpy(), "tfa")
    np.testing.assert_allclose(gf2, gf.numpy(), atol=1e-14)
    np.testing.assert_allclose(gi2, gi.numpy(), atol=1e-14)
    assert np.all(gf2[0, :3, :4] == 0)


def test_warp_onehot_3x3_kat():
    """qpwcnet/app/optical_flow/test_warp.py:28-33: flow (x=+1, y=0) => out[i,j] = nxt[i,j+1]:
    the hot pixel moves from column 1 to column 0."""
    nxt = np.float32([[0, 0, 0], [0, 1, 0], [0, 0, 0]]).reshape(1, 3, 3, 1)
    flo = np.broadcast_to(np.float32([1, 0]).reshape(1, 1, 1, 2), (1, 3, 3, 2)).copy()
    out = oracle.warp(nxt, flo, "tfa")[0, :, :, 0]
    assert out[1, 0] == 1.0 and out[1, 1] == 0.0       # moved to column 0
    assert out[1, 2] == 0.0                            # x=3 clamps to the border column (value 0)
    out_tf = oracle.warp(nxt, flo, "tf")[0, :, :, 0]
    assert out_tf[1, 0] == 1.0 and out_tf[1, 1] == 0.0


def test_warp_zero_flow_kats():
    """zero flow => tfa mode is the identity; tf mode zeroes the last row/column."""
    img = rng(8).random((1, 5, 6, 3)).astype(np.float32)
    z = np.zeros((1, 5, 6, 2), np.float32)
    o2 = oracle.warp(img, z, "tfa")
    np.testing.assert_array_equal(o2[:, :-1, :-1], img[:, :-1, :-1])
    # last row/col: floor clamps to size-2, alpha == 1 => 1*(TR-TL)+TL, identity up to one rounding
    np.testing.assert_allclose(o2, img, rtol=2e-7, atol=0)
    o = oracle.warp(img, z, "tf")
    np.testing.assert_array_equal(o[:, :-1, :-1], img[:, :-1, :-1])
    assert np.all(o[:, -1] == 0) and np.all(o[:, :, -1] == 0)


def test_warp_far_oob_and_extrapolation_kats():
    img = rng(9).random((1, 4, 4, 2)).astype(np.float32)
    far = np.full((1, 4, 4, 2), 100.0, np.float32)
    assert np.all(oracle.warp(img, far, "tf") == 0)                      # fully outside => 0
    np.testing.assert_array_equal(oracle.warp(img, far, "tfa"),
                                  np.broadcast_to(img[:, -1:, -1:], img.shape))  # border replicate
    # x in (-1, 0): truncation gives x0 = 0, x1 = 1 => weights (1 - x) and x < 0 (extrapolation)
    fl = np.zeros((1, 4, 4, 2), np.float32)
    fl[0, 0, 0, 0] = -0.5
    o = oracle.warp(img, fl, "tf")
    np.testing.assert_allclose(o[0, 0, 0], 1.5 * img[0, 0, 0] - 0.5 * img[0, 0, 1], rtol=1e-6)


def test_warp_tfa_rejects_degenerate_grid():
    with pytest.raises(ValueError):
        oracle.warp(np.zeros((1, 1, 4, 2), np.float32), np.zeros((1, 1, 4, 2), np.float32), "tfa")


@pytest.mark.parametrize("mode", ["tf", "tfa"])
def test_fused_matches_composition_and_autograd(mode):
    r = rng(10)
    B, H, W, C, d = 1, 7, 8, 4, 4
    prv = torch.from_numpy(r.standard_normal((B, H, W, C))).requires_grad_()
    nxt = torch.from_numpy(r.standard_normal((B, H, W, C))).requires_grad_()
    flo = torch.from_numpy(r.standard_normal((B, H, W, 2)) * 2).requires_grad_()
    out = ref_torch.warp_cost_volume(prv, nxt, flo, mode, d)
    g = torch.from_numpy(r.standard_normal(tuple(out.shape)))
    gp, gn, gf = torch.autograd.grad(out, (prv, nxt, flo), g)
    a = [t.detach().numpy() for t in (prv, nxt, flo)]
    np.testing.assert_allclose(oracle.warp_cost_volume(*a, mode, d), out.detach().numpy(), atol=1e-14)
    gp2, gn2, gf2 = oracle.warp_cost_volume_bwd(*a, g.numpy(), mode, d)
    np.testing.assert_allclose(gp2, gp.numpy(), atol=1e-13)
    np.testing.assert_allclose(gn2, gn.numpy(), atol=1e-13)
    np.testing.assert_allclose(gf2, gf.numpy(), atol=1e-12)


# ---------------------------------------------------------------------------------------------
# committed golden fixtures (tests/golden/, minted by oracle/make_golden.py)
# ---------------------------------------------------------------------------------------------
import os

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_golden_fixtures_frozen():
    g = np.load(os.path.join(GOLD, "qpwc_golden.npz"))
    for name in ("cv_a", "cv_b", "cv_c", "cv_d"):
        d = int(g[f"{name}/d"])
        prv, nxt = g[f"{name}/prv"], g[f"{name}/nxt"]
        out32 = oracle.cost_volume(prv, nxt, d)
        scale = np.abs(g[f"{name}/out"]).max()
        assert np.abs(out32 - g[f"{name}/out"]).max() <= 1e-5 * scale
        gp, gn = oracle.cost_volume_bwd(prv, nxt, out32, g[f"{name}/g_out"], d)
        assert np.abs(gp - g[f"{name}/g_prv"]).max() <= 1e-5 * np.abs(g[f"{name}/g_prv"]).max()
        assert np.abs(gn - g[f"{name}/g_nxt"]).max() <= 1e-5 * np.abs(g[f"{name}/g_nxt"]).max()
    for name in ("warp_a", "warp_b", "warp_c"):
        for mode in oracle.MODES:
            out = oracle.warp(g[f"{name}/img"], g[f"{name}/flow"], mode)
            # fp32 vs fp64 arithmetic on the same fp32 inputs
            np.testing.assert_allclose(out, g[f"{name}/{mode}/out"], rtol=0, atol=3e-6)
            gi, gf = oracle.warp_bwd(g[f"{name}/img"], g[f"{name}/flow"], g[f"{name}/g_out"], mode)
            np.testing.assert_allclose(gi, g[f"{name}/{mode}/g_img"], rtol=0, atol=1e-5)
            np.testing.assert_allclose(gf, g[f"{name}/{mode}/g_flow"], rtol=0, atol=1e-5)


def test_golden_cfg1_reference_test_shape():
    """(4,32,64,3), d=4: the shape of test/test_cost_volume.py:20-21 and test/test_warp.py:24-25."""
    c = np.load(os.path.join(GOLD, "qpwc_cfg1.npz"))
    r1 = np.random.default_rng(int(c["seed"]))
    prv = r1.standard_normal((4, 32, 64, 3)).astype(np.float32)
    nxt = r1.standard_normal((4, 32, 64, 3)).astype(np.float32)
    img = r1.random((4, 32, 64, 3)).astype(np.float32)
    flo = r1.standard_normal((4, 32, 64, 2)).astype(np.float32)
    np.testing.assert_array_equal(prv[0, 0, :4], c["prv/head"])
    cv = oracle.cost_volume(prv, nxt, 4)
    assert cv.shape == (4, 32, 64, 81)
    np.testing.assert_allclose(cv[:, ::5, ::7], c["cv/sample"], rtol=0, atol=2e-6)
    assert abs(cv.astype(np.float64).sum() - float(c["cv/sum"])) < 1e-2
    for mode in oracle.MODES:
        w = oracle.warp(img, flo, mode)
        np.testing.assert_allclose(w[:, ::5, ::7], c[f"warp/{mode}/sample"], rtol=0, atol=2e-6)


# ------------------------------------------------------------------ x2 bilinear upsampling (Upsample)
def test_upsample2x_oracle_matches_torch_half_pixel_rule_and_adjoint():
    """oracle.upsample2x restates tf.image.resize(bilinear, half-pixel centres) x2 (Upsample,
    qpwcnet/core/non_layers.py:183-193).  TF is absent, so the restatement is cross-checked against
    torch's independent implementation of the same rule (align_corners=False), hand-derived values at
    the borders, and the adjoint identity <up(x), g> == <x, up_bwd(g)>."""
    import torch
    r = np.random.default_rng(5)
    x = r.standard_normal((2, 5, 7, 3))
    y = oracle.upsample2x(x, 2.0)
    t = torch.nn.functional.interpolate(torch.from_numpy(x).permute(0, 3, 1, 2), scale_factor=2, mode="bilinear",
                                        align_corners=False).permute(0, 2, 3, 1).numpy() * 2.0
    assert y.shape == (2, 10, 14, 3)
    np.testing.assert_allclose(y, t, rtol=0, atol=1e-13)
    # borders replicate, interior weights are 0.25 / 0.75
    np.testing.assert_allclose(y[:, 0, 0], 2.0 * x[:, 0, 0], atol=1e-15)
    np.testing.assert_allclose(y[:, -1, -1], 2.0 * x[:, -1, -1], atol=1e-15)
    np.testing.assert_allclose(y[:, 0, 1], 2.0 * (0.75 * x[:, 0, 0] + 0.25 * x[:, 0, 1]), atol=1e-14)
    np.testing.assert_allclose(y[:, 0, 2], 2.0 * (0.25 * x[:, 0, 0] + 0.75 * x[:, 0, 1]), atol=1e-14)
    g = r.standard_normal(y.shape)
    gi = oracle.upsample2x_bwd(g, 2.0)
    assert abs((y * g).sum() - (x * gi).sum()) < 1e-10
    x32 = x.astype(np.float32)
    assert np.abs(oracle.upsample2x(x32, 2.0) - y).max() < 2e-6


def test_oracle_reproduces_extras_golden():
    """tests/golden/qpwc_golden_extras.npz (oracle/make_golden.py --extras-only): Upsample, the
    half-flow warp pair and the warp on the upsampled coarse flow."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "qpwc_golden_extras.npz"))
    for name in ("up_a", "up_b"):
        x = g[f"{name}/x"].astype(np.float64)
        np.testing.assert_allclose(oracle.upsample2x(x, 2.0), g[f"{name}/out"], rtol=0, atol=1e-14)
        np.testing.assert_allclose(oracle.upsample2x_bwd(g[f"{name}/g_out"].astype(np.float64), 2.0), g[f"{name}/g_x"], rtol=0, atol=1e-13)
    for name in ("pair_a", "pair_b"):
        up = oracle.upsample2x(g[f"{name}/flow_coarse"].astype(np.float64), 2.0)
        for mode in oracle.MODES:
            np.testing.assert_allclose(oracle.warp(g[f"{name}/prv"].astype(np.float64), 0.5 * g[f"{name}/flo_10"].astype(np.float64), mode),
                                       g[f"{name}/{mode}/prv_w"], rtol=0, atol=1e-14)
            np.testing.assert_allclose(oracle.warp(g[f"{name}/nxt"].astype(np.float64), up, mode),
                                       g[f"{name}/{mode}/nxt_up_w"], rtol=0, atol=1e-14)


def test_occlusion_map_properties():
    """estimate_occlusion_map (occlusion.py:27-118) restated on the pinned tf-mode warp."""
    B, H, W = 2, 12, 17
    # zero flow: every pixel is hit by its own inverse flow and nothing leaves the image
    assert (oracle.occlusion_map(np.zeros((B, H, W, 2), np.float32)) == 0).all()
    # a uniform shift by (+3, +2): sources whose target leaves the image are flagged (oob) ...
    flow = np.zeros((B, H, W, 2), np.float32)
    flow[..., 0], flow[..., 1] = 3.0, 2.0
    m = oracle.occlusion_map(flow)
    assert set(np.unique(m)) <= {0.0, 1.0}
    assert (m[:, H - 2:, :] == 1).all() and (m[:, :, W - 3:] == 1).all()
    # ... and an explicit evaluation of the definition agrees pixel by pixel on a random flow
    r = np.random.default_rng(5)
    flow = (r.standard_normal((1, 6, 8, 2)) * 2).astype(np.float32)
    inv = -oracle.warp(flow, flow, "tf")
    want = np.ones((6, 8), np.float32)
    for i in range(6):
        for j in range(8):
            p = int(np.clip(np.trunc(np.float32(i) + inv[0, i, j, 1]), 0, 5))
            q = int(np.clip(np.trunc(np.float32(j) + inv[0, i, j, 0]), 0, 7))
            want[p, q] = 0
    for i in range(6):
        for j in range(8):
            i2, j2 = np.float32(i) + flow[0, i, j, 1], np.float32(j) + flow[0, i, j, 0]
            if i2 < 0 or i2 >= 6 or j2 < 0 or j2 >= 8:
                want[i, j] = 1
    np.testing.assert_array_equal(oracle.occlusion_map(flow)[0], want)


def test_oracle_reproduces_occlusion_golden():
    """tests/golden/qpwc_golden_occlusion.npz (oracle/make_golden.py --occlusion-only)."""
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "qpwc_golden_occlusion.npz"))
    for name in ("occ_a", "occ_b"):
        m = oracle.occlusion_map(g[f"{name}/flow"])
        np.testing.assert_array_equal(m, g[f"{name}/map"])
        assert 0.0 < m.mean() < 1.0                       # both classes present
