"""Parity at the FULL sizes of BASELINE.json's configurations, checked on crops: the op runs on the
whole (B, H, W, C) tensors on the GPU, the fp64 CPU oracle on crops of a few frame pairs cut with a
margin that covers every dependency of the crop's interior (search range, flow reach, scatter
reach); crops at image corners keep the border semantics (zero padding / clamping) in play.

config 2: 436x1024 (448x1024) B=8 -- 14x32x256, 28x64x256, 56x128x128, 112x256x64, 224x512x32
config 3: 256x448, per-GPU batch 8 -- 8x14x256 ... 128x224x32, FrameInterpolate warps with C = 3 ... 32
config 4: canonical PWC-Net sizes x C in {16..196}, d = 4 and d = 8
config 5: one 3840x2160 (2176-row) pair -- 68x120x256 ... 1088x1920x32, B=1
Forward AND backward, cost volume, warp and the UpFlow pair, both cost-volume engines."""
import numpy as np
import pytest
import torch

import oracle
from qpwcnet_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
MAXFLOW = 6.0
MARGIN = 20            # >= 2*d + MAXFLOW + 2 for d = 4: covers corr halo + warp reach + scatter reach


def gpu_inputs(B, H, W, C, seed, d=4):
    g = torch.Generator(device=DEV).manual_seed(seed)
    prv = torch.randn((B, H, W, C), device=DEV, generator=g)
    nxt = torch.randn((B, H, W, C), device=DEV, generator=g)
    flo = (torch.randn((B, H, W, 2), device=DEV, generator=g) * 2).clamp_(-MAXFLOW, MAXFLOW)
    D = (2 * d + 1) ** 2
    g_out = torch.randn((B, H, W, D), device=DEV, generator=g)
    return prv, nxt, flo, g_out


def crops(H, W, ch=14, cw=22):
    """(i0, i1, j0, j1) interiors: top-left corner, bottom-right corner, one interior window."""
    ch, cw = min(ch, H), min(cw, W)
    out = [(0, ch, 0, cw), (H - ch, H, W - cw, W)]
    if H > 3 * ch and W > 3 * cw:
        out.append((H // 2 - 3, H // 2 - 3 + ch, W // 3 + 5, W // 3 + 5 + cw))
    return out


def ext(i0, i1, j0, j1, H, W, m):
    """Crop extended by margin m, clipped at the image; returns the extended box and the interior's
    offset inside it."""
    a0, a1, b0, b1 = max(0, i0 - m), min(H, i1 + m), max(0, j0 - m), min(W, j1 + m)
    return a0, a1, b0, b1, i0 - a0, j0 - b0


def cut(t, b, box):
    a0, a1, b0, b1 = box
    return t[b:b + 1, a0:a1, b0:b1].detach().cpu().numpy().astype(np.float64)


def rel_ok(got, ref, tol=1e-5, what=""):
    err = float(np.abs(got - ref).max())
    assert err <= tol * max(float(np.abs(ref).max()), 1e-30), f"{what}: max|delta| {err:.3e} vs {tol:g} * {np.abs(ref).max():.3e}"


def full(t, b, dtype=np.float64):
    return t[b:b + 1].detach().cpu().numpy().astype(dtype)


def check_level(B, H, W, C, d, seed, engines=("auto", "ffma"), backward=True):
    """Cost volume: oracle on crops (the op is translation invariant).  Everything that goes through
    the warp: oracle on the WHOLE image of a frame pair -- the reference adds the flow to the absolute
    pixel index in fp32, so its rounding depends on where in the image a pixel sits."""
    prv, nxt, flo, g_out = gpu_inputs(B, H, W, C, seed, d)
    tp, tn, tf_ = prv.clone().requires_grad_(), nxt.clone().requires_grad_(), flo.clone().requires_grad_()
    margin = 2 * d + 2
    whole_bwd = backward and H * W * C * (2 * d + 1) ** 2 <= 1.3e9      # full-image oracle backward stays in seconds
    for eng in engines:
        ops.set_corr_engine(eng)
        cv = ops.cost_volume(tp, tn, d)
        fused = ops.warp_cost_volume(tp, tn, tf_, "tfa", d)
        if backward:
            gp, gn = torch.autograd.grad(cv, (tp, tn), g_out)
            fp, fn_, ff = torch.autograd.grad(fused, (tp, tn, tf_), g_out)
        for b in sorted({0, B - 1}):
            n32, f32 = full(nxt, b, np.float32), full(flo, b, np.float32)
            nw_full = oracle.warp(n32, f32, "tfa").astype(np.float64)      # fp32 arithmetic = the reference's
            for (i0, i1, j0, j1) in crops(H, W):
                a0, a1, b0, b1, oi, oj = ext(i0, i1, j0, j1, H, W, margin)
                box = (a0, a1, b0, b1)
                inner = (slice(None), slice(oi, oi + i1 - i0), slice(oj, oj + j1 - j0))
                p64, n64, g64 = (cut(t, b, box) for t in (prv, nxt, g_out))
                tag = f"{H}x{W}x{C} d={d} {eng} b={b} crop=({i0},{j0})"
                # ---- cost volume forward: max-norm and per element against the condition number
                ref = oracle.cost_volume(p64, n64, d)
                got = cut(cv, b, box)
                rel_ok(got[inner], ref[inner], what="cv fwd " + tag)
                pad = np.zeros((1, a1 - a0 + 2 * d, b1 - b0 + 2 * d, C))
                pad[:, d:d + a1 - a0, d:d + b1 - b0] = np.abs(n64)
                q = 2 * d + 1
                cond = np.stack([(np.abs(p64) * pad[:, u:u + a1 - a0, v:v + b1 - b0]).mean(-1) for u in range(q) for v in range(q)], -1)
                err = np.abs(got - ref)[inner]
                assert np.all(err <= 1e-5 * cond[inner] + 1e-30), f"cv fwd per-element {tag}: worst {np.max(err / (cond[inner] + 1e-30)):.2e}"
                # ---- UpFlow pair forward: exact cost volume of the fp32-warped frame
                reff = oracle.cost_volume(p64, nw_full[:, a0:a1, b0:b1], d)
                rel_ok(cut(fused, b, box)[inner], reff[inner], what="fused fwd " + tag)
                if backward:    # the leaky mask comes from the GPU's own forward output
                    rp, rn = oracle.cost_volume_bwd(p64, n64, got, g64, d)
                    rel_ok(cut(gp, b, box)[inner], rp[inner], what="cv g_prv " + tag)
                    rel_ok(cut(gn, b, box)[inner], rn[inner], what="cv g_nxt " + tag)
            if not whole_bwd:
                continue
            # ---- gradients of the pair, whole image: corr backward on the warped frame, then the warp's adjoint
            rpf, rnw = oracle.cost_volume_bwd(full(prv, b), nw_full, full(fused, b), full(g_out, b), d)
            rel_ok(full(fp, b), rpf, what=f"fused g_prv {H}x{W}x{C} {eng} b={b}")
            rnf, rff = oracle.warp_bwd(n32.astype(np.float64), f32.astype(np.float64), rnw, "tfa")
            rnf32, rff32 = oracle.warp_bwd(n32, f32, rnw.astype(np.float32), "tfa")
            for name, gpu, r64, r32 in (("g_nxt", full(fn_, b), rnf, rnf32), ("g_flow", full(ff, b), rff, rff32)):
                e_gpu = float(np.abs(gpu - r64).max())
                e_ref = float(np.abs(r32.astype(np.float64) - r64).max())
                assert e_gpu <= max(1e-5 * float(np.abs(r64).max()), 4 * e_ref), \
                    f"fused {name} {H}x{W}x{C} {eng} b={b}: gpu off exact by {e_gpu:.3e}, reference fp32 arithmetic by {e_ref:.3e}"
    ops.set_corr_engine("auto")


CFG2 = [(8, 14, 32, 256), (8, 28, 64, 256), (8, 56, 128, 128), (8, 112, 256, 64), (8, 224, 512, 32)]
CFG3 = [(8, 8, 14, 256), (8, 16, 28, 256), (8, 32, 56, 128), (8, 64, 112, 64), (8, 128, 224, 32)]
CFG4 = [(8, 218, 512, 16), (8, 109, 256, 32), (8, 55, 128, 64), (8, 28, 64, 96), (8, 14, 32, 128), (8, 7, 16, 196)]
CFG5 = [(1, 68, 120, 256), (1, 136, 240, 256), (1, 272, 480, 128), (1, 544, 960, 64), (1, 1088, 1920, 32)]


@pytest.mark.parametrize("B,H,W,C", CFG2)
def test_config2_levels_full_size(B, H, W, C):
    check_level(B, H, W, C, 4, seed=200 + C + H)


@pytest.mark.parametrize("B,H,W,C", CFG3)
def test_config3_levels_full_size(B, H, W, C):
    check_level(B, H, W, C, 4, seed=300 + C + H)


@pytest.mark.parametrize("B,H,W,C", CFG4)
def test_config4_levels_full_size(B, H, W, C):
    check_level(B, H, W, C, 4, seed=400 + C + H)


@pytest.mark.parametrize("B,H,W,C", [(8, 218, 512, 16), (8, 20, 30, 32), (2, 109, 256, 32)])
def test_config4_search_range_8(B, H, W, C):
    check_level(B, H, W, C, 8, seed=480 + C + H, engines=("auto",))


@pytest.mark.parametrize("B,H,W,C", CFG5)
def test_config5_levels_full_size(B, H, W, C):
    check_level(B, H, W, C, 4, seed=500 + C + H, engines=("auto",))


@pytest.mark.parametrize("mode", ["tf", "tfa"])
@pytest.mark.parametrize("B,H,W,C", [(8, 224, 512, 32), (8, 8, 14, 3), (8, 128, 224, 32), (8, 16, 28, 256), (1, 1088, 1920, 32)])
def test_warp_full_size(B, H, W, C, mode):
    """Stand-alone warp and FrameInterpolate's half-flow pair at full size, forward bit-exact on crops,
    gradients against exact arithmetic; prints the achieved max-abs errors."""
    g = torch.Generator(device=DEV).manual_seed(77 + C + H)
    img = torch.rand((B, H, W, C), device=DEV, generator=g)
    img2 = torch.rand((B, H, W, C), device=DEV, generator=g)
    flo = (torch.randn((B, H, W, 2), device=DEV, generator=g) * 2).clamp_(-MAXFLOW, MAXFLOW)
    flo2 = (torch.randn((B, H, W, 2), device=DEV, generator=g) * 2).clamp_(-MAXFLOW, MAXFLOW)
    gw = torch.randn((B, H, W, C), device=DEV, generator=g)
    ti, tf_ = img.clone().requires_grad_(), flo.clone().requires_grad_()
    out = ops.warp(ti, tf_, mode)
    gi, gf = torch.autograd.grad(out, (ti, tf_), gw)
    pair = ops.half_flow_warps(img, img2, flo, flo2, mode)      # [warp(prv, .5*flo_10) | warp(nxt, .5*flo_01)]
    worst = {"g_img": 0.0, "g_flow": 0.0}
    for b in sorted({0, B - 1}):      # whole images: the reference's fp32 coordinate arithmetic is position dependent
        i32, f32, i2, f2 = (full(t, b, np.float32) for t in (img, flo, img2, flo2))
        np.testing.assert_array_equal(full(out, b, np.float32), oracle.warp(i32, f32, mode))
        pr = full(pair, b, np.float32)
        np.testing.assert_array_equal(pr[..., :C], oracle.warp(i32, np.float32(0.5) * f2, mode))
        np.testing.assert_array_equal(pr[..., C:], oracle.warp(i2, np.float32(0.5) * f32, mode))
        g64 = full(gw, b)
        ri64, rf64 = oracle.warp_bwd(i32.astype(np.float64), f32.astype(np.float64), g64, mode)
        ri32, rf32 = oracle.warp_bwd(i32, f32, g64.astype(np.float32), mode)
        for name, got, r32, r64 in (("g_img", full(gi, b), ri32, ri64), ("g_flow", full(gf, b), rf32, rf64)):
            e_gpu = float(np.abs(got - r64).max())
            e_ref = float(np.abs(r32.astype(np.float64) - r64).max())
            worst[name] = max(worst[name], e_gpu)
            floor = 1e-6 * (max(1.0, C / 8) if name == "g_flow" else 1.0)
            assert e_gpu <= max(floor, 4 * e_ref), f"{name} {H}x{W}x{C} {mode}: gpu {e_gpu:.3e}, fp32 reference arithmetic {e_ref:.3e}"
    print(f"[achieved] warp bwd {H}x{W}x{C} {mode}: max|gpu - exact| g_img {worst['g_img']:.3e}  g_flow {worst['g_flow']:.3e}")
