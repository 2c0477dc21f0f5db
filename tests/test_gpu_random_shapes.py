"""Randomised shape sweep of the hot-path entry points against the oracle (GPU): ragged tiles, odd
heights/widths, channel tails, tiny maps, every kernel variant.  Seeds are fixed: the cases are the
same on every run."""
import os

import numpy as np
import pytest
import torch

import oracle
from qpwcnet_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(DEV)


def host(t):
    return t.detach().cpu().numpy()


def _shapes(seed, n, cmul):
    r = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        B = int(r.integers(1, 3))
        H = int(r.integers(1, 40))
        W = int(r.integers(1, 150))
        C = int(r.integers(1, 13)) * cmul
        out.append((B, H, W, C))
    return out


@pytest.mark.parametrize("engine", ["auto", "ffma"])
def test_random_cost_volume_forward(engine):
    ops.set_corr_engine(engine)
    for (B, H, W, C) in _shapes(11, 14, 4) + _shapes(12, 6, 1):
        r = np.random.default_rng(B * 7 + H * 13 + W * 17 + C)
        prv = r.standard_normal((B, H, W, C)).astype(np.float32)
        nxt = r.standard_normal((B, H, W, C)).astype(np.float32)
        ref = oracle.cost_volume(prv.astype(np.float64), nxt.astype(np.float64), 4)
        got = host(ops.cost_volume(dev(prv), dev(nxt), 4))
        assert np.abs(got - ref).max() <= 1e-5 * np.abs(ref).max(), (engine, B, H, W, C)
    ops.set_corr_engine("auto")


def test_random_cost_volume_backward():
    for (B, H, W, C) in _shapes(21, 12, 4) + _shapes(22, 4, 1):
        r = np.random.default_rng(B * 7 + H * 13 + W * 17 + C + 1)
        prv = r.standard_normal((B, H, W, C)).astype(np.float32)
        nxt = r.standard_normal((B, H, W, C)).astype(np.float32)
        tp, tn = dev(prv).requires_grad_(), dev(nxt).requires_grad_()
        out = ops.cost_volume(tp, tn, 4)
        g = r.standard_normal(tuple(out.shape)).astype(np.float32)
        gp, gn = torch.autograd.grad(out, (tp, tn), dev(g))
        rp, rn = oracle.cost_volume_bwd(prv.astype(np.float64), nxt.astype(np.float64),
                                        host(out).astype(np.float64), g.astype(np.float64), 4)
        for got, ref in ((gp, rp), (gn, rn)):
            assert np.abs(host(got) - ref).max() <= 1e-5 * max(np.abs(ref).max(), 1e-6), (B, H, W, C)


@pytest.mark.parametrize("mode", ["tf", "tfa"])
def test_random_warp_and_fused(mode):
    for (B, H, W, C) in _shapes(31, 10, 4) + _shapes(32, 6, 1):
        if mode == "tfa" and (H < 2 or W < 2):
            continue
        r = np.random.default_rng(B * 7 + H * 13 + W * 17 + C + 2)
        prv = r.standard_normal((B, H, W, C)).astype(np.float32)
        nxt = r.standard_normal((B, H, W, C)).astype(np.float32)
        flo = (r.standard_normal((B, H, W, 2)) * 3).astype(np.float32)
        np.testing.assert_array_equal(host(ops.warp(dev(nxt), dev(flo), mode)), oracle.warp(nxt, flo, mode))
        np.testing.assert_array_equal(host(ops.warp(dev(nxt), dev(flo), mode, flow_scale=0.5)),
                                      oracle.warp(nxt, np.float32(0.5) * flo, mode))
        pair = host(ops.half_flow_warps(dev(prv), dev(nxt), dev(flo), dev(-flo), mode))
        np.testing.assert_array_equal(pair[..., :C], oracle.warp(prv, np.float32(-0.5) * flo, mode))
        np.testing.assert_array_equal(pair[..., C:], oracle.warp(nxt, np.float32(0.5) * flo, mode))
        ref = oracle.warp_cost_volume(prv.astype(np.float64), nxt.astype(np.float64), flo.astype(np.float64), mode, 4)
        got = host(ops.warp_cost_volume(dev(prv), dev(nxt), dev(flo), mode, 4))
        assert np.abs(got - ref).max() <= max(1e-5 * np.abs(ref).max(), 1e-6), (mode, B, H, W, C)   # floor: fp32 cancellation residue of degenerate (W == 1) warps
        if H % 2 == 0 and W % 2 == 0:
            fc = (r.standard_normal((B, H // 2, W // 2, 2)) * 2).astype(np.float32)
            up = oracle.upsample2x(fc, 2.0)
            np.testing.assert_array_equal(host(ops.upsample2x(dev(fc), 2.0)), up)
            np.testing.assert_array_equal(host(ops.warp_up(dev(nxt), dev(fc), mode)), oracle.warp(nxt, up, mode))
            refu = oracle.warp_cost_volume(prv.astype(np.float64), nxt.astype(np.float64), up.astype(np.float64), mode, 4)
            gotu = host(ops.warp_cost_volume_up(dev(prv), dev(nxt), dev(fc), mode, 4))
            assert np.abs(gotu - refu).max() <= max(1e-5 * np.abs(refu).max(), 1e-6), (mode, B, H, W, C)
