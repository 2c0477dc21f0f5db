#!/usr/bin/env python3
"""Mint the committed golden fixtures under tests/golden/  (TEST INFRASTRUCTURE).

The reference ships no golden vectors and cannot be executed here (TensorFlow / tensorflow_addons
are absent, no network), so the fixtures are produced by the CPU oracle (fp64 arithmetic on
fp32-representable inputs, then stored as fp64) -- they freeze the oracle's behaviour and give the
GPU tests a file-based target that does not need gcc on the box.  Re-run:  python oracle/make_golden.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def f32(a):
    return np.asarray(a, dtype=np.float32)


def main():
    os.makedirs(OUT, exist_ok=True)
    r = np.random.default_rng(20261018)
    fx = {}
    # --- cost volume: (B,H,W,C,d) incl. the reference's own test shape family (C=3, d=4)
    for name, (B, H, W, C, d) in {
        "cv_a": (2, 8, 10, 3, 4), "cv_b": (1, 6, 7, 8, 2), "cv_c": (1, 9, 12, 32, 4),
        "cv_d": (1, 10, 11, 5, 8),
    }.items():
        prv, nxt = f32(r.standard_normal((B, H, W, C))), f32(r.standard_normal((B, H, W, C)))
        out = oracle.cost_volume(prv.astype(np.float64), nxt.astype(np.float64), d)
        g = f32(r.standard_normal(out.shape))
        gp, gn = oracle.cost_volume_bwd(prv.astype(np.float64), nxt.astype(np.float64), out,
                                        g.astype(np.float64), d)
        fx.update({f"{name}/prv": prv, f"{name}/nxt": nxt, f"{name}/d": np.int32(d),
                   f"{name}/out": out, f"{name}/g_out": g, f"{name}/g_prv": gp, f"{name}/g_nxt": gn})
    # --- warp, both border rules
    for name, (B, H, W, C, s) in {
        "warp_a": (2, 8, 10, 3, 1.0), "warp_b": (1, 7, 9, 8, 3.0), "warp_c": (1, 5, 6, 2, 6.0),
    }.items():
        img, flo = f32(r.random((B, H, W, C))), f32(r.standard_normal((B, H, W, 2)) * s)
        g = f32(r.standard_normal(img.shape))
        fx.update({f"{name}/img": img, f"{name}/flow": flo, f"{name}/g_out": g})
        for mode in oracle.MODES:
            out = oracle.warp(img.astype(np.float64), flo.astype(np.float64), mode)
            gi, gf = oracle.warp_bwd(img.astype(np.float64), flo.astype(np.float64),
                                     g.astype(np.float64), mode)
            fx.update({f"{name}/{mode}/out": out, f"{name}/{mode}/g_img": gi,
                       f"{name}/{mode}/g_flow": gf})
    # --- fused warp -> cost volume (UpFlow), non_layers.py:377-380
    for name, (B, H, W, C, d) in {"fused_a": (1, 9, 11, 8, 4), "fused_b": (2, 6, 7, 3, 4)}.items():
        prv, nxt = f32(r.standard_normal((B, H, W, C))), f32(r.standard_normal((B, H, W, C)))
        flo = f32(r.standard_normal((B, H, W, 2)) * 2.0)
        g = f32(r.standard_normal((B, H, W, (2 * d + 1) ** 2)))
        fx.update({f"{name}/prv": prv, f"{name}/nxt": nxt, f"{name}/flow": flo, f"{name}/g_out": g,
                   f"{name}/d": np.int32(d)})
        a64 = [t.astype(np.float64) for t in (prv, nxt, flo)]
        for mode in oracle.MODES:
            out = oracle.warp_cost_volume(*a64, mode, d)
            gp, gn, gf = oracle.warp_cost_volume_bwd(*a64, g.astype(np.float64), mode, d)
            fx.update({f"{name}/{mode}/out": out, f"{name}/{mode}/g_prv": gp,
                       f"{name}/{mode}/g_nxt": gn, f"{name}/{mode}/g_flow": gf})
    np.savez_compressed(os.path.join(OUT, "qpwc_golden.npz"), **fx)

    # --- config 1 (test/test_cost_volume.py:20-21, test/test_warp.py:24-25): (4,32,64,3), d=4.
    # Inputs are regenerated from the seed; only strided samples + sums of the outputs are stored.
    r1 = np.random.default_rng(1)
    prv, nxt = f32(r1.standard_normal((4, 32, 64, 3))), f32(r1.standard_normal((4, 32, 64, 3)))
    img, flo = f32(r1.random((4, 32, 64, 3))), f32(r1.standard_normal((4, 32, 64, 2)))
    cfg1 = {"seed": np.int32(1)}
    cv = oracle.cost_volume(prv.astype(np.float64), nxt.astype(np.float64), 4)
    cfg1["cv/sample"] = cv[:, ::5, ::7]
    cfg1["cv/sum"] = cv.sum()
    cfg1["prv/head"] = prv[0, 0, :4]          # guards the RNG stream
    for mode in oracle.MODES:
        w = oracle.warp(img.astype(np.float64), flo.astype(np.float64), mode)
        cfg1[f"warp/{mode}/sample"] = w[:, ::5, ::7]
        cfg1[f"warp/{mode}/sum"] = w.sum()
    np.savez_compressed(os.path.join(OUT, "qpwc_cfg1.npz"), **cfg1)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


def extras():
    """Fixtures for the entry points added on top of the first set (own file and RNG stream, so the
    first file stays byte-identical): Upsample (non_layers.py:183-193), FrameInterpolate's half-flow
    warp pair (non_layers.py:303-311) and the warp fed by the upsampled coarse flow (pwcnet.py:49-56)."""
    r = np.random.default_rng(20261019)
    fx = {}
    for name, (B, H, W, C) in {"up_a": (2, 5, 7, 2), "up_b": (1, 4, 6, 3)}.items():
        x = f32(r.standard_normal((B, H, W, C)))
        y = oracle.upsample2x(x.astype(np.float64), 2.0)
        g = f32(r.standard_normal(y.shape))
        fx.update({f"{name}/x": x, f"{name}/out": y, f"{name}/g_out": g,
                   f"{name}/g_x": oracle.upsample2x_bwd(g.astype(np.float64), 2.0)})
    for name, (B, H, W, C) in {"pair_a": (1, 8, 10, 4), "pair_b": (2, 6, 6, 3)}.items():
        prv, nxt = f32(r.random((B, H, W, C))), f32(r.random((B, H, W, C)))
        f01, f10 = f32(r.standard_normal((B, H, W, 2)) * 3), f32(r.standard_normal((B, H, W, 2)) * 3)
        fc = f32(r.standard_normal((B, H // 2, W // 2, 2)) * 1.5)
        fx.update({f"{name}/prv": prv, f"{name}/nxt": nxt, f"{name}/flo_01": f01, f"{name}/flo_10": f10,
                   f"{name}/flow_coarse": fc})
        up = oracle.upsample2x(fc.astype(np.float64), 2.0)
        for mode in oracle.MODES:
            fx[f"{name}/{mode}/prv_w"] = oracle.warp(prv.astype(np.float64), 0.5 * f10.astype(np.float64), mode)
            fx[f"{name}/{mode}/nxt_w"] = oracle.warp(nxt.astype(np.float64), 0.5 * f01.astype(np.float64), mode)
            fx[f"{name}/{mode}/nxt_up_w"] = oracle.warp(nxt.astype(np.float64), up, mode)
    np.savez_compressed(os.path.join(OUT, "qpwc_golden_extras.npz"), **fx)


def occlusion():
    """estimate_occlusion_map (occlusion.py:27-118) fixtures, own file and RNG stream."""
    r = np.random.default_rng(20261020)
    fx = {}
    for name, (B, H, W, sigma) in {"occ_a": (2, 9, 13, 2.0), "occ_b": (1, 16, 12, 7.0)}.items():
        flow = f32(r.standard_normal((B, H, W, 2)) * sigma)
        fx.update({f"{name}/flow": flow, f"{name}/map": oracle.occlusion_map(flow)})
    np.savez_compressed(os.path.join(OUT, "qpwc_golden_occlusion.npz"), **fx)


if __name__ == "__main__":
    if "--occlusion-only" in sys.argv:
        occlusion()
        sys.exit(0)
    if "--extras-only" not in sys.argv:
        main()
    extras()
    occlusion()
