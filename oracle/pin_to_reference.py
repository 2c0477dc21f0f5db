#!/usr/bin/env python3
"""Pin the oracle by EXECUTING THE REFERENCE'S OWN SOURCE  (TEST INFRASTRUCTURE, runs in the
authoring container only -- /root/reference does not exist on the GPU box).

    python oracle/pin_to_reference.py            # writes tests/golden/ref_*.npz
    python oracle/pin_to_reference.py --check    # re-runs and compares with the committed files

The unmodified modules /root/reference/qpwcnet/core/{warp,layers,non_layers,occlusion}.py are
imported by path and run under `oracle/_ref_shim` (a `tensorflow` stand-in over torch-CPU that
implements only the tf.* primitives those files call).  Outputs are the reference's fp32 results;
gradients are torch autograd's derivative of the very graph the reference code builds (what TF
autodiff derives: gather_nd -> scatter-add, slice -> zero-pad, casts/clips carry no gradient).  Each
case is also run in `exact` mode (tf.float32 := fp64) to record how far the reference's own fp32
arithmetic is from exact arithmetic -- the yardstick for order-dependent sums.

What this pins (by execution):  CostVolume (Keras layer and functor; CostVolumeV2 through the
reference's own "0.0" equivalence, app/test/test_cvol_equal.py:25), tf_warp / Warp,
estimate_occlusion_map, the UpFlow composition with Warp, FrameInterpolate's 0.5*flow warps, and the
glue of WarpV2 (sign flip, channel reversal, NCHW transposes).  What it cannot pin: the arithmetic
inside tensorflow_addons (`interpolate_bilinear`) and TensorFlow's resize kernel -- restated, marked
"tfa"/unpinned in the fixtures.
"""
import os
import sys
import warnings

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("QPWC_REFERENCE", "/root/reference")
GOLD = os.path.join(ROOT, "tests", "golden")


def _import_reference():
    """Import the reference modules under the shim; our own `qpwcnet` alias package must not win."""
    for m in [m for m in sys.modules if m == "qpwcnet" or m.startswith("qpwcnet.")]:
        del sys.modules[m]
    sys.path[:] = [p for p in sys.path if os.path.abspath(p or ".") != ROOT]
    sys.path[0:0] = [os.path.join(HERE, "_ref_shim"), REF]
    warnings.simplefilter("ignore", SyntaxWarning)
    import tensorflow as tf
    from qpwcnet.core import layers, non_layers, occlusion, warp
    for mod in (layers, non_layers, occlusion, warp):
        assert os.path.abspath(mod.__file__).startswith(os.path.abspath(REF) + os.sep), mod.__file__
    return tf, warp, layers, non_layers, occlusion


import contextlib  # noqa: E402
import io  # noqa: E402

import numpy as np  # noqa: E402
import torch  # noqa: E402

tf, ref_warp, ref_layers, ref_non_layers, ref_occlusion = _import_reference()


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def to_fmt(a, fmt):
    return np.ascontiguousarray(a.transpose(0, 3, 1, 2)) if fmt == "channels_first" else a


def from_fmt(a, fmt):
    return np.ascontiguousarray(a.transpose(0, 2, 3, 1)) if fmt == "channels_first" and a.ndim == 4 else a


def run(fn, inputs, g_out, fmt, exact=False):
    """Run reference callable `fn(*tensors)` on NHWC numpy `inputs` in data format `fmt`; returns the
    output and the gradients wrt every input (NHWC, numpy) for the upstream gradient `g_out`."""
    tf.keras.backend.set_image_data_format(fmt)
    tf.set_real(torch.float64 if exact else torch.float32)
    try:
        ts = [tf.constant(to_fmt(a.astype(np.float64) if exact else a, fmt)).requires_grad_() for a in inputs]
        with contextlib.redirect_stdout(io.StringIO()):       # tf_warp prints its shape
            out = fn(*ts)
        res = [from_fmt(out.numpy(), fmt)]
        if g_out is not None:
            g = tf.constant(to_fmt(g_out.astype(np.float64) if exact else g_out, fmt))
            grads = torch.autograd.grad(out, ts, g, allow_unused=True)
            res += [from_fmt(gr.numpy(), fmt) for gr in grads]
        return res
    finally:
        tf.set_real(torch.float32)
        tf.keras.backend.set_image_data_format("channels_last")


def cv_layer(d):
    return lambda prv, nxt: ref_layers.CostVolume(d)((prv, nxt))                    # layers.py:32-109


def cv_functor(d):
    return lambda prv, nxt: ref_non_layers.CostVolume(d)((prv, nxt))                # non_layers.py:51-104


def warp_layer():
    return lambda img, flo: ref_layers.Warp()((img, flo))                           # layers.py:144-168


def warp_fn(fmt):
    return lambda img, flo: ref_warp.tf_warp(img, flo, fmt)                         # warp.py:63-153


def warpv2_layer():
    return lambda img, flo: ref_layers.WarpV2()((img, flo))                         # layers.py:171-186


def upflow_pair(d):
    """UpFlow's composition (non_layers.py:377-380) with the in-repo flavours of both ops."""
    return lambda prv, nxt, flo: ref_non_layers.CostVolume(d)((prv, ref_non_layers.Warp()((nxt, flo))))


def half_warp():
    return lambda img, flo: ref_non_layers.Warp()((img, 0.5 * flo))                 # non_layers.py:303-304


def mint():
    fx = {}
    r = np.random.default_rng(20261101)

    def record(name, fn, inputs, names, g_shape_from_out=True, formats=("channels_last", "channels_first"),
               store_inputs=True, sample=None):
        out, = run(fn, inputs, None, "channels_last")
        g = f32(r.standard_normal(out.shape))
        res = run(fn, inputs, g, "channels_last")
        res64 = run(fn, inputs, g, "channels_last", exact=True)
        for fmt in formats[1:]:
            alt = run(fn, inputs, g, fmt)
            for a, b in zip(res, alt):      # the reference's two layouts agree up to summation order
                assert np.abs(a - b).max() <= 2e-6 * max(1.0, np.abs(a).max()), (name, fmt, np.abs(a - b).max())
        pick = (lambda a: a) if sample is None else (lambda a: np.ascontiguousarray(a[sample]))
        if store_inputs:
            for n, a in zip(names, inputs):
                fx[f"{name}/{n}"] = a
            fx[f"{name}/g_out"] = g
        fx[f"{name}/out"] = pick(res[0])
        fx[f"{name}/out_exact"] = pick(res64[0])
        for n, a, a64 in zip(names, res[1:], res64[1:]):
            fx[f"{name}/g_{n}"] = pick(a)
            fx[f"{name}/g_{n}_exact"] = pick(a64)

    # ---- cost volume (layers.py:72-100): small full cases, d = 1, 2, 4, 8; C%4 != 0; C = 1
    for name, (B, H, W, C, d) in {"cv_a": (2, 8, 10, 3, 4), "cv_b": (1, 6, 7, 8, 2), "cv_c": (1, 9, 12, 32, 4),
                                  "cv_d": (1, 10, 11, 5, 8), "cv_e": (1, 3, 2, 1, 1), "cv_f": (1, 12, 20, 64, 4)}.items():
        prv, nxt = f32(r.standard_normal((B, H, W, C))), f32(r.standard_normal((B, H, W, C)))
        fx[f"{name}/d"] = np.int32(d)
        record(name, cv_layer(d), [prv, nxt], ["prv", "nxt"])
        o2, = run(cv_functor(d), [prv, nxt], None, "channels_last")
        assert np.array_equal(o2, fx[f"{name}/out"])            # Keras layer == functor
    # ---- tf_warp (warp.py:63-153): noise flows of growing size, incl. far out-of-bounds
    for name, (B, H, W, C, s) in {"warp_a": (2, 8, 10, 3, 1.0), "warp_b": (1, 7, 9, 8, 3.0), "warp_c": (1, 5, 6, 2, 6.0),
                                  "warp_d": (1, 6, 5, 32, 0.3), "warp_e": (1, 2, 2, 1, 1.0)}.items():
        img, flo = f32(r.random((B, H, W, C))), f32(r.standard_normal((B, H, W, 2)) * s)
        record(name, warp_layer(), [img, flo], ["img", "flow"])
        o2, = run(warp_fn("channels_last"), [img, flo], None, "channels_last")
        assert np.array_equal(o2, fx[f"{name}/out"])
        # WarpV2: the reference's glue around tfa executed, tfa arithmetic restated (UNPINNED)
        record(name + "/tfa", warpv2_layer(), [img, flo], ["img", "flow"], store_inputs=False)
    # ---- UpFlow composition and FrameInterpolate's half-flow warp, in-repo flavours
    for name, (B, H, W, C, d) in {"fused_a": (1, 9, 11, 8, 4), "fused_b": (2, 6, 7, 3, 4)}.items():
        prv, nxt = f32(r.standard_normal((B, H, W, C))), f32(r.standard_normal((B, H, W, C)))
        flo = f32(r.standard_normal((B, H, W, 2)) * 2.0)
        fx[f"{name}/d"] = np.int32(d)
        record(name, upflow_pair(d), [prv, nxt, flo], ["prv", "nxt", "flow"])
    for name, (B, H, W, C) in {"half_a": (1, 8, 10, 4), "half_b": (2, 6, 6, 3)}.items():
        img, flo = f32(r.random((B, H, W, C))), f32(r.standard_normal((B, H, W, 2)) * 3)
        record(name, half_warp(), [img, flo], ["img", "flow"])
    # ---- estimate_occlusion_map (occlusion.py:27-118)
    for name, (B, H, W, s) in {"occ_a": (2, 9, 13, 2.0), "occ_b": (1, 16, 12, 7.0), "occ_c": (1, 4, 4, 0.0)}.items():
        flow = f32(r.standard_normal((B, H, W, 2)) * s)
        fx[f"{name}/flow"] = flow
        for fmt in ("channels_last", "channels_first"):
            m, = run(lambda f: ref_occlusion.estimate_occlusion_map(f), [flow], None, fmt)
            if fmt == "channels_last":
                fx[f"{name}/map"] = m
            else:
                assert np.array_equal(m, fx[f"{name}/map"])
    # ---- known-answer cases through the reference code
    one_hot = f32([[0, 0, 0], [0, 1, 0], [0, 0, 0]]).reshape(1, 3, 3, 1)             # app/optical_flow/test_warp.py:25-33
    flo = np.broadcast_to(f32([1, 0]).reshape(1, 1, 1, 2), (1, 3, 3, 2)).copy()
    fx["kat_onehot/img"], fx["kat_onehot/flow"] = one_hot, flo
    fx["kat_onehot/tf"], = run(warp_layer(), [one_hot, flo], None, "channels_last")
    fx["kat_onehot/tfa"], = run(warpv2_layer(), [one_hot, flo], None, "channels_last")
    img = f32(r.random((1, 5, 6, 2)))
    for kname, flow in {"zero": np.zeros((1, 5, 6, 2)), "far": np.full((1, 5, 6, 2), 100.0),
                        "neg_far": np.full((1, 5, 6, 2), -100.0), "extrap": np.full((1, 5, 6, 2), -0.25),
                        "int_shift": np.broadcast_to(f32([2, -1]), (1, 5, 6, 2))}.items():
        flow = f32(flow)
        fx[f"kat_{kname}/img"], fx[f"kat_{kname}/flow"] = img, flow
        fx[f"kat_{kname}/tf"], = run(warp_layer(), [img, flow], None, "channels_last")
        fx[f"kat_{kname}/tfa"], = run(warpv2_layer(), [img, flow], None, "channels_last")
    np.savez_compressed(os.path.join(GOLD, "ref_pin.npz"), **fx)

    # ---- config 1 = the reference's own test shapes, seeded: (4,32,64,3), d=4
    # (test/test_cost_volume.py:20-21, test/test_warp.py:24-25); inputs regenerated from the seed,
    # outputs stored as strided samples + full-tensor sums (fp64 accumulation).
    cfg1 = {"seed": np.int32(1)}
    r1 = np.random.default_rng(1)
    prv, nxt = f32(r1.standard_normal((4, 32, 64, 3))), f32(r1.standard_normal((4, 32, 64, 3)))
    img, flo = f32(r1.random((4, 32, 64, 3))), f32(r1.standard_normal((4, 32, 64, 2)))
    g_cv = f32(r1.standard_normal((4, 32, 64, 81)))
    g_w = f32(r1.standard_normal((4, 32, 64, 3)))
    cfg1["prv/head"] = prv[0, 0, :4]
    samp = (slice(None), slice(None, None, 5), slice(None, None, 7))
    for fmt in ("channels_last", "channels_first"):
        cv, gp, gn = run(cv_layer(4), [prv, nxt], g_cv, fmt)
        w, gi, gf = run(warp_layer(), [img, flo], g_w, fmt)
        w2, gi2, gf2 = run(warpv2_layer(), [img, flo], g_w, fmt)
        cur = {"cv/sample": cv[samp], "cv/sum": cv.sum(dtype=np.float64),
               "cv/g_prv": gp[samp], "cv/g_nxt": gn[samp],
               "warp/tf/sample": w[samp], "warp/tf/sum": w.sum(dtype=np.float64),
               "warp/tf/g_img": gi[samp], "warp/tf/g_flow": gf[samp],
               "warp/tfa/sample": w2[samp], "warp/tfa/sum": w2.sum(dtype=np.float64),
               "warp/tfa/g_img": gi2[samp], "warp/tfa/g_flow": gf2[samp]}
        if fmt == "channels_last":
            cfg1.update(cur)
        else:
            for k, v in cur.items():
                assert np.abs(v - cfg1[k]).max() <= 2e-6 * max(1.0, np.abs(v).max()), k
    # app/test/test_cvol_equal.py:10: (1,128,256,3) both layouts
    r2 = np.random.default_rng(2)
    prv, nxt = f32(r2.standard_normal((1, 128, 256, 3))), f32(r2.standard_normal((1, 128, 256, 3)))
    cv, = run(cv_layer(4), [prv, nxt], None, "channels_last")
    cvf, = run(cv_layer(4), [prv, nxt], None, "channels_first")
    assert np.abs(cv - cvf).max() <= 2e-6 * np.abs(cv).max()
    cfg1["cvol_equal/sample"] = cv[:, ::9, ::11]
    cfg1["cvol_equal/sum"] = cv.sum(dtype=np.float64)
    np.savez_compressed(os.path.join(GOLD, "ref_cfg1.npz"), **cfg1)
    return {"ref_pin.npz": fx, "ref_cfg1.npz": cfg1}


def check():
    """Re-execute the reference and compare with the committed fixtures (bit for bit)."""
    import tempfile
    global GOLD
    committed = GOLD
    with tempfile.TemporaryDirectory() as tmp:
        GOLD = tmp
        fresh = mint()
        bad = 0
        for fname, fx in fresh.items():
            old = np.load(os.path.join(committed, fname))
            assert sorted(old.files) == sorted(fx), fname
            for k in old.files:
                if not np.array_equal(old[k], np.asarray(fx[k])):
                    print("MISMATCH", fname, k)
                    bad += 1
    GOLD = committed
    print("reference re-executed: fixtures", "identical" if not bad else f"{bad} mismatches")
    return bad


if __name__ == "__main__":
    if "--check" in sys.argv:
        sys.exit(1 if check() else 0)
    os.makedirs(GOLD, exist_ok=True)
    mint()
    for f in ("ref_pin.npz", "ref_cfg1.npz"):
        print(f, os.path.getsize(os.path.join(GOLD, f)))
