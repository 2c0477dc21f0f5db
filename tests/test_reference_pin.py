"""The oracle against fixtures minted by EXECUTING the reference's own source
(oracle/pin_to_reference.py: unmodified /root/reference/qpwcnet/core/{warp,layers,non_layers,
occlusion}.py under the tf shim).  This is what pins the oracle; tests/test_gpu_parity.py then
compares the CUDA library with the same fixtures and with the oracle at larger sizes.

Criteria: tf_warp forward and the occlusion map bit-exact (elementwise fp32 op sequences); the cost
volume <= 1e-6 relative in fp32 (the channel sum's order is not fixed by the reference) and
<= 1e-13 against the `exact` (fp64) run; gradients <= 1e-6 absolute-or-relative in fp32, <= 1e-12 in
fp64.  `*/tfa` entries ran the reference's WarpV2 glue around a restated tensorflow_addons
(arithmetic unpinned, the sign / channel-order / transpose handling pinned)."""
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PIN = np.load(os.path.join(GOLD, "ref_pin.npz"))
CFG1 = np.load(os.path.join(GOLD, "ref_cfg1.npz"))


def names(prefix):
    return sorted({k.split("/")[0] for k in PIN.files if k.startswith(prefix)})


def close(got, ref, tol):
    err = float(np.abs(got.astype(np.float64) - ref.astype(np.float64)).max()) if got.size else 0.0
    scale = max(1.0, float(np.abs(ref).max())) if ref.size else 1.0
    assert got.shape == ref.shape
    assert err <= tol * scale, f"max|delta| = {err:.3e} > {tol:g} * {scale:.3g}"


def as_accurate(got, ref32, exact, tol=1e-6, factor=4.0):
    """For sums whose fp32 value depends on the summation order (scatter-add image gradients, flow
    gradients summed over channels and weighted by extrapolation weights ~|flow|): the candidate may
    be off EXACT arithmetic by `tol` (abs-or-rel) or by `factor` x what the reference's own fp32 run
    is off, whichever is larger."""
    e_got = float(np.abs(got.astype(np.float64) - exact).max())
    e_ref = float(np.abs(ref32.astype(np.float64) - exact).max())
    scale = max(1.0, float(np.abs(exact).max()))
    assert e_got <= max(tol * scale, factor * e_ref), f"candidate off exact by {e_got:.3e}, reference fp32 by {e_ref:.3e}"


@pytest.mark.parametrize("name", names("cv_"))
def test_cost_volume_reproduces_reference(name):
    prv, nxt, g, d = (PIN[f"{name}/{k}"] for k in ("prv", "nxt", "g_out", "d"))
    d = int(d)
    out32 = oracle.cost_volume(prv, nxt, d)
    close(out32, PIN[f"{name}/out"], 1e-6)
    out64 = oracle.cost_volume(prv.astype(np.float64), nxt.astype(np.float64), d)
    close(out64, PIN[f"{name}/out_exact"], 1e-13)
    gp, gn = oracle.cost_volume_bwd(prv, nxt, out32, g, d)
    close(gp, PIN[f"{name}/g_prv"], 1e-6)
    close(gn, PIN[f"{name}/g_nxt"], 1e-6)
    gp, gn = oracle.cost_volume_bwd(prv.astype(np.float64), nxt.astype(np.float64), out64, g.astype(np.float64), d)
    close(gp, PIN[f"{name}/g_prv_exact"], 1e-12)
    close(gn, PIN[f"{name}/g_nxt_exact"], 1e-12)


@pytest.mark.parametrize("name", names("warp_") + names("half_"))
def test_warp_reproduces_reference(name):
    img, flow, g = (PIN[f"{name}/{k}"] for k in ("img", "flow", "g_out"))
    s = np.float32(0.5) if name.startswith("half_") else np.float32(1.0)     # non_layers.py:303-304
    fl = (s * flow).astype(np.float32)
    out = oracle.warp(img, fl, "tf")
    np.testing.assert_array_equal(out, PIN[f"{name}/out"])                   # bit-exact
    gi, gf = oracle.warp_bwd(img, fl, g, "tf")
    as_accurate(gi, PIN[f"{name}/g_img"], PIN[f"{name}/g_img_exact"])
    as_accurate(gf * s, PIN[f"{name}/g_flow"], PIN[f"{name}/g_flow_exact"])
    i64, f64, g64 = img.astype(np.float64), fl.astype(np.float64), g.astype(np.float64)
    close(oracle.warp(i64, f64, "tf"), PIN[f"{name}/out_exact"], 1e-13)
    gi, gf = oracle.warp_bwd(i64, f64, g64, "tf")
    close(gi, PIN[f"{name}/g_img_exact"], 1e-12)
    close(gf * float(s), PIN[f"{name}/g_flow_exact"], 1e-12)
    if f"{name}/tfa/out" in PIN.files:      # WarpV2 glue (layers.py:177-186) around restated tfa
        np.testing.assert_array_equal(oracle.warp(img, fl, "tfa"), PIN[f"{name}/tfa/out"])
        gi, gf = oracle.warp_bwd(img, fl, g2 := PIN[f"{name}/g_out"], "tfa")
        # the tfa run drew its own upstream gradient
        assert g2.shape == img.shape


@pytest.mark.parametrize("name", names("fused_"))
def test_upflow_pair_reproduces_reference(name):
    prv, nxt, flow, g, d = (PIN[f"{name}/{k}"] for k in ("prv", "nxt", "flow", "g_out", "d"))
    d = int(d)
    close(oracle.warp_cost_volume(prv, nxt, flow, "tf", d), PIN[f"{name}/out"], 1e-6)
    gp, gn, gf = oracle.warp_cost_volume_bwd(prv, nxt, flow, g, "tf", d)
    close(gp, PIN[f"{name}/g_prv"], 1e-6)
    close(gn, PIN[f"{name}/g_nxt"], 1e-6)
    as_accurate(gf, PIN[f"{name}/g_flow"], PIN[f"{name}/g_flow_exact"])
    a64 = [t.astype(np.float64) for t in (prv, nxt, flow)]
    close(oracle.warp_cost_volume(*a64, "tf", d), PIN[f"{name}/out_exact"], 1e-13)
    gp, gn, gf = oracle.warp_cost_volume_bwd(*a64, g.astype(np.float64), "tf", d)
    close(gp, PIN[f"{name}/g_prv_exact"], 1e-12)
    close(gn, PIN[f"{name}/g_nxt_exact"], 1e-12)
    close(gf, PIN[f"{name}/g_flow_exact"], 1e-12)


@pytest.mark.parametrize("name", names("occ_"))
def test_occlusion_map_reproduces_reference(name):
    np.testing.assert_array_equal(oracle.occlusion_map(PIN[f"{name}/flow"]), PIN[f"{name}/map"])


@pytest.mark.parametrize("name", names("kat_"))
def test_known_answer_cases_through_reference_code(name):
    img, flow = PIN[f"{name}/img"], PIN[f"{name}/flow"]
    np.testing.assert_array_equal(oracle.warp(img, flow, "tf"), PIN[f"{name}/tf"])
    np.testing.assert_array_equal(oracle.warp(img, flow, "tfa"), PIN[f"{name}/tfa"])


def test_known_answers_say_what_the_survey_says():
    """zero flow => tf_warp zeroes the last row/column, WarpV2 is the identity; the hot
    pixel of the 3x3 case moves from column 1 to column 0 (app/optical_flow/test_warp.py:25-33)."""
    z = PIN["kat_zero/tf"]
    img = PIN["kat_zero/img"]
    assert np.all(z[:, -1] == 0) and np.all(z[:, :, -1] == 0)
    np.testing.assert_array_equal(z[:, :-1, :-1], img[:, :-1, :-1])
    # tfa: identity, except that the clamped last row/column is `1*(b-a)+a` -- b up to one rounding
    np.testing.assert_array_equal(PIN["kat_zero/tfa"][:, :-1, :-1], img[:, :-1, :-1])
    np.testing.assert_allclose(PIN["kat_zero/tfa"], img, rtol=0, atol=1.2e-7)
    assert np.all(PIN["kat_far/tf"] == 0) and np.all(PIN["kat_neg_far/tf"] == 0)
    np.testing.assert_allclose(PIN["kat_far/tfa"], np.broadcast_to(img[:, -1:, -1:], img.shape), rtol=0, atol=2.4e-7)
    np.testing.assert_array_equal(PIN["kat_neg_far/tfa"], np.broadcast_to(img[:, :1, :1], img.shape))
    for mode in ("tf", "tfa"):
        hot = PIN[f"kat_onehot/{mode}"][0, :, :, 0]
        assert hot[1, 0] == 1.0 and hot.sum() == 1.0


def test_config1_reference_shapes():
    """test/test_cost_volume.py:20-21 and test/test_warp.py:24-25 shapes, seeded."""
    r1 = np.random.default_rng(int(CFG1["seed"]))
    f32 = lambda a: np.asarray(a, dtype=np.float32)  # noqa: E731
    prv, nxt = f32(r1.standard_normal((4, 32, 64, 3))), f32(r1.standard_normal((4, 32, 64, 3)))
    img, flo = f32(r1.random((4, 32, 64, 3))), f32(r1.standard_normal((4, 32, 64, 2)))
    g_cv, g_w = f32(r1.standard_normal((4, 32, 64, 81))), f32(r1.standard_normal((4, 32, 64, 3)))
    np.testing.assert_array_equal(prv[0, 0, :4], CFG1["prv/head"])
    samp = (slice(None), slice(None, None, 5), slice(None, None, 7))
    cv = oracle.cost_volume(prv, nxt, 4)
    close(cv[samp], CFG1["cv/sample"], 1e-6)
    assert abs(cv.sum(dtype=np.float64) - float(CFG1["cv/sum"])) <= 1e-6 * np.abs(cv).sum(dtype=np.float64)
    gp, gn = oracle.cost_volume_bwd(prv, nxt, cv, g_cv, 4)
    close(gp[samp], CFG1["cv/g_prv"], 1e-6)
    close(gn[samp], CFG1["cv/g_nxt"], 1e-6)
    for mode in ("tf", "tfa"):
        w = oracle.warp(img, flo, mode)
        np.testing.assert_array_equal(w[samp], CFG1[f"warp/{mode}/sample"])
        assert w.sum(dtype=np.float64) == float(CFG1[f"warp/{mode}/sum"])
        gi, gf = oracle.warp_bwd(img, flo, g_w, mode)
        close(gi[samp], CFG1[f"warp/{mode}/g_img"], 1e-6)
        close(gf[samp], CFG1[f"warp/{mode}/g_flow"], 2e-6)
    r2 = np.random.default_rng(2)
    prv, nxt = f32(r2.standard_normal((1, 128, 256, 3))), f32(r2.standard_normal((1, 128, 256, 3)))
    cv = oracle.cost_volume(prv, nxt, 4)                       # app/test/test_cvol_equal.py:10
    close(cv[:, ::9, ::11], CFG1["cvol_equal/sample"], 1e-6)


@pytest.mark.skipif(not os.path.isdir("/root/reference/qpwcnet"), reason="reference checkout absent (GPU box)")
def test_fixtures_are_what_the_reference_computes_today():
    """Re-executes the reference under the shim and compares with the committed fixtures."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, "-B", os.path.join(root, "oracle", "pin_to_reference.py"), "--check"],
                       capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
