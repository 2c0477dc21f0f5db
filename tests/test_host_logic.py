"""CPU-side checks that need no GPU: the C-ABI library loads and exports every symbol declared in
include/qpwc.h, the DLPack bridge reads tensors correctly, the drop-in layer surface mirrors the
reference's constructors / config / error behaviour, and argument validation rejects bad calls
before any device work."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "qpwc.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(qpwc_[a-z_0-9]+)\s*\(", txt)))


def test_library_builds_loads_and_exports_header_symbols():
    from qpwcnet_b200 import _cabi, build
    so = build.build()
    assert os.path.exists(so)
    L = ctypes.CDLL(so)
    syms = _header_symbols()
    assert len(syms) >= 12
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/qpwc.h but not exported"
    assert sorted(_cabi.EXPORTED_SYMBOLS) == syms
    assert _cabi.lib().qpwc_version() >= 100
    assert _cabi.lib().qpwc_last_error() == b""


def test_argument_validation_needs_no_device():
    from qpwcnet_b200 import _cabi
    L = _cabi.lib()
    # null pointers / bad ranges are rejected before any CUDA call
    assert L.qpwc_corr_fwd(None, None, None, 1, 4, 4, 3, 4, 0.1, 81, None) == _cabi.QPWC_ERR_INVALID
    assert b"prv is NULL" in L.qpwc_last_error()
    assert L.qpwc_corr_fwd(None, None, None, 1, 4, 4, 3, 0, 0.1, 81, None) == _cabi.QPWC_ERR_INVALID
    assert b"search_range" in L.qpwc_last_error()
    assert L.qpwc_corr_fwd(None, None, None, 1, 4, 4, 3, 4, 0.1, 80, None) == _cabi.QPWC_ERR_INVALID
    assert L.qpwc_warp_fwd(None, None, None, 1, 1, 4, 3, 1, None) == _cabi.QPWC_ERR_INVALID
    assert b"2x2" in L.qpwc_last_error()
    assert L.qpwc_warp_fwd(None, None, None, 1, 4, 4, 3, 7, None) == _cabi.QPWC_ERR_INVALID
    assert L.qpwc_corr_fwd(None, None, None, 0, 4, 4, 3, 4, 0.1, 81, None) == _cabi.QPWC_OK   # empty batch
    assert L.qpwc_warp_corr_bwd_workspace(2, 3, 4, 5) == 0      # since 0.2: no caller workspace
    assert L.qpwc_set_option(0, 7) == _cabi.QPWC_ERR_INVALID and L.qpwc_set_option(0, 0) == _cabi.QPWC_OK
    with pytest.raises(_cabi.QpwcError, match="NULL"):
        _cabi.check(L.qpwc_warp_bwd(None, None, None, None, None, 1, 4, 4, 3, 0, None))


def test_dlpack_view_reads_pointer_shape_device():
    from qpwcnet_b200._cabi import dlview, kDLCPU
    t = torch.arange(24, dtype=torch.float32).reshape(1, 2, 3, 4)
    v = dlview(t)
    assert v.ptr == t.data_ptr() and v.shape == (1, 2, 3, 4) and v.device_type == kDLCPU
    assert v.is_contiguous() and v.on_host and not v.on_cuda
    sub = t[:, :, 1:, :]
    vs = dlview(sub)
    assert vs.ptr == sub.data_ptr() and not vs.is_contiguous()
    with pytest.raises(TypeError):
        dlview(t.double())
    vg = dlview(t.clone().requires_grad_())        # detached export of a leaf that requires grad
    assert vg.shape == (1, 2, 3, 4)
    a = np.zeros((2, 2), np.float32)
    assert dlview(a).ptr == a.ctypes.data          # any __dlpack__ producer works


def test_layer_surface_mirrors_reference():
    import qpwcnet_b200
    from qpwcnet.core import layers, non_layers
    from qpwcnet.core.warp import tf_warp  # noqa: F401
    qpwcnet_b200.set_image_data_format("channels_last")
    for mod in (layers, non_layers):
        for name in ("CostVolume", "CostVolumeV2", "Warp", "WarpV2"):
            assert hasattr(mod, name)
    cv = layers.CostVolume(search_range=3, name="cv")
    assert cv.search_range == 3 and cv.data_format == "channels_last" and cv.axis == 3
    cfg = cv.get_config()
    assert cfg["search_range"] == 3 and cfg["name"] == "cv"
    assert layers.CostVolume.from_config(cfg).search_range == 3
    assert layers.CostVolumeV2().search_range == 4
    qpwcnet_b200.set_image_data_format("channels_first")
    try:
        assert layers.Warp().data_format == "channels_first" and layers.Warp().axis == 1
        assert non_layers.WarpV2().data_format == "channels_first"
        assert layers.WarpV2(data_format="channels_last").data_format == "channels_last"
    finally:
        qpwcnet_b200.set_image_data_format("channels_last")
    with pytest.raises(ValueError, match="Unsupported data format"):
        layers.CostVolume(data_format="NCHW")
    with pytest.raises(ValueError, match="Unsupported data format"):
        qpwcnet_b200.set_image_data_format("bogus")


def test_ops_validate_before_touching_the_device():
    from qpwcnet_b200 import ops
    x = torch.zeros((1, 4, 5, 3))
    with pytest.raises(ValueError):
        ops.cost_volume(x, torch.zeros((1, 4, 6, 3)), 4)
    with pytest.raises(TypeError):
        ops.cost_volume(x.double(), x.double(), 4)
    with pytest.raises(ValueError):
        ops.warp(x, torch.zeros((1, 4, 5, 3)), "tf")
    with pytest.raises(ValueError, match="mode"):
        ops.warp(x, torch.zeros((1, 4, 5, 2)), "bilinear")
    with pytest.raises(RuntimeError, match="inference-only"):
        ops.cost_volume(x.clone().requires_grad_(), x, 4)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "qpwcnet_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "libqpwc_emu" not in src and "libqpwc_oracle" not in src, f


def test_bind_host_thread_near_is_best_effort():
    """No GPU / no NVML: the affinity helper must return None and leave the affinity untouched."""
    import os

    from qpwcnet_b200 import ops
    before = os.sched_getaffinity(0)
    old = ops.bind_host_thread_near(0)
    assert old is None or isinstance(old, set)
    if old is not None:
        os.sched_setaffinity(0, old)
    assert os.sched_getaffinity(0) == before
