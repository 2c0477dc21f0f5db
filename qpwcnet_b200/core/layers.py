"""Layer classes with the reference's names, constructor arguments and tuple-call convention
(qpwcnet/core/layers.py:32-186), as ``torch.nn.Module``s over the sm_100a kernels.

    CostVolume(search_range=4)((prv, nxt))      CostVolumeV2(search_range=4)((prv, nxt))
    Warp()((img, flo))                          WarpV2()((img, flo))

Layout follows the process-global ``image_data_format()`` read at construction, exactly like the
reference (layers.py:41,119,146,173); the optional ``data_format=`` keyword that the reference's own
tests pass (test/test_cost_volume.py:10-11, test/test_warp.py:14-15) overrides it.  ``name=`` and
other Keras ``Layer`` kwargs are accepted and kept in ``get_config()``.  ``WarpCostVolume`` is the
fused UpFlow pair (non_layers.py:377-380) -- an addition, not a reference class.
"""
from __future__ import annotations

import torch

from . import _impl
from ..backend import get_axis

_get_axis = get_axis


def lrelu(x):
    return torch.nn.functional.leaky_relu(x, 0.1)


class _Layer(torch.nn.Module):
    def __init__(self, *args, data_format=None, name=None, **kwargs):
        super().__init__()
        self.data_format = _impl.resolve_format(data_format)
        self.axis = get_axis(self.data_format)
        self.name = name
        self._keras_kwargs = dict(kwargs)

    def get_config(self):
        cfg = {"name": self.name}
        cfg.update(self._keras_kwargs)
        cfg.update(getattr(self, "_config", {}))
        return cfg

    @classmethod
    def from_config(cls, config):
        return cls(**config)


class CostVolume(_Layer):
    """Local correlation cost volume, d = search_range, leaky_relu(0.1) included
    (qpwcnet/core/layers.py:32-109)."""

    def __init__(self, search_range=4, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self._config = {"search_range": search_range}
        self.search_range = int(search_range)

    def forward(self, inputs):
        return _impl.cost_volume(inputs, self.search_range, self.data_format)


class CostVolumeV2(CostVolume):
    """Same function as CostVolume (the reference pins tfa's CorrelationCost to it:
    test/test_cost_volume.py:22-24) -- qpwcnet/core/layers.py:112-141."""


class Warp(_Layer):
    """tf_warp semantics (truncate, clip, weights from clipped corners) --
    qpwcnet/core/layers.py:144-168."""

    mode = "tf"

    def forward(self, inputs):
        return _impl.warp(inputs, self.mode, self.data_format)


class WarpV2(Warp):
    """tfa.image.dense_image_warp(img, -flo[..., ::-1]) semantics (floor, edge clamp) --
    qpwcnet/core/layers.py:171-186."""

    mode = "tfa"


class WarpCostVolume(_Layer):
    """Fused ``CostVolumeV2((prv, WarpV2((nxt, flo))))`` -- the UpFlow pair,
    qpwcnet/core/non_layers.py:377-380; call with ``(prv, nxt, flo)``."""

    def __init__(self, search_range=4, warp_mode="tfa", *args, **kwargs):
        super().__init__(*args, **kwargs)
        self._config = {"search_range": search_range, "warp_mode": warp_mode}
        self.search_range = int(search_range)
        self.warp_mode = warp_mode

    def forward(self, inputs):
        return _impl.warp_cost_volume(inputs, self.warp_mode, self.search_range, self.data_format)
