"""Plain-callable twins of the layer classes -- the ones the reference's models actually use
(qpwcnet/core/non_layers.py:51-158, imported by qpwcnet/core/pwcnet.py:7-17)."""
from __future__ import annotations

from . import _impl
from ..backend import get_axis

_get_axis = get_axis


class _Functor:
    def __init__(self, *args, data_format=None, **kwargs):
        self.data_format = _impl.resolve_format(data_format)
        self.axis = get_axis(self.data_format)


class CostVolume(_Functor):
    """qpwcnet/core/non_layers.py:51-104."""

    def __init__(self, search_range=4, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.search_range = int(search_range)

    def __call__(self, inputs):
        return _impl.cost_volume(inputs, self.search_range, self.data_format)


class CostVolumeV2(CostVolume):
    """qpwcnet/core/non_layers.py:107-123."""


class Warp(_Functor):
    """qpwcnet/core/non_layers.py:126-134."""

    mode = "tf"

    def __call__(self, inputs):
        return _impl.warp(inputs, self.mode, self.data_format)


class WarpV2(Warp):
    """qpwcnet/core/non_layers.py:137-158."""

    mode = "tfa"


class WarpCostVolume(_Functor):
    """Fused UpFlow pair (non_layers.py:377-380); call with ``(prv, nxt, flo)``."""

    def __init__(self, search_range=4, warp_mode="tfa", *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.search_range = int(search_range)
        self.warp_mode = warp_mode

    def __call__(self, inputs):
        return _impl.warp_cost_volume(inputs, self.warp_mode, self.search_range, self.data_format)


class HalfFlowWarps(_Functor):
    """The two half-flow warps at the head of ``FrameInterpolate.__call__``
    (qpwcnet/core/non_layers.py:303-311): call with ``(prv, nxt, flo_01, flo_10)``; returns
    ``concat([warp((prv, 0.5*flo_10)), warp((nxt, 0.5*flo_01))], channel axis)`` -- the reference's
    ``[prv_w, nxt_w]`` -- from one kernel launch, with the 0.5 scaling fused."""

    def __init__(self, warp_mode="tfa", flow_scale=0.5, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.warp_mode = warp_mode
        self.flow_scale = float(flow_scale)

    def __call__(self, inputs):
        return _impl.half_flow_warps(inputs, self.warp_mode, self.data_format, self.flow_scale)


class Upsample(_Functor):
    """qpwcnet/core/non_layers.py:183-193: ``scale * UpSampling2D(interpolation='bilinear')(x)``."""

    def __init__(self, scale: float = 1.0, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.scale = float(scale)

    def __call__(self, x):
        return _impl.upsample(x, self.scale, self.data_format)
