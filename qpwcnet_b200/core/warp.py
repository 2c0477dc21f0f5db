"""``tf_warp(img, flow, data_format=None)`` -- qpwcnet/core/warp.py:63-153 -- and the tfa-style
``dense_image_warp`` of warp.py:156-211, both over libqpwc."""
from __future__ import annotations

from . import _impl


def tf_warp(img, flow, data_format=None):
    """Bilinear backward warp with the reference's truncate/clip border rule."""
    return _impl.warp((img, flow), "tf", _impl.resolve_format(data_format))


def dense_image_warp(image, flow, name=None):
    """The reference's local variant of tfa's routine: NHWC only, flow is (dy, dx) and the query is
    ``grid + flow`` (qpwcnet/core/warp.py:201), tfa border rule."""
    return _impl.warp((image, flow.flip(-1)), "tfa", "channels_last")
