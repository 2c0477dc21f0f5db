"""``qpwcnet.core.occlusion`` call surface (qpwcnet/core/occlusion.py:27-118), backed by libqpwc."""
from __future__ import annotations

from .. import ops
from ._impl import resolve_format


def estimate_occlusion_map(flow, data_format: str = None):
    """Occlusion map of an optical flow: which pixels of the `next` frame cannot be determined from
    the previous frame (1) -- same arguments and result as the reference function.  ``flow`` follows
    ``prv[i,j] = nxt[i+f[i,j,1], j+f[i,j,0]]``; NHWC or NCHW per ``data_format`` (default: the
    backend's image_data_format(), like the reference)."""
    return ops.occlusion_map(flow, resolve_format(data_format))
