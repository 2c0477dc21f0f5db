"""`tensorflow_addons` stand-in  --  TEST INFRASTRUCTURE ONLY.

tensorflow_addons is third-party, un-vendored and unpinned by the reference (setup.py:19), so its
arithmetic CANNOT be pinned by executing anything under /root/reference.  What is here:

* `image.dense_image_warp` / `image.interpolate_bilinear`: restated from tfa's published algorithm
  (floor clamped to [0, size-2], alpha clamped to [0, 1], nested lerps) so that the reference's
  *glue* around it (`WarpV2.call`: sign flip, channel reversal, NCHW transposes, layers.py:177-186)
  can be executed.  The glue is pinned that way; the tfa arithmetic itself stays PARITY UNPINNED.
* `layers.optical_flow.CorrelationCost`: deliberately absent (raises).  The reference's own tests pin
  it to the in-repo `CostVolume` ("0.0", app/test/test_cvol_equal.py:25), which IS executed.
"""
import torch as _t

import tensorflow as _tf


class _Image:
    @staticmethod
    def interpolate_bilinear(grid, query_points, indexing="ij"):
        if indexing not in ("ij", "xy"):
            raise ValueError("Indexing mode must be 'ij' or 'xy'")
        B, H, W, C = grid.shape
        if H < 2 or W < 2:
            raise ValueError("Grid must be at least 2x2.")
        q = _t.unbind(query_points, dim=2)
        if indexing == "xy":
            q = q[::-1]
        alphas, floors, ceils = [], [], []
        for dim, size in ((0, H), (1, W)):
            qd = q[dim]
            max_floor = _t.tensor(size - 2, dtype=qd.dtype)
            min_floor = _t.tensor(0.0, dtype=qd.dtype)
            floor = _t.minimum(_t.maximum(min_floor, _t.floor(qd)), max_floor)
            int_floor = floor.to(_t.int64)
            floors.append(int_floor)
            ceils.append(int_floor + 1)
            alpha = qd - floor
            alpha = _t.minimum(_t.maximum(_t.tensor(0.0, dtype=qd.dtype), alpha), _t.tensor(1.0, dtype=qd.dtype))
            alphas.append(alpha.unsqueeze(2))
        flat = grid.reshape(B * H * W, C)
        boff = (_t.arange(B) * H * W).reshape(B, 1)

        def gather(y, x):
            return flat[(boff + y * W + x)]

        tl, tr = gather(floors[0], floors[1]), gather(floors[0], ceils[1])
        bl, br = gather(ceils[0], floors[1]), gather(ceils[0], ceils[1])
        top = alphas[1] * (tr - tl) + tl
        bot = alphas[1] * (br - bl) + bl
        return _tf._wrap(alphas[0] * (bot - top) + top)

    @staticmethod
    def dense_image_warp(image, flow, name=None):
        B, H, W, C = image.shape
        gx, gy = _t.meshgrid(_t.arange(W), _t.arange(H), indexing="xy")
        grid = _t.stack([gy, gx], dim=2).to(flow.dtype).unsqueeze(0)
        query = (grid - flow).reshape(B, H * W, 2)
        out = _Image.interpolate_bilinear(image, query)
        return _tf._wrap(out.reshape(B, H, W, C))


class _CorrelationCost:
    def __init__(self, *args, **kwargs):
        self.args = args

    def __call__(self, inputs):
        raise NotImplementedError(
            "tfa CorrelationCost is un-vendored third-party code; the reference pins it to "
            "CostVolume (app/test/test_cvol_equal.py:25), which is what the pin script executes")


class _OpticalFlow:
    CorrelationCost = _CorrelationCost


class _Layers:
    optical_flow = _OpticalFlow()


image = _Image()
layers = _Layers()
