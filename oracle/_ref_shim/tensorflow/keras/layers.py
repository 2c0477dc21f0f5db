import torch as _t

from . import backend as _backend

# keras.layers.Layer.__init__ validates its kwargs against this allow-list and raises TypeError for
# anything else -- which is why the reference's test scripts (`CostVolume(..., data_format=...)`,
# test/test_cost_volume.py:10) are stale against layers.py as shipped.
_ALLOWED = {"input_dim", "input_shape", "batch_input_shape", "batch_size", "weights",
            "activity_regularizer", "autocast", "implementation"}


class Layer:
    def __init__(self, trainable=True, name=None, dtype=None, dynamic=False, **kwargs):
        for k in kwargs:
            if k not in _ALLOWED:
                raise TypeError("Keyword argument not understood:", k)
        self.trainable, self.name, self.built = trainable, name or type(self).__name__.lower(), False
        self._dtype = dtype or "float32"

    def build(self, input_shape):
        self.built = True

    def call(self, inputs):
        return inputs

    def __call__(self, inputs, *args, **kwargs):
        if not self.built:
            shapes = (tuple(tuple(x.shape) for x in inputs) if isinstance(inputs, (tuple, list))
                      else tuple(inputs.shape))
            self.build(shapes)
            self.built = True
        return self.call(inputs, *args, **kwargs)

    def get_config(self):
        return {"name": self.name, "trainable": self.trainable, "dtype": self._dtype}


class Activation(Layer):
    def __init__(self, activation, **kwargs):
        super().__init__(**kwargs)
        self.activation = activation

    def call(self, inputs):
        return self.activation(inputs)


class Lambda(Layer):
    def __init__(self, function, **kwargs):
        super().__init__(**kwargs)
        self.function = function

    def call(self, inputs):
        return self.function(inputs)


class ZeroPadding2D(Layer):
    """Zero rows/columns around the two spatial axes; int padding = symmetric on both axes."""

    def __init__(self, padding=(1, 1), data_format=None, **kwargs):
        super().__init__(**kwargs)
        if isinstance(padding, int):
            padding = ((padding, padding), (padding, padding))
        elif isinstance(padding[0], int):
            padding = ((padding[0], padding[0]), (padding[1], padding[1]))
        self.padding = padding
        self.data_format = data_format or _backend.image_data_format()

    def call(self, x):
        (t, b), (l, r) = self.padding
        if self.data_format == "channels_first":       # (N, C, H, W)
            return _t.nn.functional.pad(x, (l, r, t, b))
        return _t.nn.functional.pad(x, (0, 0, l, r, t, b))   # (N, H, W, C)
