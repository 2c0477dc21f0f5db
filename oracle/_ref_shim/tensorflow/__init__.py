"""Minimal `tensorflow` stand-in over torch-CPU  --  TEST INFRASTRUCTURE ONLY (oracle pinning).

TensorFlow is not installable in this image, so the reference's own hot-path source
(/root/reference/qpwcnet/core/{warp,layers,non_layers,occlusion}.py) cannot run as shipped.  This
package implements, op for op and with TF's eager semantics, exactly the ~35 `tf.*` symbols those
four files touch, so that `oracle/pin_to_reference.py` can import the UNMODIFIED reference modules
and execute them.  Tensors are torch-CPU tensors (a thin subclass that adds TF's negative-step
slicing and TF's "no implicit dtype promotion" rule), every op rounds to the tensor dtype like a TF
eager op does, and because the ops are differentiable torch ops the gradients TF autodiff derives
for the same graph (gather_nd -> scatter-add, slice -> pad, ...) come out of torch autograd.

Nothing here restates reference code: it only gives names like `tf.gather_nd` their documented
meaning.  Never imported by the product package, by bench.py or by the GPU tests.
"""
import builtins as _b

import torch as _t

from . import keras  # noqa: F401  (tf.keras.*)

__version__ = "0.0-shim"

_REAL = {"float32": _t.float32}     # `exact` mode maps tf.float32 -> torch.float64 (see set_real)


class DType:
    def __init__(self, name):
        self.name = name

    @property
    def torch(self):
        return {"float32": _REAL["float32"], "float64": _t.float64, "float16": _t.float16,
                "int32": _t.int32, "int64": _t.int64, "bool": _t.bool}[self.name]

    def __repr__(self):
        return "tf." + self.name


float32, float64, half, int32, int64 = (DType(n) for n in ("float32", "float64", "float16", "int32", "int64"))
float16 = half
bool = DType("bool")  # noqa: A001


def set_real(dtype):
    """pin_to_reference's `exact` mode: run the same reference code with tf.float32 meaning fp64, to
    measure how far the reference's own fp32 arithmetic is from exact arithmetic."""
    _REAL["float32"] = dtype


_NO_PROMOTE = {"add", "sub", "mul", "div", "true_divide", "maximum", "minimum", "__add__", "__radd__",
               "__sub__", "__rsub__", "__mul__", "__rmul__", "__truediv__", "__rtruediv__", "lt", "le",
               "gt", "ge", "__lt__", "__le__", "__gt__", "__ge__"}


class Tensor(_t.Tensor):
    """torch tensor + TF indexing (negative steps) + TF's refusal to mix dtypes in binary ops."""

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        name = getattr(func, "__name__", "")
        if name in _NO_PROMOTE:
            ts = [a for a in args if isinstance(a, _t.Tensor)]
            if len(ts) == 2 and ts[0].dtype != ts[1].dtype:
                raise TypeError(f"tf shim: {name} on {ts[0].dtype} and {ts[1].dtype} "
                                "(TensorFlow does not promote dtypes implicitly)")
        return super().__torch_function__(func, types, args, kwargs or {})

    def __getitem__(self, idx):
        items = idx if isinstance(idx, tuple) else (idx,)
        if not any(isinstance(s, _b.slice) and s.step is not None and s.step < 0 for s in items):
            return super().__getitem__(idx)
        # expand Ellipsis, then realise each negative-step slice as a positive slice of a flip
        n_real = sum(1 for s in items if s is not None and s is not Ellipsis)
        full = []
        for s in items:
            if s is Ellipsis:
                full.extend([_b.slice(None)] * (self.dim() - n_real))
            else:
                full.append(s)
        out, dim, fwd = self, 0, []
        for s in full:
            if s is None:
                fwd.append(None)
                continue
            if isinstance(s, _b.slice) and s.step is not None and s.step < 0:
                n = out.shape[dim]
                sel = list(_b.range(n))[s]
                out = _t.index_select(out, dim, _t.tensor(sel, dtype=_t.int64))
                fwd.append(_b.slice(None))
            else:
                fwd.append(s)
            if not isinstance(s, int):
                dim += 1
        return out.__getitem__(tuple(fwd))

    def numpy(self):
        return self.detach().as_subclass(_t.Tensor).numpy()


def _wrap(x):
    return x.as_subclass(Tensor) if isinstance(x, _t.Tensor) and not isinstance(x, Tensor) else x


def _dt(dtype):
    return dtype.torch if isinstance(dtype, DType) else dtype


def convert_to_tensor(x, dtype=None):
    if isinstance(x, _t.Tensor):
        return _wrap(x if dtype is None else x.to(_dt(dtype)))
    import numpy as np
    a = np.asarray(x)
    if dtype is None and a.dtype == np.float64 and not isinstance(x, np.ndarray):
        a = a.astype(np.float32)                      # TF: python floats become float32
    if dtype is None and a.dtype == np.int64 and not isinstance(x, np.ndarray):
        a = a.astype(np.int32)                        # TF: python ints become int32
    t = _t.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(_dt(dtype))
    elif t.dtype == _t.float32:
        t = t.to(_REAL["float32"])
    return _wrap(t)


def constant(value, dtype=None):
    return convert_to_tensor(value, dtype)


def _like(ref, v):
    """A python scalar / list operand takes the dtype of the tensor operand (TF's rule for
    python constants in `tf.clip_by_value(x, [0, 0], [h-1, w-1])` and friends)."""
    if isinstance(v, _t.Tensor):
        return v
    return _wrap(_t.as_tensor(v, dtype=ref.dtype))


def rank(x):
    return x.dim()


def shape(x):
    return tuple(x.shape)


def range(*args):  # noqa: A001
    return _wrap(_t.arange(*args, dtype=_t.int32))


def meshgrid(*xs, indexing="xy"):
    return [_wrap(g) for g in _t.meshgrid(*xs, indexing=indexing)]


def expand_dims(x, axis):
    return _wrap(_t.unsqueeze(convert_to_tensor(x), axis))


def cast(x, dtype):
    """tf.cast: float -> int truncates toward zero (C conversion), bool -> float gives 0/1."""
    return _wrap(convert_to_tensor(x).to(_dt(dtype)))


def concat(values, axis):
    return _wrap(_t.cat([convert_to_tensor(v) for v in values], dim=axis))


def stack(values, axis=0):
    return _wrap(_t.stack([convert_to_tensor(v) for v in values], dim=axis))


def unstack(x, axis=0):
    return [_wrap(v) for v in _t.unbind(x, dim=axis)]


def split(x, num, axis=0):
    return [_wrap(v) for v in _t.chunk(x, num, dim=axis)]


def reshape(x, shp):
    return _wrap(_t.reshape(x, tuple(int(s) for s in shp)))


def transpose(x, perm):
    return _wrap(x.permute(*perm))


def tile(x, multiples):
    return _wrap(x.repeat(*multiples))


def clip_by_value(x, lo, hi):
    return _wrap(_t.minimum(_t.maximum(x, _like(x, lo)), _like(x, hi)))


def maximum(a, b):
    return _wrap(_t.maximum(a, _like(a, b)))


def zeros_like(x, dtype=None):
    return _wrap(_t.zeros_like(x, dtype=None if dtype is None else _dt(dtype)))


def ones_like(x, dtype=None):
    return _wrap(_t.ones_like(x, dtype=None if dtype is None else _dt(dtype)))


def add_n(xs):
    """AddN: one left-to-right elementwise sum ((a+b)+c)+d, each partial rounded to the dtype."""
    out = xs[0]
    for v in xs[1:]:
        out = out + v
    return out


def slice(x, begin, size):  # noqa: A001
    idx = tuple(_b.slice(b, None if s == -1 else b + s) for b, s in zip(begin, size))
    return x[idx]


def reduce_mean(x, axis=None, keepdims=False):
    """Mean reducer: sum of the elements divided by their count."""
    return _wrap(_t.sum(x, dim=axis, keepdim=keepdims) / x.shape[axis])


def reduce_sum(x, axis=None, keepdims=False):
    return _wrap(_t.sum(x) if axis is None else _t.sum(x, dim=axis, keepdim=keepdims))


def reduce_any(x, axis=None):
    if isinstance(x, (list, tuple)):
        x = _t.stack(list(x), dim=0)
    return _wrap(_t.any(x, dim=axis))


def logical_or(a, b):
    return _wrap(_t.logical_or(a, b))


def round(x):  # noqa: A001
    return _wrap(_t.round(x))


def gather_nd(params, indices, batch_dims=0):
    """out[b..., i...] = params[b..., indices[b..., i..., 0], ..., indices[b..., i..., K-1], ...]."""
    indices = indices.to(_t.int64)
    K = indices.shape[-1]
    bshape = params.shape[:batch_dims]
    ix = []
    for k, n in enumerate(bshape):
        view = [1] * (indices.dim() - 1)
        view[k] = n
        ix.append(_t.arange(n).view(view).expand(indices.shape[:-1]))
    ix.extend(indices[..., k] for k in _b.range(K))
    return _wrap(params.as_subclass(_t.Tensor)[tuple(ix)])


def tensor_scatter_nd_min(tensor, indices, updates):
    """out = tensor; out[indices[n]] = min(out[indices[n]], updates[n]) for every n."""
    indices = indices.to(_t.int64).reshape(-1, indices.shape[-1])
    strides = _t.tensor([int(_t.tensor(tensor.shape[k + 1:]).prod()) if k + 1 < tensor.dim() else 1
                         for k in _b.range(tensor.dim())], dtype=_t.int64)
    flat = (indices * strides[: indices.shape[-1]]).sum(-1)
    out = tensor.as_subclass(_t.Tensor).reshape(-1).clone()
    out.scatter_reduce_(0, flat, updates.as_subclass(_t.Tensor).reshape(-1), reduce="amin", include_self=True)
    return _wrap(out.reshape(tensor.shape))


class _NameScope:
    def __init__(self, *_a, **_k):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


name_scope = _NameScope


class _NN:
    @staticmethod
    def leaky_relu(x, alpha=0.2):
        """LeakyRelu kernel: x > 0 ? x : alpha * x (gradient alpha at x == 0)."""
        return _wrap(_t.nn.functional.leaky_relu(x, alpha))


class _Math:
    tanh = staticmethod(lambda x: _wrap(_t.tanh(x)))
    softplus = staticmethod(lambda x: _wrap(_t.nn.functional.softplus(x)))


nn = _NN()
math = _Math()


Variable = Tensor


def is_tensor(x):
    return isinstance(x, _t.Tensor)


def executing_eagerly():
    return True


# einops picks its backend by scanning sys.modules; a module called `tensorflow` would make it try
# TF ops on these torch tensors.  Bind the shim's tensor type to einops' torch backend explicitly.
try:
    import einops._backends as _eb
    _eb._type2backend[Tensor] = _eb.TorchBackend()
except Exception:  # pragma: no cover
    pass
