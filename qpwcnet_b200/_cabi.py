"""ctypes binding of libqpwc.so (include/qpwc.h) and the DLPack -> raw pointer bridge.

Tensors cross the Python/C boundary as DLPack capsules: the producer (PyTorch here; anything that
implements ``__dlpack__``, e.g. ``tf.experimental.dlpack.to_dlpack`` on the reference side) exports
a ``DLManagedTensor``; this module reads pointer / shape / strides / device straight out of that
struct with ctypes and hands plain pointers and sizes to the C ABI.  There is no CPU fallback and no
alternative backend: if the CUDA library is missing the import of any op fails loudly.
"""
from __future__ import annotations

import ctypes
import os
import struct
from ctypes import (POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_int64, c_longlong,
                    c_size_t, c_uint8, c_uint16, c_uint32, c_uint64, c_void_p)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libqpwc.so")

QPWC_OK, QPWC_ERR_INVALID, QPWC_ERR_UNSUPPORTED, QPWC_ERR_CUDA = 0, 1, 2, 3
WARP_MODES = {"tf": 0, "tfa": 1}

# DLPack device types
kDLCPU, kDLCUDA, kDLCUDAHost, kDLCUDAManaged = 1, 2, 3, 13


class QpwcError(RuntimeError):
    """Non-zero status from libqpwc (message from qpwc_last_error())."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"libqpwc error {code}: {msg}")
        self.code = code


_lib = None


def lib() -> ctypes.CDLL:
    """The loaded C-ABI library.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m qpwcnet_b200.build` "
            "(qpwcnet_b200 has no CPU or non-CUDA path)")
    L = ctypes.CDLL(LIB_PATH)
    fp, vp, i, f, ll = c_void_p, c_void_p, c_int, c_float, c_longlong
    sigs = {
        "qpwc_version": ([], c_int),
        "qpwc_last_error": ([], c_char_p),
        "qpwc_corr_fwd": ([fp, fp, fp, i, i, i, i, i, f, ll, vp], c_int),
        "qpwc_corr_fwd_nchw": ([fp, fp, fp, i, i, i, i, i, f, vp], c_int),
        "qpwc_corr_bwd": ([fp, fp, fp, fp, fp, fp, i, i, i, i, i, f, ll, vp], c_int),
        "qpwc_warp_fwd": ([fp, fp, fp, i, i, i, i, i, vp], c_int),
        "qpwc_warp_fwd_nchw": ([fp, fp, fp, i, i, i, i, i, f, vp], c_int),
        "qpwc_warp_bwd": ([fp, fp, fp, fp, fp, i, i, i, i, i, vp], c_int),
        "qpwc_warp_fwd_rows": ([fp, fp, fp, i, i, i, i, i, i, i, vp], c_int),
        "qpwc_warp_fwd_ex": ([fp, fp, fp, i, i, i, i, i, f, ll, vp], c_int),
        "qpwc_warp_pair_fwd": ([fp, fp, fp, fp, fp, i, i, i, i, i, f, ll, vp], c_int),
        "qpwc_warp_bwd_ex": ([fp, fp, fp, fp, fp, i, i, i, i, i, f, ll, vp], c_int),
        "qpwc_upsample2x_fwd": ([fp, fp, i, i, i, i, f, vp], c_int),
        "qpwc_upsample2x_bwd": ([fp, fp, i, i, i, i, f, vp], c_int),
        "qpwc_corr_bwd_nchw": ([fp, fp, fp, fp, fp, fp, i, i, i, i, i, f, vp], c_int),
        "qpwc_warp_bwd_nchw": ([fp, fp, fp, fp, fp, i, i, i, i, i, f, vp], c_int),
        "qpwc_occlusion_map": ([fp, fp, i, i, i, i, vp], c_int),
        "qpwc_warp_fwd_up": ([fp, fp, fp, i, i, i, i, i, f, vp], c_int),
        "qpwc_warp_corr_fwd_up": ([fp, fp, fp, fp, i, i, i, i, i, f, i, ll, f, vp], c_int),
        "qpwc_warp_corr_fwd": ([fp, fp, fp, fp, i, i, i, i, i, f, i, ll, vp], c_int),
        "qpwc_warp_corr_bwd_workspace": ([i, i, i, i], c_size_t),
        "qpwc_warp_corr_bwd": ([fp, fp, fp, fp, fp, fp, fp, fp, vp, c_size_t, i, i, i, i, i, f, i, ll, vp], c_int),
        "qpwc_corr_fwd_host": ([fp, fp, fp, i, i, i, i, i, f, i], c_int),
        "qpwc_warp_fwd_host": ([fp, fp, fp, i, i, i, i, i, i], c_int),
        "qpwc_warp_corr_fwd_host": ([fp, fp, fp, fp, i, i, i, i, i, f, i, i], c_int),
        "qpwc_host_set_deferred": ([i], c_int),
        "qpwc_host_sync": ([i], c_int),
        "qpwc_set_option": ([i, i], c_int),
        "qpwc_get_option": ([i], c_int),
    }
    for name, (argtypes, restype) in sigs.items():
        fn = getattr(L, name)          # AttributeError here = header/library mismatch: be loud
        fn.argtypes = argtypes
        fn.restype = restype
    _lib = L
    return L


EXPORTED_SYMBOLS = (
    "qpwc_version", "qpwc_last_error", "qpwc_corr_fwd", "qpwc_corr_fwd_nchw", "qpwc_corr_bwd", "qpwc_warp_fwd",
    "qpwc_warp_fwd_nchw", "qpwc_warp_bwd", "qpwc_warp_fwd_rows", "qpwc_warp_fwd_ex", "qpwc_warp_pair_fwd", "qpwc_warp_bwd_ex", "qpwc_upsample2x_fwd", "qpwc_upsample2x_bwd", "qpwc_occlusion_map", "qpwc_warp_bwd_nchw", "qpwc_corr_bwd_nchw",
    "qpwc_warp_fwd_up", "qpwc_warp_corr_fwd_up", "qpwc_warp_corr_fwd", "qpwc_warp_corr_bwd_workspace", "qpwc_warp_corr_bwd",
    "qpwc_corr_fwd_host", "qpwc_warp_fwd_host", "qpwc_warp_corr_fwd_host",
    "qpwc_host_set_deferred", "qpwc_host_sync", "qpwc_set_option", "qpwc_get_option",
)


def check(rc: int) -> None:
    if rc != QPWC_OK:
        msg = lib().qpwc_last_error()
        raise QpwcError(rc, msg.decode() if msg else "")


# --------------------------------------------------------------------------------------- DLPack
class _DLDevice(Structure):
    _fields_ = [("device_type", c_int32), ("device_id", c_int32)]


class _DLDataType(Structure):
    _fields_ = [("code", c_uint8), ("bits", c_uint8), ("lanes", c_uint16)]


class _DLTensor(Structure):
    _fields_ = [("data", c_void_p), ("device", _DLDevice), ("ndim", c_int32),
                ("dtype", _DLDataType), ("shape", POINTER(c_int64)),
                ("strides", POINTER(c_int64)), ("byte_offset", c_uint64)]


class _DLManagedTensor(Structure):
    _fields_ = [("dl_tensor", _DLTensor), ("manager_ctx", c_void_p), ("deleter", c_void_p)]


class _DLPackVersion(Structure):
    _fields_ = [("major", c_uint32), ("minor", c_uint32)]


class _DLManagedTensorVersioned(Structure):
    _fields_ = [("version", _DLPackVersion), ("manager_ctx", c_void_p), ("deleter", c_void_p),
                ("flags", c_uint64), ("dl_tensor", _DLTensor)]


_DLT = struct.Struct("<QiiiBBHQQQ")
_I64 = [struct.Struct("<%dq" % n) for n in range(9)]
_string_at = ctypes.string_at
_capi = ctypes.pythonapi
_PyCapsule_GetName = _capi.PyCapsule_GetName
_PyCapsule_GetName.restype = c_char_p
_PyCapsule_GetName.argtypes = [ctypes.py_object]
_PyCapsule_GetPointer = _capi.PyCapsule_GetPointer
_PyCapsule_GetPointer.restype = c_void_p
_PyCapsule_GetPointer.argtypes = [ctypes.py_object, c_char_p]


class DLView:
    """Pointer/shape/device of a tensor read out of its DLPack capsule.  Holds the capsule (and so
    the producer's memory) alive for as long as the view lives.  Pointer, device and dtype are read
    eagerly (one 48-byte unpack -- this sits on the launch path); shape/strides on demand."""

    __slots__ = ("ptr", "device_type", "device_id", "ndim", "_shape_p", "_strides_p", "_capsule")

    def __init__(self, capsule, name=b"dltensor"):
        raw = _PyCapsule_GetPointer(capsule, name)
        # DLTensor = {void* data; int32 device_type, device_id; int32 ndim; uint8 code, bits;
        # uint16 lanes; int64* shape; int64* strides; uint64 byte_offset}: 48 bytes, first member
        # of DLManagedTensor, at offset 32 of DLManagedTensorVersioned.
        (data, self.device_type, self.device_id, self.ndim, code, bits, lanes, self._shape_p,
         self._strides_p, byte_offset) = _DLT.unpack(_string_at(raw if name == b"dltensor" else raw + 32, 48))
        if code != 2 or bits != 32 or lanes != 1:
            raise TypeError(f"libqpwc takes float32 tensors (DLPack dtype code={code} bits={bits})")
        self.ptr = data + byte_offset
        self._capsule = capsule

    @property
    def shape(self):
        nd = self.ndim
        return _I64[nd].unpack(_string_at(self._shape_p, 8 * nd)) if nd else ()

    @property
    def strides(self):
        nd = self.ndim
        return _I64[nd].unpack(_string_at(self._strides_p, 8 * nd)) if (self._strides_p and nd) else None

    def is_contiguous(self) -> bool:
        strides = self.strides
        if strides is None:
            return True
        expect = 1
        for n, s in zip(reversed(self.shape), reversed(strides)):
            if n != 1 and s != expect:
                return False
            expect *= n
        return True

    @property
    def on_cuda(self) -> bool:
        return self.device_type in (kDLCUDA, kDLCUDAManaged)

    @property
    def on_host(self) -> bool:
        return self.device_type in (kDLCPU, kDLCUDAHost)


try:  # torch is the usual producer, but any __dlpack__ object works
    import torch as _torch
    _to_dlpack = _torch.utils.dlpack.to_dlpack
except ImportError:  # pragma: no cover
    _torch = None


def dlptr(x):
    """Launch-path shortcut of dlview() for torch tensors the caller keeps alive: returns
    ``(device pointer, DLPack device_type)`` read from the tensor's DLPack export."""
    cap = _to_dlpack(x.detach() if x.requires_grad else x)
    (data, devt, _id, _nd, code, bits, lanes, _s, _st, off) = _DLT.unpack(
        _string_at(_PyCapsule_GetPointer(cap, b"dltensor"), 48))
    if code != 2 or bits != 32 or lanes != 1:
        raise TypeError(f"libqpwc takes float32 tensors (DLPack dtype code={code} bits={bits})")
    return data + off, devt


def dlview(x) -> DLView:
    """DLPack view of a tensor-like (torch.Tensor, or any object with ``__dlpack__``)."""
    if _torch is not None and isinstance(x, _torch.Tensor):
        # torch exports the legacy (unversioned) capsule; detach(): exporting a tensor that requires
        # grad is refused by torch, the autograd Functions own the graph anyway
        return DLView(_to_dlpack(x.detach() if x.requires_grad else x))
    if hasattr(x, "__dlpack__"):
        cap = x.__dlpack__()
        return DLView(cap, _PyCapsule_GetName(cap))
    raise TypeError(f"cannot export {type(x).__name__} through DLPack")
