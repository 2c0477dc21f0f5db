"""qpwcnet_b200 -- B200-native (sm_100a) cost-volume / warp hot path of yycho0108/qpwcnet.

Layout: ``csrc/`` CUDA kernels + C ABI (include/qpwc.h) -> ``lib/libqpwc.so``; ``_cabi`` ctypes +
DLPack bridge; ``ops`` autograd functions; ``core`` the reference's layer/functor call surface.
"""
from .backend import image_data_format, set_image_data_format  # noqa: F401

__all__ = ["image_data_format", "set_image_data_format"]
__version__ = "0.1.0"
