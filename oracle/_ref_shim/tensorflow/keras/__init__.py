"""tf.keras stand-in (see ../__init__.py): only the backend global, the Layer call protocol and the
one layer (ZeroPadding2D) the hot-path source uses.  Everything else the reference's *other* classes
name in their constructors (Conv2D, BatchNormalization, ...) is absent on purpose: those classes are
out of scope and are never instantiated by the pinning script."""
from . import backend, layers, regularizers, utils  # noqa: F401
