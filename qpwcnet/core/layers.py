from qpwcnet_b200.core.layers import *  # noqa: F401,F403
from qpwcnet_b200.core import layers as _m

globals().update({k: v for k, v in vars(_m).items() if not k.startswith("__")})
