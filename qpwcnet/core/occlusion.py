from qpwcnet_b200.core.occlusion import *  # noqa: F401,F403
from qpwcnet_b200.core import occlusion as _m

globals().update({k: v for k, v in vars(_m).items() if not k.startswith("__")})
