"""`networks` surface of the reference (mixing_style_transfer/networks/__init__.py star-imports both modules)."""
from .architectures import *  # noqa: F401,F403
from .architectures import FXencoder, TCNBlock, TCNModel  # noqa: F401
from .network_utils import *  # noqa: F401,F403
from .network_utils import Conv1d_layer, FiLM, Res_ConvBlock  # noqa: F401
