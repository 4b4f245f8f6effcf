from . import fused_layer_norm  # noqa: F401
from .fused_layer_norm import FusedLayerNorm  # noqa: F401
