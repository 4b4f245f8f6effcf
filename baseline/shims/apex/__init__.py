"""torch-native stand-in for NVIDIA apex (reference arm only; see ../README.md)."""
from . import normalization, optimizers, multi_tensor_apply, parallel, amp  # noqa: F401
