import torch
import torch.nn.functional as F


class FusedLayerNormAffineFunction:
    """Call-compatible with apex's autograd Function: ``.apply(x, weight, bias, shape, eps)``."""

    @staticmethod
    def apply(x, weight, bias, normalized_shape, eps):
        return F.layer_norm(x, tuple(normalized_shape), weight, bias, eps)


class FusedLayerNorm(torch.nn.LayerNorm):
    pass
