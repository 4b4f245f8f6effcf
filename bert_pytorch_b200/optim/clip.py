"""Multi-tensor L2 norm / scale and the SQuAD ``GradientClipper``.

Parity: apex ``amp_C.multi_tensor_l2norm`` / ``multi_tensor_scale`` as bound in
src/optimization.py:26-33 and used by run_squad.py:703-725.  On CUDA
the kernels in ops/csrc/optim.cu run; the clip coefficient stays on the device (the
reference does a ``.item()`` host sync every micro-step).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch


def multi_tensor_l2norm(tensors: Sequence[torch.Tensor], per_tensor: bool = False
                        ) -> Tuple[torch.Tensor, torch.Tensor]:
    """(total L2 norm, per-tensor norms or empty)."""
    tensors = [t for t in tensors if t is not None]
    if not tensors:
        z = torch.zeros(())
        return z, z.new_zeros(0)
    if tensors[0].is_cuda:
        from .. import ops
        if ops.available():
            return ops.multi_tensor_l2norm(list(tensors), per_tensor)
    norms = torch.stack([t.float().norm() for t in tensors])
    return norms.norm(), (norms if per_tensor else norms.new_zeros(0))


def multi_tensor_scale(src: Sequence[torch.Tensor], dst: Sequence[torch.Tensor], scale) -> torch.Tensor:
    """dst[i] = src[i] * scale; returns an overflow flag tensor (1 if any inf/nan seen)."""
    if src and src[0].is_cuda:
        from .. import ops
        if ops.available():
            return ops.multi_tensor_scale(list(src), list(dst), scale)
    flag = torch.zeros((), dtype=torch.int32, device=src[0].device if src else "cpu")
    for s, d in zip(src, dst):
        v = s.float() * (scale if not torch.is_tensor(scale) else scale.to(s.device))
        if not bool(torch.isfinite(v).all()):
            flag.fill_(1)
        d.copy_(v.to(d.dtype))
    return flag


class GradientClipper:
    """Clip the global L2 norm of a parameter list's gradients to ``max_grad_norm``."""

    def __init__(self, max_grad_norm: float):
        self.max_norm = float(max_grad_norm)

    @torch.no_grad()
    def step(self, parameters) -> torch.Tensor:
        grads: List[torch.Tensor] = [p.grad for p in parameters if p.grad is not None]
        if not grads:
            return torch.zeros(())
        total, _ = multi_tensor_l2norm(grads, per_tensor=False)
        coef = torch.clamp(self.max_norm / (total + 1e-6), max=1.0)   # stays on device
        multi_tensor_scale(grads, grads, coef)
        return total
