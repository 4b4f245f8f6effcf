"""Dynamic loss scaler with the state-dict layout of ``torch.cuda.amp.GradScaler``
(keys ``scale, growth_factor, backoff_factor, growth_interval, _growth_tracker`` --
saved in checkpoints by run_pretraining.py:513-523 and read back at :317-318).

Differences from torch's: it works on any device (CPU/gloo plumbing config), and the
``inv_scale`` / ``found_inf`` device scalars are handed to our fused optimisers so
unscale + overflow check + skipped step run inside the optimiser kernels with no
host synchronisation.  In bf16 mode construct it with
``enabled=False``: every method becomes the identity but ``state_dict`` keeps its
keys so the checkpoint layout does not change.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch


class GradScaler:
    def __init__(self, init_scale: float = 2.0 ** 16, growth_factor: float = 2.0,
                 backoff_factor: float = 0.5, growth_interval: int = 2000, enabled: bool = True,
                 device: Optional[torch.device] = None):
        self._enabled = enabled
        self._init_scale = float(init_scale)
        self._growth_factor, self._backoff_factor = float(growth_factor), float(backoff_factor)
        self._growth_interval = int(growth_interval)
        self._device = torch.device(device) if device is not None else None
        self._scale: Optional[torch.Tensor] = None
        self._growth_tracker: Optional[torch.Tensor] = None
        self._found_inf: Optional[torch.Tensor] = None
        self._unscaled: set = set()
        self._init_growth_tracker = 0

    # -- lazy state ---------------------------------------------------------
    def _lazy_init(self, device: torch.device) -> None:
        if self._scale is None:
            dev = self._device or device
            self._scale = torch.full((), self._init_scale, dtype=torch.float32, device=dev)
            self._growth_tracker = torch.full((), self._init_growth_tracker, dtype=torch.int32, device=dev)
            self._found_inf = torch.zeros((), dtype=torch.float32, device=dev)

    def is_enabled(self) -> bool:
        return self._enabled

    def get_scale(self) -> float:
        if not self._enabled:
            return 1.0
        return self._init_scale if self._scale is None else float(self._scale)

    @property
    def found_inf(self) -> Optional[torch.Tensor]:
        return self._found_inf

    def inv_scale(self) -> torch.Tensor:
        return self._scale.reciprocal()

    # -- API ------------------------------------------------------------------
    def scale(self, loss: torch.Tensor) -> torch.Tensor:
        if not self._enabled:
            return loss
        self._lazy_init(loss.device)
        return loss * self._scale.to(loss.dtype)

    def unscale_(self, optimizer) -> None:
        """grads *= 1/scale and record inf/nan into ``found_inf`` (idempotent per step)."""
        if not self._enabled or id(optimizer) in self._unscaled:
            return
        arena = getattr(optimizer, "_arena", None)
        dev = None
        for group in optimizer.param_groups:
            for p in group["params"]:
                dev = p.device
                break
            if dev is not None:
                break
        self._lazy_init(dev)
        inv = self.inv_scale()
        if arena is not None and arena.fused_optimizer_ok():
            arena.unscale_(inv, self._found_inf)
        else:
            for group in optimizer.param_groups:
                for p in group["params"]:
                    if p.grad is None:
                        continue
                    p.grad.mul_(inv.to(p.grad.dtype))
                    if not bool(torch.isfinite(p.grad).all()):
                        self._found_inf.fill_(1.0)
        self._unscaled.add(id(optimizer))

    def step(self, optimizer, *args, **kwargs):
        if not self._enabled:
            return optimizer.step(*args, **kwargs)
        fused = getattr(optimizer, "_arena", None) is not None and optimizer._arena.fused_optimizer_ok()
        if fused and id(optimizer) not in self._unscaled:
            # unscale + inf check + conditional skip all inside the optimiser kernels
            dev = optimizer.param_groups[0]["params"][0].device
            self._lazy_init(dev)
            out = optimizer.step(*args, inv_scale=self.inv_scale(), found_inf=self._found_inf, **kwargs)
            self._unscaled.add(id(optimizer))
            return out
        self.unscale_(optimizer)
        if fused:
            return optimizer.step(*args, found_inf=self._found_inf, **kwargs)
        if float(self._found_inf) == 0.0:
            return optimizer.step(*args, **kwargs)
        return None

    def update(self, new_scale: Optional[float] = None) -> None:
        if not self._enabled or self._scale is None:
            return
        if new_scale is not None:
            self._scale.fill_(float(new_scale))
        else:
            # device-side, branch free: no .item() on the step path
            inf = self._found_inf > 0
            tracker = torch.where(inf, torch.zeros_like(self._growth_tracker), self._growth_tracker + 1)
            grow = tracker >= self._growth_interval
            scale = torch.where(inf, self._scale * self._backoff_factor,
                                torch.where(grow, self._scale * self._growth_factor, self._scale))
            self._scale.copy_(scale)
            self._growth_tracker.copy_(torch.where(grow, torch.zeros_like(tracker), tracker))
        self._found_inf.zero_()
        self._unscaled.clear()

    # -- (de)serialisation ------------------------------------------------------
    def state_dict(self) -> Dict[str, float]:
        return {
            "scale": self.get_scale() if self._enabled else self._init_scale,
            "growth_factor": self._growth_factor,
            "backoff_factor": self._backoff_factor,
            "growth_interval": self._growth_interval,
            "_growth_tracker": (int(self._growth_tracker) if self._growth_tracker is not None
                                else self._init_growth_tracker),
        }

    def load_state_dict(self, state: Dict[str, float]) -> None:
        if not state:
            return
        self._init_scale = float(state["scale"])
        self._growth_factor = float(state["growth_factor"])
        self._backoff_factor = float(state["backoff_factor"])
        self._growth_interval = int(state["growth_interval"])
        self._init_growth_tracker = int(state["_growth_tracker"])
        if self._scale is not None:
            self._scale.fill_(self._init_scale)
            self._growth_tracker.fill_(self._init_growth_tracker)
