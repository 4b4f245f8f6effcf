"""Warm-up learning-rate schedules.

Parity: src/schedulers.py:21-158 and the schedule functions of
src/optimization.py:36-62.  The schedules are closed-form functions of
``progress = (optimizer_step + 1) / total_steps``; Poly/Linear read the step counter
the optimiser keeps in ``param_groups[0]['step']`` (that coupling is what lets a
skipped overflow step not advance the schedule) and fall back to
``last_epoch = 1`` when the key is absent (quirk Q18).  The reference's cosine
schedule is unusable (``torch.cos`` on a float, Q17); here it is the intended
half-cosine decay.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional


def _warm(progress: float, warmup: float) -> Optional[float]:
    if warmup > 0 and progress < warmup:
        return progress / warmup
    return None


def cosine_factor(progress: float, warmup: float) -> float:
    w = _warm(progress, warmup)
    if w is not None:
        return w
    t = min(max((progress - warmup) / max(1.0 - warmup, 1e-12), 0.0), 1.0)
    return 0.5 * (1.0 + math.cos(math.pi * t))


def constant_factor(progress: float, warmup: float) -> float:
    w = _warm(progress, warmup)
    return 1.0 if w is None else w


def linear_factor(progress: float, warmup: float) -> float:
    w = _warm(progress, warmup)
    if w is not None:
        return w
    return max((progress - 1.0) / (warmup - 1.0), 0.0)


def poly_factor(progress: float, warmup: float, degree: float = 0.5) -> float:
    w = _warm(progress, warmup)
    if w is not None:
        return w
    return max(1.0 - progress, 0.0) ** degree


# BertAdam-style schedules: f(x, warmup) with x = step / t_total (src/optimization.py:36-62)
def warmup_cosine(x: float, warmup: float = 0.002) -> float:
    return x / warmup if x < warmup else 0.5 * (1.0 + math.cos(math.pi * x))


def warmup_constant(x: float, warmup: float = 0.002) -> float:
    return x / warmup if x < warmup else 1.0


def warmup_linear(x: float, warmup: float = 0.002) -> float:
    return x / warmup if x < warmup else max((x - 1.0) / (warmup - 1.0), 0.0)


def warmup_poly(x: float, warmup: float = 0.002, degree: float = 0.5) -> float:
    return x / warmup if x < warmup else (1.0 - x) ** degree


SCHEDULES: Dict[str, Callable[..., float]] = {
    "warmup_cosine": warmup_cosine, "warmup_constant": warmup_constant,
    "warmup_linear": warmup_linear, "warmup_poly": warmup_poly,
}


def warmup_exp_decay_exp(global_step, decay_rate, decay_steps, total_steps, warmup=0.002, degree=2.0):
    x = global_step / total_steps
    if warmup == 0.0:
        return 1.0
    if x < warmup:
        return (x / warmup) ** degree
    return decay_rate ** ((global_step - warmup * total_steps) / decay_steps)


class LRScheduler:
    """Base: owns ``base_lrs`` and writes ``param_group['lr']`` on every ``step``.
    ``optimizer`` can be anything with ``param_groups`` (an optimiser or the K-FAC
    preconditioner, run_pretraining.py:347-349).  Like the reference (which inherits
    torch's ``_LRScheduler``) construction performs one ``step()``."""

    def __init__(self, optimizer, last_epoch: int = -1):
        if not hasattr(optimizer, "param_groups"):
            raise TypeError(f"{type(optimizer).__name__} has no param_groups")
        self.optimizer = optimizer
        for g in optimizer.param_groups:
            g.setdefault("initial_lr", g["lr"])
        self.base_lrs: List[float] = [g["initial_lr"] for g in optimizer.param_groups]
        self.last_epoch = last_epoch
        self.step()

    def _advance(self, epoch: Optional[int]) -> None:
        self.last_epoch = epoch if epoch is not None else self.last_epoch + 1

    def factor(self) -> float:  # pragma: no cover - abstract
        raise NotImplementedError

    def get_lr(self) -> List[float]:
        f = self.factor()
        return [b * f for b in self.base_lrs]

    def get_last_lr(self) -> List[float]:
        return [g["lr"] for g in self.optimizer.param_groups]

    def step(self, epoch: Optional[int] = None) -> None:
        self._advance(epoch)
        for g, lr in zip(self.optimizer.param_groups, self.get_lr()):
            g["lr"] = lr

    def state_dict(self) -> dict:
        return {k: v for k, v in self.__dict__.items() if k != "optimizer"}

    def load_state_dict(self, state: dict) -> None:
        self.__dict__.update(state)


class _WarmUp(LRScheduler):
    def __init__(self, optimizer, warmup: float, total_steps: float, last_epoch: int = -1):
        self.warmup, self.total_steps = warmup, total_steps
        super().__init__(optimizer, last_epoch)

    @property
    def progress(self) -> float:
        return self.last_epoch / self.total_steps


class _StepCoupled(_WarmUp):
    """last_epoch follows the optimiser's own step counter."""

    def _advance(self, epoch: Optional[int]) -> None:
        g = self.optimizer.param_groups[0]
        self.last_epoch = int(g["step"]) + 1 if "step" in g else 1


class CosineWarmUpScheduler(_WarmUp):
    def factor(self) -> float:
        return cosine_factor(self.progress, self.warmup)


class ConstantWarmUpScheduler(_WarmUp):
    def factor(self) -> float:
        return constant_factor(self.progress, self.warmup)


class LinearWarmUpScheduler(_StepCoupled):
    def factor(self) -> float:
        return linear_factor(self.progress, self.warmup)


class PolyWarmUpScheduler(_StepCoupled):
    def __init__(self, optimizer, warmup, total_steps, degree: float = 0.5, last_epoch: int = -1):
        self.degree = degree
        super().__init__(optimizer, warmup, total_steps, last_epoch)

    def factor(self) -> float:
        return poly_factor(self.progress, self.warmup, self.degree)
