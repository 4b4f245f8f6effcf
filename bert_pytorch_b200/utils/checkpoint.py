"""Checkpoint layout + resume logic of the reference (SURVEY.md 2.5.2 / 5.4;
run_pretraining.py:243-265, 298-309, 496-533):

``<output_dir>/pretrain_ckpts/ckpt_<N>.pt`` with ``N = global_step + previous_phase_end_step``;
a dict ``{model, optimizer, sampler, epoch[, preconditioner][, scaler]}`` written by rank 0;
the newest ``N`` wins on resume; only the last three files written by *this* session are
kept; on a phase change the optimizer's step/lr/warmup/t_total are overridden from the
current arguments while the LAMB moments carry over.
"""
from __future__ import annotations

import os
import re
from typing import Any, Dict, List, Optional, Tuple

import torch

_CKPT_RE = re.compile(r"^ckpt_(\d+)\.pt$")


def checkpoint_path(ckpt_dir: str, step: int) -> str:
    return os.path.join(ckpt_dir, f"ckpt_{int(step)}.pt")


def list_checkpoints(ckpt_dir: str) -> List[Tuple[int, str]]:
    if not os.path.isdir(ckpt_dir):
        return []
    out = []
    for f in os.listdir(ckpt_dir):
        m = _CKPT_RE.match(f)
        if m:
            out.append((int(m.group(1)), os.path.join(ckpt_dir, f)))
    return sorted(out)


def find_latest(ckpt_dir: str) -> Optional[Tuple[int, str]]:
    c = list_checkpoints(ckpt_dir)
    return c[-1] if c else None


def load_latest(ckpt_dir: str) -> Tuple[Optional[Dict[str, Any]], int]:
    """(checkpoint dict or None, resume_step)."""
    latest = find_latest(ckpt_dir)
    if latest is None:
        return None, 0
    step, path = latest
    return torch.load(path, map_location="cpu", weights_only=False), step


def override_optimizer_hparams(ckpt: Dict[str, Any], *, global_steps: int, max_steps, warmup: float,
                               lr: float) -> None:
    """Optimizer-state surgery applied on resume (run_pretraining.py:298-308)."""
    opt = ckpt["optimizer"]
    for st in opt["state"].values():
        st["step"] = global_steps
    for g in opt["param_groups"]:
        g["step"] = global_steps
        g["t_total"] = max_steps
        g["warmup"] = warmup
        g["lr"] = lr
        # the reference leaves the scheduler's ``initial_lr`` of the old phase in place, so its new schedule keeps
        # decaying from the OLD base LR although ``lr`` was just overwritten (tests/test_reference_parity.py);
        # dropping it makes the new phase start from the configured learning rate
        g.pop("initial_lr", None)


class CheckpointManager:
    """Rolling window of the checkpoints written in this session."""

    def __init__(self, ckpt_dir: str, keep: int = 3):
        self.ckpt_dir, self.keep = ckpt_dir, keep
        self.written: List[str] = []

    def save(self, step: int, payload: Dict[str, Any]) -> str:
        os.makedirs(self.ckpt_dir, exist_ok=True)
        path = checkpoint_path(self.ckpt_dir, step)
        tmp = path + ".tmp"
        torch.save(payload, tmp)
        os.replace(tmp, path)           # never leave a half written ckpt_N.pt behind
        self.written.append(path)
        while len(self.written) > self.keep:
            old = self.written.pop(0)
            if old != path and os.path.exists(old):
                os.remove(old)
        return path
