"""Safe wrappers around torch.distributed (src/utils.py:29-74) plus rendezvous helpers.
They work before/without ``init_process_group`` so single-process CPU runs need no setup."""
from __future__ import annotations

import os
from pathlib import Path
from typing import Optional

import torch
import torch.distributed as dist


def is_initialized() -> bool:
    return dist.is_available() and dist.is_initialized()


def get_rank() -> int:
    return dist.get_rank() if is_initialized() else 0


def get_world_size() -> int:
    return dist.get_world_size() if is_initialized() else 1


def get_local_rank() -> int:
    return int(os.environ.get("LOCAL_RANK", 0))


def is_main_process() -> bool:
    return get_rank() == 0


def barrier() -> None:
    if is_initialized():
        dist.barrier()


def format_step(step) -> str:
    """dllogger step formatter (src/utils.py:54-64)."""
    if isinstance(step, str):
        return step
    s = ""
    if len(step) > 0:
        s += f"Training Epoch: {step[0]} "
    if len(step) > 1:
        s += f"Training Iteration: {step[1]} "
    if len(step) > 2:
        s += f"Validation Iteration: {step[2]} "
    return s


def mkdir(path: str) -> None:
    Path(path).mkdir(parents=True, exist_ok=True)


def mkdir_by_main_process(path: str) -> None:
    if is_main_process():
        mkdir(path)
    barrier()


def init_distributed(backend: Optional[str] = None, device: Optional[torch.device] = None) -> str:
    """``env://`` rendezvous (torchrun exports RANK/WORLD_SIZE/MASTER_*).  ``backend=None``
    picks nccl on CUDA and gloo on CPU -- the reference hard-codes nccl and asserts CUDA
    (run_pretraining.py:181,185), so the CPU/gloo plumbing configuration is new here."""
    if backend is None:
        backend = "nccl" if (device is not None and device.type == "cuda") or (
            device is None and torch.cuda.is_available()) else "gloo"
    if is_initialized():
        return dist.get_backend()
    if "RANK" not in os.environ or "WORLD_SIZE" not in os.environ:
        # single process: behave like a world of one without touching the network
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        os.environ.setdefault("RANK", "0")
        os.environ.setdefault("WORLD_SIZE", "1")
    kw = {}
    if backend == "nccl" and device is not None and device.type == "cuda":
        kw["device_id"] = device
    dist.init_process_group(backend=backend, init_method="env://", **kw)
    return backend


class WorkerInitObj:
    """``worker_init_fn`` for torch DataLoaders: gives every worker its own numpy / python RNG stream
    (src/utils.py:22-27 defines this helper but the reference never uses it, quirk Q5)."""

    def __init__(self, seed: int):
        self.seed = seed

    def __call__(self, worker_id: int) -> None:
        import random

        import numpy as np
        np.random.seed(seed=self.seed + worker_id)
        random.seed(self.seed + worker_id)
