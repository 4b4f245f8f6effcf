"""Device-side timing (the reference only has host wall clocks).

``DeviceTimer`` brackets a region with CUDA events on the launching stream;
``max_over_ranks`` reduces a local duration to the slowest rank, which is what every
multi-GPU number in this repo reports.  ``ClockSampler`` polls nvidia-smi during a timed
region for the bench's ``clocks`` record.  ``L2Flusher`` writes a > L2 sized buffer between
timed iterations.  ``nvtx_range`` is a no-op off CUDA."""
from __future__ import annotations

import contextlib
import statistics
import subprocess
import threading
import time
from typing import Dict, List, Optional

import torch
import torch.distributed as dist


class DeviceTimer:
    def __init__(self, device: Optional[torch.device] = None):
        self.cuda = torch.cuda.is_available() and (device is None or device.type == "cuda")
        self._t0 = 0.0
        if self.cuda:
            self._e0 = torch.cuda.Event(enable_timing=True)
            self._e1 = torch.cuda.Event(enable_timing=True)

    def start(self) -> None:
        if self.cuda:
            torch.cuda.synchronize()
            self._e0.record()
        else:
            self._t0 = time.perf_counter()

    def stop(self) -> float:
        """milliseconds"""
        if self.cuda:
            self._e1.record()
            torch.cuda.synchronize()
            return self._e0.elapsed_time(self._e1)
        return (time.perf_counter() - self._t0) * 1e3


class StepClock:
    """Device-side split of one optimizer step without adding a synchronisation: ``mark()`` records an event on the
    current stream (or the host clock on CPU); ``intervals()`` -- call it after something else has synchronised,
    e.g. the loss read-back -- returns the milliseconds between consecutive marks."""

    def __init__(self, device: Optional[torch.device] = None):
        self.cuda = torch.cuda.is_available() and (device is None or device.type == "cuda")
        self._marks: List = []

    def reset(self) -> None:
        self._marks = []

    def mark(self) -> None:
        if self.cuda:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            self._marks.append(e)
        else:
            self._marks.append(time.perf_counter())

    def intervals(self) -> List[float]:
        m = self._marks
        if self.cuda:
            return [a.elapsed_time(b) for a, b in zip(m[:-1], m[1:])]
        return [(b - a) * 1e3 for a, b in zip(m[:-1], m[1:])]


def max_over_ranks(value: float, device: Optional[torch.device] = None) -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    dev = device or (torch.device("cuda", torch.cuda.current_device())
                     if dist.get_backend() == "nccl" else torch.device("cpu"))
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


class L2Flusher:
    """Overwrites a buffer larger than the 126 MB L2 so the next timed iteration starts cold."""

    def __init__(self, device: torch.device, nbytes: int = 256 << 20):
        self.buf = torch.empty(nbytes // 4, dtype=torch.float32, device=device) if device.type == "cuda" else None

    def flush(self) -> None:
        if self.buf is not None:
            self.buf.fill_(1.0)


class ClockSampler:
    """Samples SM clocks / throttle reasons with nvidia-smi while a region runs."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int = 0, period_ms: int = 200):
        self.gpu_index, self.period_ms = gpu_index, period_ms
        self.samples: List[List[str]] = []
        self._proc: Optional[subprocess.Popen] = None
        self._thr: Optional[threading.Thread] = None

    def start(self) -> None:
        try:
            self._proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu_index}", f"--query-gpu={self.QUERY}",
                 "--format=csv,noheader,nounits", "-lms", str(self.period_ms)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except (OSError, FileNotFoundError):
            self._proc = None
            return

        def pump():
            assert self._proc is not None and self._proc.stdout is not None
            for line in self._proc.stdout:
                parts = [p.strip() for p in line.split(",")]
                if len(parts) >= 8:
                    self.samples.append(parts)
        self._thr = threading.Thread(target=pump, daemon=True)
        self._thr.start()

    def stop(self) -> Dict[str, object]:
        if self._proc is not None:
            self._proc.terminate()
            try:
                self._proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self._proc.kill()
        if self._thr is not None:
            self._thr.join(timeout=2)
        sm = []
        reasons = set()
        sm_max = None
        power = []
        for s in self.samples:
            try:
                sm.append(float(s[1])); sm_max = float(s[2]); power.append(float(s[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": sm_max,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


@contextlib.contextmanager
def nvtx_range(name: str):
    if torch.cuda.is_available():
        torch.cuda.nvtx.range_push(name)
        try:
            yield
        finally:
            torch.cuda.nvtx.range_pop()
    else:
        yield
