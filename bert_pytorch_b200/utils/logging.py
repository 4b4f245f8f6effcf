"""In-repo replacement for the two logging packages the reference pulls in (neither is
installable here):

* loggerplus style (run_pretraining.py:21,191-204,554-564): ``init(handlers=[...])``,
  ``info(msg)``, ``log(tag=, step=, **metrics)`` fanned out to Stream / File /
  TensorBoard / CSV handlers, each with a ``verbose`` switch so only rank 0 emits.
* dllogger style (run_squad.py:45,890-895,1206-1228): ``JSONStreamBackend`` +
  ``StdOutBackend``, ``dllog.log(step=, data=)``, ``flush()``.

Output file names/format follow SURVEY.md 2.5.2: ``<prefix>.txt``, ``<prefix>_metrics.csv``,
``tensorboard/``.
"""
from __future__ import annotations

import csv
import datetime as _dt
import json
import os
import sys
import time
from typing import Any, Dict, Iterable, List, Optional


def _now() -> str:
    return _dt.datetime.now().strftime("%Y-%m-%d %H:%M:%S")


def _fmt_metrics(tag, step, metrics: Dict[str, Any]) -> str:
    parts = []
    if tag is not None:
        parts.append(f"[{tag}]")
    if step is not None:
        parts.append(f"step={step}")
    for k, v in metrics.items():
        parts.append(f"{k}={v:.6g}" if isinstance(v, float) else f"{k}={v}")
    return " ".join(parts)


class Handler:
    def __init__(self, verbose: bool = True):
        self.verbose = verbose

    def info(self, msg: str) -> None: ...
    def log(self, tag, step, metrics: Dict[str, Any]) -> None: ...
    def flush(self) -> None: ...
    def close(self) -> None: ...


class StreamHandler(Handler):
    def __init__(self, stream=None, verbose: bool = True):
        super().__init__(verbose)
        self.stream = stream or sys.stdout

    def info(self, msg):
        if self.verbose:
            print(f"{_now()} {msg}", file=self.stream, flush=True)

    def log(self, tag, step, metrics):
        self.info(_fmt_metrics(tag, step, metrics))


class FileHandler(Handler):
    def __init__(self, path: str, overwrite: bool = False, verbose: bool = True):
        super().__init__(verbose)
        self.path = path
        self._fh = None
        if verbose:
            os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
            self._fh = open(path, "w" if overwrite else "a", encoding="utf-8")

    def info(self, msg):
        if self._fh:
            self._fh.write(f"{_now()} {msg}\n")
            self._fh.flush()

    def log(self, tag, step, metrics):
        self.info(_fmt_metrics(tag, step, metrics))

    def close(self):
        if self._fh:
            self._fh.close()
            self._fh = None


class CSVHandler(Handler):
    """One row per ``log`` call; the header is (re)written when the column set grows."""

    def __init__(self, path: str, overwrite: bool = False, verbose: bool = True):
        super().__init__(verbose)
        self.path = path
        self.fields: List[str] = []
        if verbose:
            os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
            if overwrite and os.path.exists(path):
                os.remove(path)
            if os.path.exists(path) and os.path.getsize(path) > 0:
                with open(path, newline="", encoding="utf-8") as f:
                    self.fields = next(csv.reader(f), [])

    def log(self, tag, step, metrics):
        if not self.verbose:
            return
        row = {"tag": tag, "step": step, **metrics}
        new = [k for k in row if k not in self.fields]
        if new:
            old_rows = []
            if self.fields and os.path.exists(self.path):
                with open(self.path, newline="", encoding="utf-8") as f:
                    old_rows = list(csv.DictReader(f))
            self.fields += new
            with open(self.path, "w", newline="", encoding="utf-8") as f:
                w = csv.DictWriter(f, fieldnames=self.fields)
                w.writeheader()
                w.writerows(old_rows)
        with open(self.path, "a", newline="", encoding="utf-8") as f:
            csv.DictWriter(f, fieldnames=self.fields).writerow(row)


class TorchTensorboardHandler(Handler):
    def __init__(self, logdir: str, verbose: bool = True):
        super().__init__(verbose)
        self.writer = None
        if verbose:
            try:
                from torch.utils.tensorboard import SummaryWriter
                self.writer = SummaryWriter(logdir)
            except Exception as e:  # tensorboard missing -> degrade to nothing
                print(f"[logging] tensorboard disabled: {e}", file=sys.stderr)

    def log(self, tag, step, metrics):
        if self.writer is None:
            return
        for k, v in metrics.items():
            if isinstance(v, (int, float)):
                self.writer.add_scalar(f"{tag}/{k}" if tag else k, v, step)

    def flush(self):
        if self.writer:
            self.writer.flush()

    def close(self):
        if self.writer:
            self.writer.close()
            self.writer = None


_HANDLERS: List[Handler] = [StreamHandler()]


def init(handlers: Optional[Iterable[Handler]] = None) -> None:
    global _HANDLERS
    for h in _HANDLERS:
        h.close()
    _HANDLERS = list(handlers) if handlers is not None else [StreamHandler()]


def info(msg: Any) -> None:
    for h in _HANDLERS:
        h.info(str(msg))


def warning(msg: Any) -> None:
    info(f"WARNING: {msg}")


def log(tag: Optional[str] = None, step: Optional[int] = None, **metrics: Any) -> None:
    for h in _HANDLERS:
        h.log(tag, step, metrics)


def flush() -> None:
    for h in _HANDLERS:
        h.flush()


def close() -> None:
    init([StreamHandler()])


# ---------------------------------------------------------------------------
# dllogger-compatible facade
# ---------------------------------------------------------------------------


class Verbosity:
    OFF, DEFAULT, VERBOSE = -1, 0, 1


class JSONStreamBackend:
    def __init__(self, verbosity: int, filename: str):
        self.verbosity = verbosity
        self._fh = open(filename, "w", encoding="utf-8")

    def log(self, step, data, verbosity):
        if verbosity <= self.verbosity:
            rec = {"type": "LOG", "datetime": _now(), "elapsedtime": f"{time.perf_counter():.6f}",
                   "step": step if not isinstance(step, tuple) else list(step), "data": data}
            self._fh.write("DLLL " + json.dumps(rec, default=str) + "\n")

    def flush(self):
        self._fh.flush()


class StdOutBackend:
    def __init__(self, verbosity: int, step_format=None):
        self.verbosity, self.step_format = verbosity, step_format

    def log(self, step, data, verbosity):
        if verbosity <= self.verbosity:
            s = self.step_format(step) if self.step_format else str(step)
            print(f"DLL {_now()} - {s} {' '.join(f'{k} : {v} ' for k, v in data.items())}", flush=True)

    def flush(self):
        sys.stdout.flush()


class DLLogger:
    def __init__(self):
        self.backends: list = []

    def init(self, backends):
        self.backends = list(backends)

    def log(self, step, data, verbosity: int = Verbosity.DEFAULT):
        for b in self.backends:
            b.log(step, data, verbosity)

    def flush(self):
        for b in self.backends:
            b.flush()


dllogger = DLLogger()
