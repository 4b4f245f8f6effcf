"""No-op stand-in for the loggerplus package (reference arm only)."""
import sys


class _H:
    def __init__(self, *a, **k):
        self.verbose = k.get("verbose", True)


StreamHandler = FileHandler = TorchTensorboardHandler = CSVHandler = _H
_verbose = True


def init(handlers=None):
    global _verbose
    _verbose = any(getattr(h, "verbose", False) for h in (handlers or []))


def info(msg, *a, **k):
    if _verbose:
        print(msg, file=sys.stderr, flush=True)


def log(tag=None, step=None, **metrics):
    pass


def warning(msg, *a, **k):
    info(msg)
