"""Import the in-tree extension ``_C.so`` (never a JIT cache: the built file must be the one that
travels with the repo snapshot and shows up in the driver's loaded-.so record)."""
from __future__ import annotations

import importlib.util
import os

_HERE = os.path.dirname(os.path.abspath(__file__))


def extension_path() -> str:
    return os.path.join(_HERE, "_C.so")


def load_extension():
    import torch  # noqa: F401  (libtorch must be loaded first so the .so resolves its symbols)
    path = extension_path()
    if not os.path.exists(path):
        raise ImportError(f"{path} not found; build it with `python -m bert_pytorch_b200.ops.build`")
    spec = importlib.util.spec_from_file_location("bert_pytorch_b200.ops._C", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
