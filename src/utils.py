"""Rank helpers (reference src/utils.py) -> bert_pytorch_b200.utils.dist."""
import bert_pytorch_b200.utils.dist as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
