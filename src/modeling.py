"""Models (reference src/modeling.py) -> bert_pytorch_b200.models."""
import bert_pytorch_b200.models as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
