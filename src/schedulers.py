"""Warm-up LR schedulers (reference src/schedulers.py) -> bert_pytorch_b200.optim.schedulers."""
import bert_pytorch_b200.optim.schedulers as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
