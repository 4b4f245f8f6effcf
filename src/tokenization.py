"""Tokenizers (reference src/tokenization.py) -> bert_pytorch_b200.data.tokenization."""
import bert_pytorch_b200.data.tokenization as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
