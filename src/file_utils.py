"""Download cache helpers (reference src/file_utils.py) -> bert_pytorch_b200.utils.file_utils."""
import bert_pytorch_b200.utils.file_utils as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
