"""BertAdam + warm-up functions (reference src/optimization.py) -> bert_pytorch_b200.optim."""
from bert_pytorch_b200.optim.adam import BertAdam, FusedAdam  # noqa: F401
from bert_pytorch_b200.optim.clip import multi_tensor_l2norm, multi_tensor_scale  # noqa: F401
from bert_pytorch_b200.optim.schedulers import (  # noqa: F401
    SCHEDULES, warmup_constant, warmup_cosine, warmup_linear, warmup_poly)
