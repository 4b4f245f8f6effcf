from .lamb import Lamb, FusedLAMB, lamb_reference_step  # noqa: F401
from .adam import Adam, FusedAdam, BertAdam  # noqa: F401
from .schedulers import (  # noqa: F401
    LRScheduler, CosineWarmUpScheduler, ConstantWarmUpScheduler, LinearWarmUpScheduler,
    PolyWarmUpScheduler, SCHEDULES,
    warmup_cosine, warmup_constant, warmup_linear, warmup_poly)
from .grad_scaler import GradScaler  # noqa: F401
from .clip import GradientClipper, multi_tensor_l2norm, multi_tensor_scale  # noqa: F401
