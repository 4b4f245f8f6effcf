"""torch-native stand-ins for apex's amp_C multi-tensor kernels (reference arm only)."""
import torch


def multi_tensor_l2norm(chunk_size, noop_flag, tensor_lists, per_tensor=False):
    tensors = tensor_lists[0]
    norms = torch._foreach_norm(tensors)
    stacked = torch.stack(norms)
    return torch.linalg.vector_norm(stacked).reshape(1), (stacked if per_tensor else stacked.new_zeros(0))


def multi_tensor_scale(chunk_size, noop_flag, tensor_lists, scale):
    src, dst = tensor_lists
    for s, d in zip(src, dst):
        d.copy_(s * scale)
        if not torch.isfinite(d).all():
            noop_flag.fill_(1)


def multi_tensor_lamb_stage1_cuda(*a, **k):
    raise NotImplementedError("bound but never called by the reference (src/optimization.py:30-33)")


multi_tensor_lamb_stage2_cuda = multi_tensor_lamb_stage1_cuda
