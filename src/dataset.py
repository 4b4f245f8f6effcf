"""HDF5 shard dataset + resumable sampler (reference src/dataset.py) -> bert_pytorch_b200.data.dataset."""
import bert_pytorch_b200.data.dataset as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
