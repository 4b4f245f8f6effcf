"""h5py stand-in: the read-only subset the reference's dataset uses, on this repo's pure-Python HDF5
reader (data loading only -- no model/kernel code)."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.append(_ROOT)
from bert_pytorch_b200.data.hdf5 import File  # noqa: E402,F401
