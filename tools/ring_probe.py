"""How much does a deeper operand ring buy the K = 1024 GEMMs?  Same NT problem through the fp32-store epilogue with
the 5-stage and the 7-stage ring (B200_GEMM_DEEP_RING=0/1, read once per process -> run twice)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bert_pytorch_b200.ops import api as K  # noqa: E402
from tools.gemm_bench import timeit  # noqa: E402

for name, m, n, k in (("ffn1_fwd", 12288, 4096, 1024), ("attn_out", 12288, 1024, 1024), ("ffn2_fwd", 12288, 1024, 4096)):
    a, b = torch.randn(m, k, device="cuda").bfloat16(), torch.randn(n, k, device="cuda").bfloat16()
    o = torch.empty(m, n, device="cuda")
    t = timeit(lambda: K.gemm(a, b, out=o, epi=K.EPI_F32, block_n=512))
    print(json.dumps({"shape": name, "deep_ring": os.environ.get("B200_GEMM_DEEP_RING", "1"), "epi": "F32 store",
                      "tflops": round(2.0 * m * n * k / t / 1e9, 1)}), flush=True)
