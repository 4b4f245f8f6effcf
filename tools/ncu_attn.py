"""One launch of each attention kernel (phase-1 and phase-2 shapes, dropout on) inside a profiler window:
  ncu --set full --profile-from-start off --clock-control none --import-source on -o gpurun_out/prof_attn python tools/ncu_attn.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bert_pytorch_b200.ops import api as K  # noqa: E402

H, h = 1024, 16
cases = []
for B, S in ((96, 128),) if os.environ.get('NCU_PHASE1_ONLY') else ((96, 128), (16, 512)):
    qkv = (torch.randn(B, S, 3 * H, device="cuda") * 0.5).bfloat16()
    lens = torch.full((B,), S, device="cuda", dtype=torch.int32)
    cases.append((qkv, lens))


def run():
    for qkv, lens in cases:
        ctx, lse = K.attention_fwd(qkv, lens, h, p_drop=0.1, seed=1, stream=1)
        K.attention_bwd(qkv, lens, ctx, torch.ones_like(ctx), lse, h, p_drop=0.1, seed=1, stream=1)


run(); run()
torch.cuda.synchronize()
torch.cuda.profiler.start()
run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
