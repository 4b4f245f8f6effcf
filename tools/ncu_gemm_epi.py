"""One launch of the pair GEMM per fused epilogue at its BERT-large shape inside a profiler window:
  ncu --set full --profile-from-start off --clock-control none --import-source on -o gpurun_out/prof_gemm_epi python tools/ncu_gemm_epi.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bert_pytorch_b200.ops import api as K  # noqa: E402

M, H, I = 12288, 1024, 4096
x = torch.randn(M, H, device="cuda").bfloat16()
wo = (torch.randn(H, H, device="cuda") * 0.02).bfloat16()
w1 = (torch.randn(I, H, device="cuda") * 0.02).bfloat16()
w2 = (torch.randn(H, I, device="cuda") * 0.02).bfloat16()
b1, bo = torch.zeros(I, device="cuda").bfloat16(), torch.zeros(H, device="cuda").bfloat16()
res = torch.randn(M, H, device="cuda").bfloat16()
dy = torch.randn(M, H, device="cuda").bfloat16()
gp = torch.rand(M, I, device="cuda").bfloat16()
aux = torch.empty(M, I, device="cuda", dtype=torch.bfloat16)
cs = torch.zeros(I, device="cuda")
gw = torch.zeros(I, H, device="cuda")
act = torch.randn(M, I, device="cuda").bfloat16()


def run():
    K.gemm(x, wo, epi=K.EPI_NONE)                                                                      # plain
    bits = K.dropout_mask(M, H, 0.1, 3, 2, "cuda")                                                     # keep bits of the site
    K.gemm(x, wo, epi=K.EPI_BIAS_DROP_RES, bias=bo, res=res, p_drop=0.1, seed=3, stream=2, mask_in=bits)   # attn-out
    K.gemm(x, w1, epi=K.EPI_BIAS_GELU_DG, bias=b1, aux_out=aux)                                        # FFN-1
    K.gemm(dy, w2, layout=K.NN, epi=K.EPI_MUL, res=gp, colsum=cs)                                      # FFN-2 dgrad
    K.wgrad_accumulate(act, x, gw)                                                                     # FFN-1 wgrad (TN)


run(); run()
torch.cuda.synchronize()
torch.cuda.profiler.start()
run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
