"""One launch of the optimizer / loss / embedding / LayerNorm kernels at BERT-large size inside a profiler window:
  ncu --set full --profile-from-start off --clock-control none --import-source on -o gpurun_out/prof_misc python tools/ncu_misc.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bert_pytorch_b200 import BertConfig  # noqa: E402
from bert_pytorch_b200.models import BertForPreTraining  # noqa: E402
from bert_pytorch_b200.models.arena import NO_DECAY_KEYS, ParamArena  # noqa: E402
from bert_pytorch_b200.ops import api as K  # noqa: E402
from bert_pytorch_b200.optim import Lamb  # noqa: E402

dev = "cuda"
cfg = BertConfig(vocab_size_or_config_json_file=30528, hidden_size=1024, num_hidden_layers=24, num_attention_heads=16,
                 intermediate_size=4096, max_position_embeddings=512)
model = BertForPreTraining(cfg).to(dev)
arena = ParamArena(model, device=torch.device(dev))
named = list(model.named_parameters())
opt = Lamb([{"params": [p for n, p in named if not any(k in n for k in NO_DECAY_KEYS)], "weight_decay": 0.01},
            {"params": [p for n, p in named if any(k in n for k in NO_DECAY_KEYS)], "weight_decay": 0.0}], lr=1e-3)
arena.bind_optimizer(opt)
M, H, V = 12288, 1024, 30528
ids = torch.randint(5, V, (M,), device=dev, dtype=torch.int32)
seg = torch.zeros(M, device=dev, dtype=torch.int32)
de = torch.randn(M, H, device=dev).bfloat16()
x = torch.randn(M, H, device=dev).bfloat16()
g, b = torch.ones(H, device=dev), torch.zeros(H, device=dev)
dg, db, dbias = (torch.zeros(H, device=dev) for _ in range(3))
logits = torch.randn(1920, V, device=dev).bfloat16()
tgt = torch.randint(0, V, (1920,), device=dev, dtype=torch.int32)
cnt = torch.tensor([1920], device=dev, dtype=torch.int32)
loss = torch.zeros(1, device=dev)
gw = arena.grad("bert.embeddings.word_embeddings.weight")
gp = arena.grad("bert.embeddings.position_embeddings.weight")
gt = arena.grad("bert.embeddings.token_type_embeddings.weight")


def run():
    arena.flat_grad.normal_(0, 1e-3)
    opt.step()                                   # flat_sumsq + lamb_stage1 + lamb_stage2
    K.softmax_ce_(logits.clone(), tgt, cnt, 1.0, loss)
    K.embedding_bwd_scatter(de, ids, seg, gw, gp, gt, 128)
    y, mean, rstd = K.layer_norm_fwd(x, g, b)
    bits = K.dropout_mask(M, H, 0.1, 3, 5, dev)   # dropout_mask_kernel
    K.layer_norm_bwd(de, x, mean, rstd, g, dgamma=dg, dbeta=db, dbias=dbias, want_dropped=True, p_drop=0.1, seed=3, drop_stream=5,
                     keep_mask=bits)


run(); run()
torch.cuda.synchronize()
torch.cuda.profiler.start()
run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
