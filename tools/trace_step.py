"""In-situ kernel time table of the flagship training micro-step (torch.profiler / CUPTI activity records:
real clocks, no serialisation) -> gpurun_out/trace_step.json + a printed table."""
import collections
import json
import os
import re
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from bert_pytorch_b200 import BertConfig, pretrain  # noqa: E402
from bert_pytorch_b200.models import BertForPreTraining, BertPretrainingCriterion  # noqa: E402
from bert_pytorch_b200.models.arena import NO_DECAY_KEYS, ParamArena  # noqa: E402
from bert_pytorch_b200.optim import GradScaler, Lamb  # noqa: E402
from bert_pytorch_b200.parallel import DataParallel  # noqa: E402

phase = int(os.environ.get("PHASE", "1"))
ph = bench.PHASES[phase]
dev = torch.device("cuda", 0)
cfg = BertConfig.from_dict(dict(bench.MODEL, next_sentence=True, hidden_act="gelu", hidden_dropout_prob=0.1,
                                attention_probs_dropout_prob=0.1))
cfg.pad_vocab(8)
cfg.max_predictions_per_seq = ph["max_pred"]
model = BertForPreTraining(cfg).to(dev)
arena = ParamArena(model, device=dev)
ddp = DataParallel(model, arena=arena)
named = list(model.named_parameters())
opt = Lamb([{"params": [p for n, p in named if not any(k in n for k in NO_DECAY_KEYS)], "weight_decay": 0.01},
            {"params": [p for n, p in named if any(k in n for k in NO_DECAY_KEYS)], "weight_decay": 0.0}], lr=1e-3)
arena.bind_optimizer(opt)
scaler = GradScaler(enabled=False)
crit = BertPretrainingCriterion(cfg.vocab_size)
pool = [[t.to(dev) for t in b] for b in bench.synth_batches(4, ph["local_batch"], ph["seq"], 30522, ph["max_pred"], 1, torch.int32)]
model.train()


def micro(i):
    return pretrain.forward_backward_pass(ddp, crit, scaler, pool[i % 4], 1, sync_grads=False, compute_dtype=torch.bfloat16)


for i in range(6):
    micro(i)
torch.cuda.synchronize()
N = 4
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    for i in range(N):
        micro(i)
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
ordered = []
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        name = re.sub(r"\(.*", "", ev.name)[:64]
        ordered.append((ev.time_range.start, name, ev.device_time if hasattr(ev, "device_time") else ev.cuda_time))
        agg[name][0] += 1
        agg[name][1] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
tot = sum(v[1] for v in agg.values())
rows = sorted(agg.items(), key=lambda kv: -kv[1][1])
print(f"phase {phase}: kernel time per micro-step {tot / N / 1e3:.2f} ms")
for n, (c, t) in rows[:24]:
    print(f"{t / N:10.1f} us {100 * t / tot:5.1f}%  x{c // N:4d}  avg {t / c:7.1f}  {n}")
os.makedirs("gpurun_out", exist_ok=True)
ordered.sort()
per_step = len(ordered) // N
last = ordered[-per_step:]
# wall span of the last micro-step on the device vs the time kernels were running: what launch gaps / tails cost
t0 = last[0][0]
end, busy, gaps = t0, 0.0, []
for st, n, t in last:
    if st > end:
        gaps.append((st - end, n))
    busy += max(0.0, st + t - max(end, st))
    end = max(end, st + t)
span = end - t0
print(f"last micro-step: span {span / 1e3:.2f} ms, some kernel running {busy / 1e3:.2f} ms, idle {100 * (span - busy) / span:.1f}% "
      f"in {len(gaps)} gaps (median {sorted(g for g, _ in gaps)[len(gaps) // 2] if gaps else 0:.1f} us)")
json.dump({"phase": phase, "span_ms_last_step": span / 1e3, "busy_ms_last_step": busy / 1e3,
           "ordered_last_step": [[n, round(t, 1), round(st - t0, 1)] for st, n, t in last], "ms_per_micro_step_kernels": tot / N / 1e3,
           "kernels": [{"name": n, "calls": c // N, "us_per_step": t / N} for n, (c, t) in rows]},
          open(f"gpurun_out/trace_step_p{phase}.json", "w"), indent=1)
