import argparse
import json
import math

import pytest
import torch

from bert_pytorch_b200.config import BertConfig, batch_arithmetic, overlay_json_config
from bert_pytorch_b200.optim import schedulers as S
from bert_pytorch_b200.pretrain import build_parser, parse_arguments


def test_bertconfig_roundtrip_and_extra_keys(tmp_path):
    p = tmp_path / "m.json"
    p.write_text(json.dumps({"vocab_size": 30522, "hidden_size": 1024, "num_attention_heads": 16,
                             "vocab_file": "/x/vocab.txt", "tokenizer": "wordpiece", "next_sentence": False}))
    c = BertConfig.from_json_file(str(p))
    assert c.vocab_file == "/x/vocab.txt" and c.tokenizer == "wordpiece" and c.next_sentence is False
    assert c.pad_vocab(8).vocab_size == 30528
    assert json.loads(c.to_json_string())["hidden_size"] == 1024
    assert BertConfig(30522).vocab_size == 30522
    with pytest.raises(ValueError):
        BertConfig(1000, hidden_size=100, num_attention_heads=7)


def test_cli_json_default_precedence(tmp_path):
    cfg = tmp_path / "t.json"
    cfg.write_text(json.dumps({"learning_rate": 6e-3, "local_batch_size": 96, "not_a_flag": 1, "lr_decay": "linear",
                               "kfac_skip_layers": ["a"]}))
    a = parse_arguments(["--config_file", str(cfg), "--local_batch_size", "7"])
    assert a.learning_rate == 6e-3          # JSON beats default
    assert a.local_batch_size == 7          # CLI beats JSON
    assert a.lr_decay == "linear"
    assert a.max_predictions_per_seq == 80  # default survives
    assert not hasattr(a, "not_a_flag")     # unknown keys ignored
    assert a.kfac_skip_layers == ["a"]


@pytest.mark.parametrize("gb,lb,ws,acc", [
    (65536, 96, 1, 683), (65536, 96, 2, 342), (65536, 96, 4, 171), (65536, 96, 8, 86),
    (32768, 16, 8, 256), (8192, 16, 8, 64), (32768, 16, 1, 2048)])
def test_batch_arithmetic_table(gb, lb, ws, acc):
    assert batch_arithmetic(gb, lb, ws)[1] == acc


class _Opt:
    def __init__(self, lr, step=None):
        self.param_groups = [{"lr": lr}, {"lr": lr}]
        if step is not None:
            for g in self.param_groups:
                g["step"] = step


def test_poly_linear_follow_optimizer_step():
    o = _Opt(6e-3)
    s = S.PolyWarmUpScheduler(o, warmup=0.2843, total_steps=7038)
    assert s.last_epoch == 1                       # no 'step' key -> 1 (quirk Q18)
    assert math.isclose(o.param_groups[0]["lr"], 6e-3 * (1 / 7038) / 0.2843)
    o.param_groups[0]["step"] = 3000
    s.step()
    prog = 3001 / 7038
    assert math.isclose(o.param_groups[1]["lr"], 6e-3 * (1 - prog) ** 0.5)
    o2 = _Opt(4e-4, step=50000)
    l = S.LinearWarmUpScheduler(o2, warmup=0.06, total_steps=100000)
    prog = 50001 / 100000
    assert math.isclose(o2.param_groups[0]["lr"], 4e-4 * max((prog - 1) / (0.06 - 1), 0))
    o2.param_groups[0]["step"] = 100
    l.step()
    assert math.isclose(o2.param_groups[0]["lr"], 4e-4 * (101 / 100000) / 0.06)


def test_cosine_constant_and_legacy_schedules():
    o = _Opt(1.0)
    c = S.CosineWarmUpScheduler(o, warmup=0.1, total_steps=100)
    lrs = []
    for _ in range(100):
        c.step()
        lrs.append(o.param_groups[0]["lr"])
    assert max(lrs) <= 1.0 + 1e-9 and lrs[-1] < 0.01 and lrs[8] > lrs[3]   # warm-up then decay to ~0
    k = S.ConstantWarmUpScheduler(_Opt(2.0), warmup=0.5, total_steps=10)
    for _ in range(9):
        k.step()
    assert k.get_last_lr()[0] == 2.0
    assert S.warmup_linear(0.05, 0.1) == 0.5 and S.warmup_linear(1.0, 0.1) == 0.0
    assert math.isclose(S.warmup_poly(0.75, 0.1), 0.5)
    assert S.SCHEDULES["warmup_constant"](0.9, 0.1) == 1.0
    sd = c.state_dict()
    c2 = S.CosineWarmUpScheduler(_Opt(1.0), warmup=0.1, total_steps=100)
    c2.load_state_dict(sd)
    assert c2.last_epoch == c.last_epoch
