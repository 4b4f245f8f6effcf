"""End to end on the GPU: the pre-training runtime (loader -> fused engine with CUDA-graph replay -> arena LAMB ->
checkpoint) on synthetic HDF5 shards, and the checkpoint loading back into the plain PyTorch modules."""
import csv
import json
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(tmp_path, extra=()):
    from bert_pytorch_b200 import pretrain
    from bert_pytorch_b200.data import synthetic
    data, model_json, _ = synthetic.make_workspace(
        str(tmp_path), num_shards=4, samples_per_shard=64, seq_len=128, vocab_size=1024, hidden_size=128,
        num_hidden_layers=2, num_attention_heads=2, intermediate_size=512, max_position_embeddings=128)
    out = str(tmp_path / "run" / "out"); os.makedirs(os.path.dirname(out), exist_ok=True)
    cfg = {"model_config_file": model_json, "max_predictions_per_seq": 20, "learning_rate": 4e-3,
           "warmup_proportion": 0.25, "global_batch_size": 32, "local_batch_size": 8, "max_steps": 12,
           "num_steps_per_checkpoint": 4, "log_prefix": "pretraining_phase1_log", "disable_progress_bar": True}
    path = os.path.join(os.path.dirname(out), "train.json")
    with open(path, "w") as f:
        json.dump(cfg, f)
    pretrain.cli(["--config_file", path, "--input_dir", data, "--output_dir", out, "--steps", "12", *extra])
    return out, model_json


@pytest.mark.parametrize("extra", [(), ("--checkpoint_activations",), ("--fp8",)])
def test_pretraining_runtime_on_gpu(tmp_path, extra):
    from bert_pytorch_b200 import BertConfig
    from bert_pytorch_b200.models import BertForPreTraining
    from bert_pytorch_b200.ops import api as K
    from bert_pytorch_b200.utils import checkpoint as ck
    n0 = K.KERNEL_LAUNCHES
    out, model_json = _run(tmp_path, extra)
    assert K.KERNEL_LAUNCHES - n0 > 12 * 4 * 50              # the sm_100a kernel program ran (4 micro-steps per step)
    rows = list(csv.DictReader(open(os.path.join(out, "pretraining_phase1_log_metrics.csv"))))
    assert [int(r["step"]) for r in rows] == list(range(1, 13))
    first, last = float(rows[0]["average_loss"]), float(rows[-1]["average_loss"])
    assert last == last and last < first + 0.5, (first, last)  # uniform random tokens: nothing to learn, nothing may diverge
    step, path = ck.find_latest(os.path.join(out, "pretrain_ckpts"))
    assert step == 12
    payload = torch.load(path, map_location="cpu", weights_only=False)
    cfg = BertConfig.from_json_file(model_json)
    cfg.pad_vocab(8)
    ref = BertForPreTraining(cfg)
    missing, unexpected = ref.load_state_dict(payload["model"], strict=False)
    assert not unexpected and all("position_ids" in m for m in missing)
    assert payload["optimizer"]["param_groups"][0]["step"] == 12
