"""BASELINE config #1: phase-1 style pre-training on CPU/gloo, world_size 1, 4 synthetic HDF5 shards, 10 steps
(+ resume equivalence, output layout, checkpoint rolling window)."""
import csv
import json
import os

import pytest
import torch

from bert_pytorch_b200 import pretrain
from bert_pytorch_b200.data import synthetic
from bert_pytorch_b200.utils import checkpoint as ck


def _workspace(root, **kw):
    return synthetic.make_workspace(str(root), num_shards=4, samples_per_shard=40, seq_len=128, vocab_size=1000,
                                    hidden_size=64, num_hidden_layers=2, num_attention_heads=4, intermediate_size=128,
                                    max_position_embeddings=128, hidden_dropout_prob=0.0,
                                    attention_probs_dropout_prob=0.0, **kw)


def _argv(data, model_json, out, steps, max_steps=10, extra=()):
    cfg = {"model_config_file": model_json, "max_predictions_per_seq": 20, "learning_rate": 6e-3,
           "warmup_proportion": 0.2843, "global_batch_size": 16, "local_batch_size": 8, "max_steps": max_steps,
           "num_steps_per_checkpoint": 2, "log_prefix": "pretraining_phase1_log", "disable_progress_bar": True}
    path = os.path.join(os.path.dirname(out), f"train_{steps}.json")
    with open(path, "w") as f:
        json.dump(cfg, f)
    return ["--config_file", path, "--input_dir", data, "--output_dir", out, "--steps", str(steps), "--device", "cpu",
            *extra]


def test_phase1_cpu_gloo_10_steps_layout_and_resume(tmp_path):
    data, model_json, _ = _workspace(tmp_path)
    out_a = str(tmp_path / "a" / "out"); os.makedirs(os.path.dirname(out_a))
    pretrain.cli(_argv(data, model_json, out_a, steps=10))
    # ---- output layout
    assert os.path.isfile(os.path.join(out_a, "pretraining_phase1_log.txt"))
    rows = list(csv.DictReader(open(os.path.join(out_a, "pretraining_phase1_log_metrics.csv"))))
    assert [int(r["step"]) for r in rows] == list(range(1, 11))
    for k in ("tag", "epoch", "average_loss", "step_loss", "learning_rate", "samples_per_second"):
        assert k in rows[0]
    assert float(rows[-1]["average_loss"]) < float(rows[0]["average_loss"]) + 1.0
    ckpts = ck.list_checkpoints(os.path.join(out_a, "pretrain_ckpts"))
    assert [s for s, _ in ckpts] == [6, 8, 10]                       # rolling window of 3
    payload = torch.load(ckpts[-1][1], map_location="cpu", weights_only=False)
    assert set(payload) == {"model", "optimizer", "sampler", "epoch", "scaler"}
    assert set(payload["sampler"]) == {"epoch", "seed", "num_replicas", "total_size", "index"}
    assert "bert.encoder.layer.0.attention.self.query.weight" in payload["model"]
    assert payload["optimizer"]["param_groups"][0]["step"] == 10
    assert set(payload["optimizer"]["state"][0]) >= {"exp_avg", "exp_avg_sq"}

    # ---- resume equivalence: 5 + 5 steps == 10 steps
    out_b = str(tmp_path / "b" / "out"); os.makedirs(os.path.dirname(out_b))
    pretrain.cli(_argv(data, model_json, out_b, steps=5))
    assert ck.find_latest(os.path.join(out_b, "pretrain_ckpts"))[0] == 5
    pretrain.cli(_argv(data, model_json, out_b, steps=5))
    pa = torch.load(ck.find_latest(os.path.join(out_a, "pretrain_ckpts"))[1], map_location="cpu", weights_only=False)
    pb = torch.load(ck.find_latest(os.path.join(out_b, "pretrain_ckpts"))[1], map_location="cpu", weights_only=False)
    assert pa["sampler"]["index"] == pb["sampler"]["index"]
    worst = max((pa["model"][k].float() - pb["model"][k].float()).abs().max().item() for k in pa["model"])
    # identical data order and identical masks (the mask RNG is keyed by the batch position): only the optimizer
    # state round trip through the checkpoint separates the two runs
    assert worst < 1e-5, worst


def test_phase2_style_step_surgery(tmp_path):
    data, model_json, _ = _workspace(tmp_path)
    out = str(tmp_path / "p" / "out"); os.makedirs(os.path.dirname(out))
    pretrain.cli(_argv(data, model_json, out, steps=4, max_steps=4))
    argv = _argv(data, model_json, out, steps=2, max_steps=3, extra=["--previous_phase_end_step", "4",
                                                                     "--learning_rate", "1e-3"])
    pretrain.cli(argv)
    ckpts = [s for s, _ in ck.list_checkpoints(os.path.join(out, "pretrain_ckpts"))]
    assert ckpts[-1] == 6                                             # names continue at 4 + k
    payload = torch.load(ck.find_latest(os.path.join(out, "pretrain_ckpts"))[1], map_location="cpu", weights_only=False)
    assert payload["optimizer"]["param_groups"][0]["step"] == 2       # schedule restarted, moments carried over
    with pytest.raises(ValueError):
        pretrain.cli(_argv(data, model_json, out, steps=1, extra=["--previous_phase_end_step", "99"]))


def test_required_arguments():
    with pytest.raises(ValueError):
        pretrain.check_required(pretrain.parse_arguments(["--output_dir", "x"]))


def test_kfac_flag_and_multi_rank_gloo(tmp_path):
    """--kfac on the CLI path (preconditioner built, stepped, checkpointed) and a 2-rank gloo launch of the same
    runtime through torchrun (sampler chunking, gradient all-reduce, rank-0-only outputs)."""
    import subprocess
    import sys
    data, model_json, _ = _workspace(tmp_path)
    out = str(tmp_path / "k" / "out"); os.makedirs(os.path.dirname(out))
    pretrain.cli(_argv(data, model_json, out, steps=3, max_steps=3, extra=["--kfac", "--kfac_inv_interval", "1",
                                                                           "--kfac_factor_interval", "1"]))
    payload = torch.load(ck.find_latest(os.path.join(out, "pretrain_ckpts"))[1], map_location="cpu", weights_only=False)
    assert "preconditioner" in payload and payload["preconditioner"]["steps"] == 3
    assert payload["preconditioner"]["layers"]                  # factors were collected through the module hooks

    out2 = str(tmp_path / "g" / "out"); os.makedirs(os.path.dirname(out2))
    argv = _argv(data, model_json, out2, steps=2, max_steps=2)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29733", os.path.join(root, "run_pretraining.py"), *argv]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=root,
                         env=dict(os.environ, OMP_NUM_THREADS="1"))
    assert res.returncode == 0, res.stderr[-3000:]
    assert ck.find_latest(os.path.join(out2, "pretrain_ckpts"))[0] == 2
    rows = list(csv.DictReader(open(os.path.join(out2, "pretraining_phase1_log_metrics.csv"))))
    assert [int(r["step"]) for r in rows] == [1, 2]
