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
    for k in ("tag", "epoch", "average_loss", "step_loss", "learning_rate", "samples_per_second",
              "device_step_ms", "optimizer_ms"):
        assert k in rows[0]
    assert all(0.0 < float(r["optimizer_ms"]) < float(r["device_step_ms"]) for r in rows)
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


def test_offline_pipeline_feeds_pretraining(tmp_path):
    """scripts/create_datasets.sh in miniature, through the CLIs: wikiextractor-style + books text -> utils/format.py
    -> utils/build_vocab.py -> utils/encode_data.py (NSP shards) -> run_pretraining.py on the result."""
    import random
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    rnd = random.Random(0)
    words = ("the quick brown fox jumps over lazy dog while rain keeps falling on quiet city streets and people walk "
             "home with umbrellas under grey clouds thinking about dinner plans music science history").split()

    def sent():
        return " ".join(rnd.choice(words) for _ in range(rnd.randint(5, 14))).capitalize() + "."

    wiki = tmp_path / "dl" / "wikicorpus" / "AA"; wiki.mkdir(parents=True)
    with open(wiki / "wiki_00", "w") as f:
        for d in range(24):
            f.write(f'<doc id="{d}" url="u" title="T{d}">\nT{d}\n\n')
            for _ in range(3):
                f.write(" ".join(sent() for _ in range(5)) + "\n")
            f.write("</doc>\n")
    books = tmp_path / "dl" / "books"; books.mkdir(parents=True)
    for b in range(4):                  # two books per shard: NSP needs >= 2 documents in a file
        with open(books / f"book{b}.txt", "w", encoding="ISO-8859-1") as f:
            for _ in range(16):
                f.write(" ".join(sent() for _ in range(4)) + "\n")

    def run(*cmd):
        r = subprocess.run([sys.executable, *cmd], cwd=root, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        return r.stdout

    fmt, vocab, enc = str(tmp_path / "fmt"), str(tmp_path / "vocab" / "wp.txt"), str(tmp_path / "enc")
    run("utils/format.py", "--dataset", "wikicorpus", "--input_dir", str(wiki.parent), "--output_dir", fmt + "/wiki",
        "--processes", "1", "--shards", "1")
    run("utils/format.py", "--dataset", "bookscorpus", "--input_dir", str(books), "--output_dir", fmt + "/books",
        "--processes", "1", "--shards", "2")
    run("utils/build_vocab.py", "-i", fmt, "-o", vocab, "-s", "200")
    tokens = open(vocab).read().splitlines()
    assert tokens[0] == "[PAD]" and {"[MASK]", "[SEP]", "[CLS]", "[UNK]"} <= set(tokens[:5])
    out = run("utils/encode_data.py", "--input_dir", fmt, "--output_dir", enc, "--vocab_file", vocab, "--max_seq_len", "64",
              "--next_seq_prob", "0.5", "--short_seq_prob", "0.1", "--processes", "1", "--seed", "1")
    shard_dir = os.path.join(enc, "sequences_lowercase_max_seq_len_64_next_seq_task_true")
    assert len([f for f in os.listdir(shard_dir) if f.endswith(".hdf5")]) == 3, out

    model_json = str(tmp_path / "model.json")
    json.dump({"attention_probs_dropout_prob": 0.1, "hidden_act": "gelu", "hidden_dropout_prob": 0.1, "hidden_size": 32,
               "initializer_range": 0.02, "intermediate_size": 64, "max_position_embeddings": 64,
               "num_attention_heads": 2, "num_hidden_layers": 2, "type_vocab_size": 2, "vocab_size": len(tokens),
               "next_sentence": True, "vocab_file": vocab, "tokenizer": "wordpiece", "lowercase": True},
              open(model_json, "w"))
    train_json = str(tmp_path / "train.json")
    json.dump({"model_config_file": model_json, "max_predictions_per_seq": 10, "masked_token_fraction": 0.15,
               "learning_rate": 1e-3, "global_batch_size": 16, "local_batch_size": 8, "max_steps": 3,
               "disable_progress_bar": True}, open(train_json, "w"))
    outdir = str(tmp_path / "out")
    pretrain.cli(["--config_file", train_json, "--input_dir", shard_dir, "--output_dir", outdir, "--device", "cpu"])
    assert ck.find_latest(os.path.join(outdir, "pretrain_ckpts"))[0] == 3
