"""Synthetic pre-training shards with the exact on-disk schema of the reference encoder
(utils/encode_data.py:183-210): ``input_ids`` int32 gzip,
``special_token_positions`` int32 gzip ([N,3] with NSP, [N,2] without),
``next_sentence_labels`` int8 gzip, files ``train_{i}.hdf5``.  Also emits a vocab file and a
model JSON whose ``vocab_file`` points at it (the shipped JSONs carry absolute cluster
paths).  Used by the tests, the CPU plumbing config and bench.py -- there
is no network for real corpora.
"""
from __future__ import annotations

import json
import os
from typing import List, Optional

import numpy as np

from . import hdf5

SPECIAL_TOKENS = ["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"]


def write_vocab(path: str, vocab_size: int) -> str:
    with open(path, "w", encoding="utf-8") as f:
        for t in SPECIAL_TOKENS:
            f.write(t + "\n")
        for i in range(vocab_size - len(SPECIAL_TOKENS)):
            f.write(f"tok{i}\n")
    return path


def learnable_tokens(n: int, seq_len: int, vocab_size: int, rng: np.random.Generator) -> np.ndarray:
    """A corpus a masked-LM can actually learn (uniform random tokens carry no signal: the best achievable loss is
    ln V): the token at position p comes from one of four 16-token groups selected by ``p mod 4`` (needs the position
    embeddings) with Zipf-distributed frequencies inside the group, and with probability 0.5 it repeats the token four
    positions earlier (needs attention).  Entropy floor ~1.7 nats against ln(1024) = 6.9 for the uniform corpus."""
    base = len(SPECIAL_TOKENS)
    groups = min(4, max(1, (vocab_size - base) // 16))
    w = 1.0 / np.arange(1, 17)
    w /= w.sum()
    pos = np.arange(seq_len)
    tok = rng.choice(16, size=(n, seq_len), p=w).astype(np.int32)
    rep = rng.random((n, seq_len)) < 0.5
    for p_ in range(4, seq_len):
        tok[:, p_] = np.where(rep[:, p_], tok[:, p_ - 4], tok[:, p_])
    return (base + (pos % groups)[None, :] * 16 + tok).astype(np.int32)


def make_samples(n: int, seq_len: int, vocab_size: int, next_sentence: bool, rng: np.random.Generator,
                 short_seq_prob: float = 0.1, learnable: bool = False):
    cls_id, sep_id = SPECIAL_TOKENS.index("[CLS]"), SPECIAL_TOKENS.index("[SEP]")
    ids = np.zeros((n, seq_len), dtype=np.int32)
    nsp = 3 if next_sentence else 2
    sp = np.zeros((n, nsp), dtype=np.int32)
    full = rng.random(n) >= short_seq_prob
    total = np.where(full, seq_len, rng.integers(max(8, seq_len // 8), seq_len + 1, size=n))
    body = (learnable_tokens(n, seq_len, vocab_size, rng) if learnable
            else rng.integers(len(SPECIAL_TOKENS), vocab_size, size=(n, seq_len), dtype=np.int32))
    for i in range(n):
        t = int(total[i])
        ids[i, :t] = body[i, :t]
        ids[i, 0] = cls_id
        ids[i, t - 1] = sep_id
        if next_sentence:
            mid = int(rng.integers(2, t - 2))
            ids[i, mid] = sep_id
            sp[i] = (0, mid, t - 1)
        else:
            sp[i] = (0, t - 1)
    labels = rng.integers(0, 2, size=n).astype(np.int8) if next_sentence else np.zeros(n, dtype=np.int8)
    return ids, sp, labels


def write_shards(out_dir: str, num_shards: int, samples_per_shard: int, seq_len: int, vocab_size: int,
                 next_sentence: bool = True, seed: int = 0, compression: Optional[str] = "gzip",
                 learnable: bool = False) -> List[str]:
    os.makedirs(out_dir, exist_ok=True)
    rng = np.random.default_rng(seed)
    paths = []
    for s in range(num_shards):
        ids, sp, nsl = make_samples(samples_per_shard, seq_len, vocab_size, next_sentence, rng, learnable=learnable)
        p = os.path.join(out_dir, f"train_{s}.hdf5")
        with hdf5.File(p, "w") as f:
            f.create_dataset("input_ids", data=ids, dtype="i4", compression=compression)
            f.create_dataset("special_token_positions", data=sp, dtype="i4", compression=compression)
            f.create_dataset("next_sentence_labels", data=nsl, dtype="i1", compression=compression)
        paths.append(p)
    return paths


def write_model_config(path: str, vocab_file: str, *, vocab_size: int = 30522, hidden_size: int = 1024,
                       num_hidden_layers: int = 24, num_attention_heads: int = 16,
                       intermediate_size: int = 4096, next_sentence: bool = True,
                       max_position_embeddings: int = 512, **extra) -> str:
    cfg = dict(attention_probs_dropout_prob=0.1, hidden_act="gelu", hidden_dropout_prob=0.1,
               hidden_size=hidden_size, initializer_range=0.02, intermediate_size=intermediate_size,
               lowercase=True, max_position_embeddings=max_position_embeddings,
               model_name="synthetic-bert", next_sentence=next_sentence,
               num_attention_heads=num_attention_heads, num_hidden_layers=num_hidden_layers,
               tokenizer="wordpiece", type_vocab_size=2, vocab_size=vocab_size, vocab_file=vocab_file)
    cfg.update(extra)
    with open(path, "w", encoding="utf-8") as f:
        json.dump(cfg, f, indent=2)
    return path


def make_workspace(root: str, *, num_shards: int = 4, samples_per_shard: int = 64, seq_len: int = 128,
                   vocab_size: int = 30522, next_sentence: bool = True, seed: int = 0, learnable: bool = False,
                   **model_kw):
    """data dir + vocab + model json under ``root``; returns (data_dir, model_json, vocab)."""
    os.makedirs(root, exist_ok=True)
    vocab = write_vocab(os.path.join(root, "vocab.txt"), vocab_size)
    data_dir = os.path.join(root, f"sequences_lowercase_max_seq_len_{seq_len}_next_seq_task_"
                                  f"{str(next_sentence).lower()}")
    write_shards(data_dir, num_shards, samples_per_shard, seq_len, vocab_size, next_sentence, seed, learnable=learnable)
    model_json = write_model_config(os.path.join(root, "model_config.json"), vocab, vocab_size=vocab_size,
                                    next_sentence=next_sentence, **model_kw)
    return data_dir, model_json, vocab
