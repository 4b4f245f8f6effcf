"""Model + training configuration.

Behavioural parity targets (reference file:line):
  * ``BertConfig`` -- src/modeling.py:188-280 (int-or-json ctor, every JSON key is
    copied onto the instance so tokenizer keys ride along, to_json_string).
  * training config precedence CLI > JSON > argparse defaults, unknown JSON keys
    silently ignored -- run_pretraining.py:159-172.
  * batch arithmetic -- run_pretraining.py:213-228.
"""
from __future__ import annotations

import argparse
import copy
import json
import math
import os
from typing import Any, Dict, Optional, Sequence


class BertConfig:
    """Architecture hyper-parameters of a BERT encoder (+ tokenizer side-info)."""

    _DEFAULTS: Dict[str, Any] = dict(
        vocab_size=30522,
        hidden_size=768,
        num_hidden_layers=12,
        num_attention_heads=12,
        intermediate_size=3072,
        hidden_act="gelu",
        hidden_dropout_prob=0.1,
        attention_probs_dropout_prob=0.1,
        max_position_embeddings=512,
        type_vocab_size=2,
        initializer_range=0.02,
        next_sentence=True,
        output_all_encoded_layers=False,
    )

    def __init__(self, vocab_size_or_config_json_file: Any = None, **kwargs: Any):
        for k, v in self._DEFAULTS.items():
            setattr(self, k, v)
        src = vocab_size_or_config_json_file
        if isinstance(src, (str, os.PathLike)):
            with open(src, "r", encoding="utf-8") as f:
                for k, v in json.load(f).items():
                    setattr(self, k, v)
        elif isinstance(src, int):
            self.vocab_size = src
        elif src is not None:
            raise ValueError("first argument must be a vocab size (int) or a path to a JSON config")
        for k, v in kwargs.items():
            setattr(self, k, v)
        if self.hidden_size % self.num_attention_heads != 0:
            raise ValueError(
                f"hidden_size={self.hidden_size} is not a multiple of "
                f"num_attention_heads={self.num_attention_heads}")

    # -- constructors -----------------------------------------------------
    @classmethod
    def from_dict(cls, obj: Dict[str, Any]) -> "BertConfig":
        cfg = cls(vocab_size_or_config_json_file=int(obj.get("vocab_size", -1)))
        for k, v in obj.items():
            setattr(cfg, k, v)
        return cfg

    @classmethod
    def from_json_file(cls, path: str) -> "BertConfig":
        with open(path, "r", encoding="utf-8") as f:
            return cls.from_dict(json.load(f))

    # -- helpers ----------------------------------------------------------
    @property
    def head_dim(self) -> int:
        return self.hidden_size // self.num_attention_heads

    def pad_vocab(self, multiple: int = 8) -> "BertConfig":
        """Round vocab up so the MLM decoder GEMM is tile friendly
        (run_pretraining.py:237-238)."""
        r = self.vocab_size % multiple
        if r:
            self.vocab_size += multiple - r
        return self

    def to_dict(self) -> Dict[str, Any]:
        return copy.deepcopy(self.__dict__)

    def to_json_string(self) -> str:
        return json.dumps(self.to_dict(), indent=2, sort_keys=True) + "\n"

    def __repr__(self) -> str:
        return self.to_json_string()


# ---------------------------------------------------------------------------
# CLI / JSON overlay
# ---------------------------------------------------------------------------

def overlay_json_config(parser: argparse.ArgumentParser, argv: Optional[Sequence[str]] = None,
                        config_key: str = "config_file") -> argparse.Namespace:
    """Parse ``argv`` with precedence CLI > JSON file > parser defaults.

    A second pass with every default suppressed tells us which flags were
    really typed by the user; only the others may be overridden by the JSON.
    Keys in the JSON that the parser does not know are ignored.
    """
    args = parser.parse_args(argv)
    probe = argparse.ArgumentParser(argument_default=argparse.SUPPRESS, add_help=False)
    for action in parser._actions:  # mirror arity so nargs/flags parse identically
        if not action.option_strings or action.dest == "help":
            continue
        kw: Dict[str, Any] = dict(dest=action.dest)
        if isinstance(action, (argparse._StoreTrueAction, argparse._StoreFalseAction)):
            kw["action"] = "store_true"
        elif action.nargs is not None:
            kw["nargs"] = action.nargs
        probe.add_argument(*action.option_strings, **kw)
    typed, _ = probe.parse_known_args(argv)
    path = getattr(args, config_key, None)
    if path is not None:
        with open(path, "r", encoding="utf-8") as f:
            overrides = json.load(f)
        for k, v in overrides.items():
            if hasattr(args, k) and not hasattr(typed, k):
                setattr(args, k, v)
    return args


def batch_arithmetic(global_batch_size: int, local_batch_size: int, world_size: int):
    """(local_accumulated_batch_size, accumulation_steps) -- ceil math of
    run_pretraining.py:218-228."""
    local_acc = math.ceil(global_batch_size / world_size)
    acc_steps = math.ceil(local_acc / local_batch_size)
    return local_acc, acc_steps
