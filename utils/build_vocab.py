#!/usr/bin/env python
"""Train a WordPiece / byte-level BPE vocabulary -- same CLI as the reference's utils/build_vocab.py."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bert_pytorch_b200.data import corpus  # noqa: E402

if __name__ == "__main__":
    p = argparse.ArgumentParser(description="Vocabulary Generator")
    p.add_argument("-i", "--input", type=str, required=True)
    p.add_argument("-o", "--output", type=str, required=True)
    p.add_argument("-s", "--size", type=int, default=30000)
    p.add_argument("--tokenizer", type=str, default="wordpiece", choices=["wordpiece", "bpe"])
    p.add_argument("--uppercase", action="store_true", default=False)
    p.add_argument("--special_tokens", nargs="+", default=["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"])
    p.add_argument("--pad_token", type=str, default="[PAD]")
    a = p.parse_args()
    print("Starting training", flush=True)
    v = corpus.build_vocab(corpus.find_txt_files(a.input), a.output, a.size, a.tokenizer, a.uppercase,
                           a.special_tokens, a.pad_token)
    print(f"Vocab written to file ({len(v)} entries)", flush=True)
