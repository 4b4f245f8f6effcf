#!/usr/bin/env python
"""Format a raw corpus into shards with one sentence per line and a blank line between articles -- same CLI as
the reference's utils/format.py."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bert_pytorch_b200.data import corpus  # noqa: E402

if __name__ == "__main__":
    p = argparse.ArgumentParser()
    p.add_argument("--input_dir", type=str, required=True, help="wikiextractor output dir / books dir")
    p.add_argument("--output_dir", type=str, required=True)
    p.add_argument("--dataset", type=str, required=True, choices=["wikicorpus", "bookscorpus"])
    p.add_argument("--processes", type=int, default=8)
    p.add_argument("--shards", type=int, default=64, help="number of output shards (-1: one per input file)")
    a = p.parse_args()
    corpus.format_corpus(a.dataset, a.input_dir, a.output_dir, a.processes, a.shards)
