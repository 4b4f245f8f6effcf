#!/usr/bin/env python
"""Split a formatted text file into size-limited shards at article boundaries -- same CLI as the reference's
utils/shard.py."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bert_pytorch_b200.data import corpus  # noqa: E402

if __name__ == "__main__":
    p = argparse.ArgumentParser(description="Text file sharder")
    p.add_argument("-i", "--input", type=str, required=True)
    p.add_argument("-o", "--output", type=str, required=True, help="output directory")
    p.add_argument("-f", "--format", type=str, default="shard_{index}.txt")
    p.add_argument("-b", "--size", type=str, default="100M", help="maximum bytes per shard (K/M/B suffixes)")
    p.add_argument("-n", "--max_shards", type=int, default=None)
    a = p.parse_args()
    print(f"Sharding {a.input} to {a.output}")
    os.makedirs(a.output, exist_ok=True)
    n = corpus.shard_text(a.input, os.path.join(a.output, a.format), corpus.parse_value_as_int(a.size), a.max_shards)
    print(f"Finished sharding ({n} shards)")
