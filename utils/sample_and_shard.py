#!/usr/bin/env python
"""Randomly sample articles up to a sentence budget and write size-limited shards -- same CLI as the
reference's utils/sample_and_shard.py."""
import argparse
import os
import random
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bert_pytorch_b200.data import corpus  # noqa: E402

if __name__ == "__main__":
    p = argparse.ArgumentParser(description="Text file sampler + sharder")
    p.add_argument("-i", "--input", type=str, required=True, help="input file or directory of *.txt")
    p.add_argument("-o", "--output", type=str, required=True)
    p.add_argument("-f", "--format", type=str, default="shard_{index}.txt")
    p.add_argument("-b", "--size", type=str, required=True, help="maximum bytes per shard")
    p.add_argument("-n", "--sentences", type=str, required=True, help="total number of sentences to sample")
    p.add_argument("--seed", type=int, default=None)
    a = p.parse_args()
    t0 = time.time()
    files = corpus.find_txt_files(a.input)
    print(f"[sampler] Found {len(files)} input files")
    n = corpus.sample_and_shard(files, os.path.join(a.output, a.format), corpus.parse_value_as_int(a.size),
                                corpus.parse_value_as_int(a.sentences), random.Random(a.seed))
    print(f"[sampler] Finished sampling and sharding: {n} shards (time={time.time() - t0:.1f})")
