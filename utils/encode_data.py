#!/usr/bin/env python
"""Encode formatted text shards (one sentence per line, blank line between documents) into the HDF5
pre-training shards read by run_pretraining.py.  Same CLI as the reference's utils/encode_data.py; the
work is done by bert_pytorch_b200/data/encode.py with the repo's native HDF5 writer."""
import argparse
import multiprocessing as mp
import os
import sys
import time
from pathlib import Path

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bert_pytorch_b200.data import encode as E  # noqa: E402


def main(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument("--input_dir", required=True, help="directory with .txt files or a single file")
    p.add_argument("--output_dir", required=True, help="output directory for the hdf5 files")
    p.add_argument("--vocab_file", required=True)
    p.add_argument("--max_seq_len", default=512, type=int)
    p.add_argument("--short_seq_prob", default=0.1, type=float)
    p.add_argument("--next_seq_prob", default=0.0, type=float, help="0 disables next-sentence prediction")
    p.add_argument("--uppercase", action="store_true", default=False)
    p.add_argument("--tokenizer", default="wordpiece", choices=["wordpiece", "bpe"])
    p.add_argument("--processes", type=int, default=4)
    p.add_argument("--seed", type=int, default=None)
    a = p.parse_args(argv)
    t0 = time.time()
    if os.path.isfile(a.input_dir):
        files = [a.input_dir]
    elif os.path.isdir(a.input_dir):
        files = sorted(str(x) for x in Path(a.input_dir).rglob("*.txt") if x.is_file())
    else:
        raise ValueError(f"{a.input_dir} is not a valid path")
    print(f"[encoder] Found {len(files)} input files")
    out_dir = os.path.join(a.output_dir, E.output_dir_name(a.uppercase, a.max_seq_len, a.next_seq_prob > 0))
    os.makedirs(out_dir, exist_ok=True)
    jobs = [(f, os.path.join(out_dir, f"train_{i}.hdf5"), a.vocab_file, a.tokenizer, a.uppercase, a.max_seq_len,
             a.next_seq_prob, a.short_seq_prob, None if a.seed is None else a.seed + i) for i, f in enumerate(files)]
    print(f"[encoder] Starting multiprocessing pool ({a.processes} processes)")
    if a.processes <= 1 or len(jobs) <= 1:
        for j in jobs:
            E.encode_file(*j)
    else:
        with mp.Pool(processes=a.processes) as pool:
            pool.starmap(E.encode_file, jobs)
    print(f"[encoder] Finished processing (time={time.time() - t0:.0f}s)")


if __name__ == "__main__":
    main()
