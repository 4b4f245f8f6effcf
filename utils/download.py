#!/usr/bin/env python
"""NLP dataset downloader (wikicorpus, bookscorpus, squad, sst-2, mrpc, weights) -- same CLI as the reference's
utils/download.py.  Needs network access."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bert_pytorch_b200.data import corpus  # noqa: E402

if __name__ == "__main__":
    p = argparse.ArgumentParser(description="NLP Dataset Downloader")
    p.add_argument("--dir", type=str, required=True, help="Directory to download datasets to.")
    p.add_argument("--datasets", type=str, required=True, nargs="+",
                   choices=["wikicorpus", "bookscorpus", "squad", "sst-2", "mprc", "mrpc", "weights"])
    a = p.parse_args()
    print(f'Downloading {a.datasets} to "{a.dir}"')
    for d in a.datasets:
        corpus.download(d, a.dir)
    print("Finished downloading")
