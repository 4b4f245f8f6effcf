#!/usr/bin/env python
"""BERT pre-training CLI -- same flags, JSON overlay, output layout and checkpoint format as
the reference's run_pretraining.py; the implementation is bert_pytorch_b200/pretrain.py."""
from bert_pytorch_b200.pretrain import cli

if __name__ == "__main__":
    cli()
