#!/usr/bin/env python
"""SQuAD fine-tuning / prediction CLI -- same flags and outputs as the reference run_squad.py; implementation in bert_pytorch_b200/finetune_squad.py."""
from bert_pytorch_b200.finetune_squad import main

if __name__ == "__main__":
    main()
