#!/usr/bin/env python
"""NER fine-tuning CLI -- same flags as the reference run_ner.py; implementation in bert_pytorch_b200/finetune_ner.py."""
from bert_pytorch_b200.finetune_ner import main

if __name__ == "__main__":
    main()
