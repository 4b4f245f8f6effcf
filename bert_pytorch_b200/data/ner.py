"""CoNLL-style NER data (reference: src/ner_dataset.py:1-110).

File format: one token per line, columns separated by space/tab, token = column 0, tag = column 3; blank
lines and ``-DOCSTART`` lines separate sentences.  Encoding per sentence: every word is expanded to its
word pieces and its tag replicated on each piece; ``[CLS]``/``[SEP]`` get label -100 (ignored by the loss);
padding gets label 0 and mask 0; label ids start at 1 in the order given on the command line.

Difference: samples are tokenised once at construction (the reference re-tokenises on every
``__getitem__``) and the module's self-test actually runs (quirk Q24).
"""
from __future__ import annotations

import re
from typing import Dict, List, Sequence, Tuple

import torch

from .squad import TokenizerAdapter

SPECIAL_LABEL = "[SPC]"


class Sample:
    def __init__(self, sentence: Sequence[str], labels: Sequence[str]):
        if len(sentence) != len(labels):
            raise ValueError("sentence and labels differ in length")
        self.sentence, self.labels = list(sentence), list(labels)

    def encoded(self, tokenizer, label_to_id: Dict[str, int], max_seq_len: int):
        tok = tokenizer if isinstance(tokenizer, TokenizerAdapter) else TokenizerAdapter(tokenizer)
        pieces: List[str] = []
        tags: List[str] = []
        for word, tag in zip(self.sentence, self.labels):
            sub = tok.tokens(word)
            pieces.extend(sub)
            tags.extend([tag] * len(sub))
        pieces, tags = pieces[:max_seq_len - 2], tags[:max_seq_len - 2]
        pieces = ["[CLS]"] + pieces + ["[SEP]"]
        tags = [SPECIAL_LABEL] + tags + [SPECIAL_LABEL]
        ids = [tok.token_to_id(t) for t in pieces]
        lab = [-100 if t == SPECIAL_LABEL else label_to_id[t] for t in tags]
        mask = [1] * len(ids)
        pad = max_seq_len - len(ids)
        return pieces, tags, ids + [0] * pad, lab + [0] * pad, mask + [0] * pad


def parse_conll(filename: str, token_col: int = 0, label_col: int = 3) -> List[Sample]:
    samples: List[Sample] = []
    words: List[str] = []
    tags: List[str] = []
    with open(filename, "r", encoding="utf-8") as f:
        for line in f:
            if line.strip() == "" or line.startswith("-DOCSTART"):
                if words:
                    samples.append(Sample(words, tags))
                    words, tags = [], []
                continue
            cols = [c.strip() for c in re.split(r" |\t", line) if c.strip() != ""]
            words.append(cols[token_col])
            tags.append(cols[label_col] if len(cols) > label_col else cols[-1])
    if words:
        samples.append(Sample(words, tags))
    return samples


class NERDataset(torch.utils.data.Dataset):
    def __init__(self, filename: str, tokenizer, labels: Sequence[str], max_seq_len: int):
        self.samples = parse_conll(filename)
        self.tokenizer = tokenizer
        self.label_to_idx = {label: i for i, label in enumerate(labels, start=1)}
        self.max_seq_len = max_seq_len
        enc = [s.encoded(tokenizer, self.label_to_idx, max_seq_len) for s in self.samples]
        self.ids = torch.tensor([e[2] for e in enc], dtype=torch.long).view(-1, max_seq_len)
        self.labels = torch.tensor([e[3] for e in enc], dtype=torch.long).view(-1, max_seq_len)
        self.mask = torch.tensor([e[4] for e in enc], dtype=torch.long).view(-1, max_seq_len)

    def __len__(self) -> int:
        return len(self.samples)

    def __getitem__(self, idx: int) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        return self.ids[idx], self.labels[idx], self.mask[idx]


if __name__ == "__main__":      # smoke test: python -m bert_pytorch_b200.data.ner FILE VOCAB LABEL...
    import sys
    from .tokenization import get_wordpiece_tokenizer
    ds = NERDataset(sys.argv[1], get_wordpiece_tokenizer(sys.argv[2]), sys.argv[3:], 32)
    for i in range(min(3, len(ds))):
        print(ds.samples[i].encoded(ds.tokenizer, ds.label_to_idx, 32)[:2], ds[i])
