"""Text shards -> pre-training samples -> HDF5 shards (library behind ``utils/encode_data.py``).

Behavioural parity with the reference encoder (utils/encode_data.py:12-210):
input = one sentence per line, blank line between documents; every sample is
``[CLS] A [SEP]`` or ``[CLS] A [SEP] B [SEP]`` zero-padded to ``max_seq_len``; consecutive sentences of a
document are packed greedily up to a target length that is the maximum (minus specials) or, with
probability ``short_seq_prob``, uniform in [2, max]; with NSP the packed chunk is cut at a random sentence
boundary and, with probability ``next_seq_prob``, segment B is replaced by text from a random *other*
document (label 1) and the displaced sentences are re-used for the next sample; samples are shuffled inside
a file; datasets ``input_ids`` (int32), ``special_token_positions`` (int32, [N,2] or [N,3]),
``next_sentence_labels`` (int8), all gzip; output dir
``sequences_{lower|upper}case_max_seq_len_{N}_next_seq_task_{true|false}/train_{i}.hdf5``.

Implementation notes: sentences are converted to token *ids* once, samples are built directly as id lists
and written through this repo's native HDF5 writer (no h5py).
"""
from __future__ import annotations

import os
import random
import time
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import hdf5

Sentence = List[int]
Document = List[Sentence]


@dataclass
class TrainingSample:
    seq_ids: List[int]
    next_seq_ids: Optional[List[int]] = None
    is_random_next: bool = False

    def layout(self, cls_id: int, sep_id: int) -> Tuple[List[int], List[int]]:
        """(token ids incl. specials, positions of the special tokens)."""
        ids = [cls_id] + list(self.seq_ids)
        special = [0]
        if self.next_seq_ids is not None:
            special.append(len(ids))
            ids.append(sep_id)
            ids.extend(self.next_seq_ids)
        special.append(len(ids))
        ids.append(sep_id)
        return ids, special


def read_documents(path: str, tokenizer, fast=None, batch_lines: int = 8192) -> List[Document]:
    """Documents = blank-line separated groups of lines; every line becomes a list of token ids.  With ``fast``
    (:class:`~.tokenization.FastWordPiece`) the lines covered by its character table (Latin scripts, punctuation) are tokenised in bulk by the native C++
    WordPiece (identical ids to the ``tokenizers`` package, tests/test_dataset.py), the rest by ``tokenizer``."""
    docs: List[Document] = [[]]
    pending: List[str] = []          # lines of the current batch; None marks a document boundary

    def flush():
        lines = [l for l in pending if l is not None]
        if fast is not None:
            enc = fast.encode_batch(lines, fallback=lambda t: tokenizer.encode(t, add_special_tokens=False).ids)
        else:
            enc = [tokenizer.encode(l, add_special_tokens=False).ids for l in lines]
        it = iter(enc)
        for l in pending:
            if l is None:
                if docs[-1]:
                    docs.append([])
                continue
            ids = next(it)
            if len(ids):
                docs[-1].append(list(ids))
        pending.clear()

    with open(path, "r", encoding="utf-8", errors="ignore") as f:
        for line in f:
            line = line.strip()
            pending.append(line if line else None)
            if len(pending) >= batch_lines:
                flush()
    flush()
    return [d for d in docs if d]


class SamplePacker:
    def __init__(self, max_seq_len: int, next_seq_prob: float, short_seq_prob: float, rng: Optional[random.Random] = None):
        self.nsp = next_seq_prob > 0
        self.next_seq_prob, self.short_seq_prob = next_seq_prob, short_seq_prob
        self.budget = max_seq_len - (3 if self.nsp else 2)
        self.rng = rng or random.Random()

    def _target(self) -> int:
        if self.rng.random() < self.short_seq_prob:
            return self.rng.randint(2, self.budget)
        return self.budget

    def _random_tail(self, docs: Sequence[Document], avoid: int, room: int) -> List[int]:
        other = self.rng.randrange(len(docs))
        while other == avoid:
            other = self.rng.randrange(len(docs))
        doc = docs[other]
        out: List[int] = []
        for sent in doc[self.rng.randrange(len(doc)):]:
            out.extend(sent)
            if len(out) >= room:
                break
        return out[:room]

    def pack_document(self, docs: Sequence[Document], d: int) -> List[TrainingSample]:
        if self.nsp and len(docs) <= 1:
            raise ValueError("File only contained one document, unable to make a random next sequence.")
        doc = docs[d]
        samples: List[TrainingSample] = []
        target = self._target()
        chunk: List[Sentence] = []
        size = 0
        i = 0
        while i < len(doc):
            sent = doc[i][:target]
            closing = bool(chunk) and (i + 1 == len(doc) or size + len(sent) >= target)
            if closing:
                if self.nsp:
                    cut = self.rng.randint(1, len(chunk) - 1) if len(chunk) >= 2 else 1
                    a = [t for s in chunk[:cut] for t in s]
                    b = [t for s in chunk[cut:] for t in s]
                    rand_next = self.rng.random() < self.next_seq_prob
                    if rand_next:
                        b = self._random_tail(docs, d, target - len(a))
                        i -= len(chunk) - cut          # give the displaced sentences another chance
                    samples.append(TrainingSample(a[:target], b[:max(target - len(a), 0)], rand_next))
                else:
                    samples.append(TrainingSample([t for s in chunk for t in s][:target]))
                target = self._target()
                chunk, size = [], 0
                sent = doc[i][:target]
            chunk.append(sent)
            size += len(sent)
            i += 1
        return samples

    def pack_file(self, docs: Sequence[Document]) -> List[TrainingSample]:
        out: List[TrainingSample] = []
        for d in range(len(docs)):
            out.extend(self.pack_document(docs, d))
        self.rng.shuffle(out)
        return out


def samples_to_arrays(samples: Sequence[TrainingSample], max_seq_len: int, cls_id: int, sep_id: int):
    nsp = any(s.next_seq_ids is not None for s in samples)
    n = len(samples)
    ids = np.zeros((n, max_seq_len), dtype=np.int32)
    special = np.zeros((n, 3 if nsp else 2), dtype=np.int32)
    labels = np.zeros(n, dtype=np.int8)
    for r, s in enumerate(samples):
        row, sp = s.layout(cls_id, sep_id)
        if len(row) > max_seq_len:
            raise ValueError(f"sample of {len(row)} tokens exceeds max_seq_len={max_seq_len}")
        ids[r, :len(row)] = row
        special[r, :len(sp)] = sp
        labels[r] = 1 if s.is_random_next else 0
    return ids, special, labels


def write_samples_to_hdf5(path: str, samples: Sequence[TrainingSample], max_seq_len: int, cls_id: int, sep_id: int) -> None:
    ids, special, labels = samples_to_arrays(samples, max_seq_len, cls_id, sep_id)
    with hdf5.File(path, "w") as f:
        f.create_dataset("input_ids", data=ids, dtype="i4", compression="gzip")
        f.create_dataset("special_token_positions", data=special, dtype="i4", compression="gzip")
        f.create_dataset("next_sentence_labels", data=labels, dtype="i1", compression="gzip")


# -- function-style entry points with the reference's names (utils/encode_data.py:38-181) ----------------------
def convert_to_unicode(text) -> str:
    from .tokenization import convert_to_unicode as _c
    return _c(text)


def get_documents_from_file(input_file: str, tokenizer) -> List[Document]:
    """Blank-line separated documents of tokenised sentences (token *ids*, see the module docstring)."""
    return read_documents(input_file, tokenizer)


def create_samples_from_document(document_idx: int, documents: Sequence[Document], max_seq_len: int,
                                 next_seq_prob: float, short_seq_prob: float,
                                 rng: Optional[random.Random] = None) -> List[TrainingSample]:
    return SamplePacker(max_seq_len, next_seq_prob, short_seq_prob, rng).pack_document(documents, document_idx)


def create_samples(input_file: str, tokenizer, max_seq_len: int, next_seq_prob: float, short_seq_prob: float,
                   rng: Optional[random.Random] = None) -> List[TrainingSample]:
    """All samples of one text file, shuffled."""
    return SamplePacker(max_seq_len, next_seq_prob, short_seq_prob, rng).pack_file(
        get_documents_from_file(input_file, tokenizer))


def output_dir_name(uppercase: bool, max_seq_len: int, nsp: bool) -> str:
    return f"sequences_{'uppercase' if uppercase else 'lowercase'}_max_seq_len_{max_seq_len}_next_seq_task_{str(nsp).lower()}"


def encode_file(input_file: str, output_file: str, vocab_file: str, tokenizer_kind: str, uppercase: bool,
                max_seq_len: int, next_seq_prob: float, short_seq_prob: float, seed: Optional[int] = None) -> int:
    from .tokenization import get_bpe_tokenizer, get_wordpiece_tokenizer
    t0 = time.time()
    tok = (get_wordpiece_tokenizer if tokenizer_kind == "wordpiece" else get_bpe_tokenizer)(vocab_file, uppercase=uppercase)
    cls_id, sep_id = tok.token_to_id("[CLS]"), tok.token_to_id("[SEP]")
    if cls_id is None or sep_id is None:
        raise ValueError("the vocabulary must contain [CLS] and [SEP]")
    print(f"[encoder] Creating instances from {input_file}", flush=True)
    fast = None
    if os.environ.get("B200_NATIVE_TOKENIZER", "1") != "0":      # native C++ encoders (identical ids, tests/test_dataset.py)
        from .tokenization import FastBPE, FastWordPiece
        fast = (FastWordPiece(vocab_file, do_lower_case=not uppercase) if tokenizer_kind == "wordpiece"
                else FastBPE(vocab_file, lowercase=not uppercase))
        if fast.native is None:
            fast = None
    docs = read_documents(input_file, tok, fast=fast)
    packer = SamplePacker(max_seq_len, next_seq_prob, short_seq_prob, random.Random(seed))
    samples = packer.pack_file(docs)
    write_samples_to_hdf5(output_file, samples, max_seq_len, cls_id, sep_id)
    print(f"[encoder] Encoded {output_file} ({len(samples)} samples, time={time.time() - t0:.0f}s)", flush=True)
    return len(samples)
