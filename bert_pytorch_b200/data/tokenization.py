"""Tokenisers.

* ``get_wordpiece_tokenizer`` / ``get_bpe_tokenizer`` -- thin factories over the HuggingFace
  ``tokenizers`` package (src/tokenization.py:42-57; Rust, kept as a dependency because it is
  offline tooling, not a hot path -- SURVEY.md N12).
* ``BasicTokenizer`` / ``WordpieceTokenizer`` / ``BertTokenizer`` -- pure-Python versions
  (src/tokenization.py:60-277); the SQuAD answer post-processing still needs the basic
  tokenizer to align predicted text with the original (run_squad.py:614).
"""
from __future__ import annotations

import collections
import os
import unicodedata
from typing import Dict, Iterable, List, Optional


def get_wordpiece_tokenizer(vocab_file: str, uppercase: bool = False):
    import tokenizers
    return tokenizers.BertWordPieceTokenizer(vocab_file, clean_text=True, handle_chinese_chars=True,
                                             lowercase=not uppercase)


def get_bpe_tokenizer(vocab_file: str, uppercase: bool = False):
    import tokenizers
    # the reference passes only ``vocab`` (src/tokenization.py:51-57), which current `tokenizers` releases turn into an
    # EMPTY model; here the merges file written next to vocab.json by utils/build_vocab.py is loaded as well.  Same
    # options as the reference: add_prefix_space=True, trim_offsets=True.
    merges = vocab_file.replace("vocab.json", "merges.txt") if vocab_file.endswith("vocab.json") else None
    try:
        return tokenizers.ByteLevelBPETokenizer(vocab_file, merges, add_prefix_space=True, lowercase=not uppercase,
                                                trim_offsets=True)
    except TypeError:
        return tokenizers.ByteLevelBPETokenizer(vocab_file, add_prefix_space=True, lowercase=not uppercase, trim_offsets=True)


def load_vocab(vocab_file: str) -> "collections.OrderedDict[str, int]":
    vocab: "collections.OrderedDict[str, int]" = collections.OrderedDict()
    with open(vocab_file, "r", encoding="utf-8") as f:
        for i, line in enumerate(f):
            tok = line.rstrip("\n")
            if tok == "" and i > 0 and not line:
                break
            vocab[tok.strip()] = i
    return vocab


def convert_to_unicode(text) -> str:
    """``str`` passes through, ``bytes`` are decoded as UTF-8 ignoring errors (src/tokenization.py helper)."""
    if isinstance(text, str):
        return text
    if isinstance(text, (bytes, bytearray)):
        return bytes(text).decode("utf-8", "ignore")
    raise ValueError(f"Unsupported string type: {type(text)}")


def whitespace_tokenize(text: str) -> List[str]:
    text = text.strip()
    return text.split() if text else []


def _is_whitespace(ch: str) -> bool:
    return ch in " \t\n\r" or unicodedata.category(ch) == "Zs"


def _is_control(ch: str) -> bool:
    if ch in "\t\n\r":
        return False
    return unicodedata.category(ch).startswith("C")


def _is_punctuation(ch: str) -> bool:
    cp = ord(ch)
    if 33 <= cp <= 47 or 58 <= cp <= 64 or 91 <= cp <= 96 or 123 <= cp <= 126:
        return True
    return unicodedata.category(ch).startswith("P")


def _is_cjk(cp: int) -> bool:
    return (0x4E00 <= cp <= 0x9FFF or 0x3400 <= cp <= 0x4DBF or 0x20000 <= cp <= 0x2A6DF
            or 0x2A700 <= cp <= 0x2B73F or 0x2B740 <= cp <= 0x2B81F or 0x2B820 <= cp <= 0x2CEAF
            or 0xF900 <= cp <= 0xFAFF or 0x2F800 <= cp <= 0x2FA1F)


class BasicTokenizer:
    """Whitespace + punctuation splitting, optional lower-casing/accent stripping, CJK
    characters isolated; tokens in ``never_split`` pass through untouched."""

    def __init__(self, do_lower_case: bool = True,
                 never_split: Iterable[str] = ("[UNK]", "[SEP]", "[PAD]", "[CLS]", "[MASK]")):
        self.do_lower_case = do_lower_case
        self.never_split = set(never_split)

    def tokenize(self, text: str) -> List[str]:
        cleaned = []
        for ch in text:
            cp = ord(ch)
            if cp == 0 or cp == 0xFFFD or _is_control(ch):
                continue
            if _is_whitespace(ch):
                cleaned.append(" ")
            elif _is_cjk(cp):
                cleaned.extend((" ", ch, " "))
            else:
                cleaned.append(ch)
        out: List[str] = []
        for tok in whitespace_tokenize("".join(cleaned)):
            if tok in self.never_split:
                out.append(tok)
                continue
            if self.do_lower_case:
                tok = "".join(c for c in unicodedata.normalize("NFD", tok.lower())
                              if unicodedata.category(c) != "Mn")
            out.extend(self._split_punct(tok))
        return whitespace_tokenize(" ".join(out))

    @staticmethod
    def _split_punct(tok: str) -> List[str]:
        pieces: List[List[str]] = []
        new_word = True
        for ch in tok:
            if _is_punctuation(ch):
                pieces.append([ch])
                new_word = True
            else:
                if new_word:
                    pieces.append([])
                    new_word = False
                pieces[-1].append(ch)
        return ["".join(p) for p in pieces]


class WordpieceTokenizer:
    """Greedy longest-match-first sub-word split with the ``##`` continuation prefix."""

    def __init__(self, vocab: Dict[str, int], unk_token: str = "[UNK]", max_input_chars_per_word: int = 100):
        self.vocab, self.unk_token, self.max_chars = vocab, unk_token, max_input_chars_per_word

    def tokenize(self, text: str) -> List[str]:
        out: List[str] = []
        for word in whitespace_tokenize(text):
            if len(word) > self.max_chars:
                out.append(self.unk_token)
                continue
            start, pieces, bad = 0, [], False
            while start < len(word):
                end, cur = len(word), None
                while start < end:
                    sub = word[start:end]
                    if start > 0:
                        sub = "##" + sub
                    if sub in self.vocab:
                        cur = sub
                        break
                    end -= 1
                if cur is None:
                    bad = True
                    break
                pieces.append(cur)
                start = end
            out.extend([self.unk_token] if bad else pieces)
        return out


class BertTokenizer:
    """End-to-end basic + wordpiece tokenizer (kept for API compatibility)."""

    def __init__(self, vocab_file: str, do_lower_case: bool = True, max_len: Optional[int] = None,
                 never_split=("[UNK]", "[SEP]", "[PAD]", "[CLS]", "[MASK]")):
        self.vocab = load_vocab(vocab_file)
        self.ids_to_tokens = collections.OrderedDict((i, t) for t, i in self.vocab.items())
        self.basic_tokenizer = BasicTokenizer(do_lower_case=do_lower_case, never_split=never_split)
        self.wordpiece_tokenizer = WordpieceTokenizer(self.vocab)
        self.max_len = max_len if max_len is not None else int(1e12)

    def tokenize(self, text: str) -> List[str]:
        return [sub for tok in self.basic_tokenizer.tokenize(text)
                for sub in self.wordpiece_tokenizer.tokenize(tok)]

    def convert_tokens_to_ids(self, tokens: Iterable[str]) -> List[int]:
        ids = [self.vocab[t] for t in tokens]
        if len(ids) > self.max_len:
            raise ValueError(f"sequence length {len(ids)} exceeds the model maximum {self.max_len}")
        return ids

    def convert_ids_to_tokens(self, ids: Iterable[int]) -> List[str]:
        return [self.ids_to_tokens[i] for i in ids]

    def token_to_id(self, token: str) -> Optional[int]:
        return self.vocab.get(token)


def _char_table(do_lower_case: bool, limit: int = 0x2100):
    """This module's character rules for code points < ``limit`` in the form the native tokenizer consumes:
    class per code point (0 normal, 1 whitespace, 2 dropped, 255 not covered) and, for normal characters, the
    replacement characters (lower-cased + accents stripped when ``do_lower_case``) each flagged punctuation or not."""
    import numpy as np
    cls = np.zeros(limit, dtype=np.uint8)
    off = np.zeros(limit + 1, dtype=np.int32)
    blob = bytearray()
    for cp in range(limit):
        ch = chr(cp)
        off[cp] = len(blob)
        if 0xD800 <= cp <= 0xDFFF or _is_cjk(cp):
            cls[cp] = 255
        elif cp == 0 or cp == 0xFFFD or _is_control(ch):
            cls[cp] = 2
        elif _is_whitespace(ch):
            cls[cp] = 1
        else:
            rep = ch
            if do_lower_case:
                rep = "".join(c for c in unicodedata.normalize("NFD", ch.lower()) if unicodedata.category(c) != "Mn")
            rec = bytearray()
            for c in rep:
                b = c.encode("utf-8")
                # a replacement that is itself whitespace / control / outside the table cannot be expressed
                if ord(c) >= limit or _is_whitespace(c) or _is_control(c) or len(b) > 3:
                    rec = None
                    break
                rec.append((0x80 if _is_punctuation(c) else 0) | len(b))
                rec += b
            if rec is None:
                cls[cp] = 255
            else:
                blob += rec
    off[limit] = len(blob)
    return cls, off, np.frombuffer(bytes(blob), dtype=np.uint8)


class FastWordPiece:
    """WordPiece ids for bulk encoding (dataset building): lines whose characters are covered by the character table
    (Latin scripts, general punctuation, ...: everything below U+2100 that this module's rules can express) go through
    the native C++ tokenizer (``ops/csrc/host.cpp``), everything else -- and every line when the helper is not
    built -- through the fallback.  The native path is driven by THIS module's character rules, so it returns exactly
    what ``BasicTokenizer`` + ``WordpieceTokenizer`` return."""

    def __init__(self, vocab_file: str, do_lower_case: bool = True):
        self.py = BertTokenizer(vocab_file, do_lower_case=do_lower_case)
        self.native = None
        try:
            from ..ops import native_host
            host = native_host.load_or_none()
            tokens = list(self.py.vocab.keys())
            # the native table is positional: only usable when ids are exactly 0..n-1 in file order
            if host is not None and all(self.py.vocab[t] == i for i, t in enumerate(tokens)) and \
                    all("\n" not in t for t in tokens):
                self.native = native_host.WordPieceEncoder(host, tokens, _char_table(do_lower_case))
        except Exception:
            self.native = None

    def encode_batch(self, texts: List[str], fallback=None) -> List[List[int]]:
        """``fallback(text) -> ids`` handles the lines the native path does not take; default: the pure-Python
        tokenizer of this module."""
        out: List[Optional[List[int]]] = [None] * len(texts)
        if self.native is not None and texts:
            for i, ids in enumerate(self.native.encode_batch(texts)):
                if ids is not None:
                    out[i] = ids.tolist()
        for i, t in enumerate(texts):
            if out[i] is None:
                out[i] = list(fallback(t)) if fallback is not None else self.py.convert_tokens_to_ids(self.py.tokenize(t))
        return out  # type: ignore[return-value]

    def encode(self, text: str) -> List[int]:
        return self.encode_batch([text])[0]


def _bpe_categories() -> "np.ndarray":
    """Category per BMP code point for the GPT-2 pre-tokenisation pattern: 1 = \\p{L}, 2 = \\p{N}, 3 = \\s, 0 = rest."""
    import numpy as np
    cat = np.zeros(0x10000, dtype=np.uint8)
    for cp in range(0x10000):
        ch = chr(cp)
        c = unicodedata.category(ch)
        if c[0] == "L":
            cat[cp] = 1
        elif c[0] == "N":
            cat[cp] = 2
        elif ch.isspace() and cp not in (0x1C, 0x1D, 0x1E, 0x1F):
            cat[cp] = 3
    return cat


class FastBPE:
    """Byte-level BPE ids for bulk encoding (``--tokenizer bpe`` shards): lines made of BMP characters go through the
    native C++ encoder (``ops/csrc/host.cpp: bpe_*``), the rest -- and everything when the helper is not built --
    through ``fallback`` (the ``tokenizers`` package).  ``vocab_file`` is the ``vocab.json`` written by the vocabulary
    builder, with ``merges.txt`` next to it."""

    def __init__(self, vocab_file: str, lowercase: bool = True, add_prefix_space: bool = True):
        import json
        self.lowercase, self.add_prefix_space = lowercase, add_prefix_space
        self.native = None
        try:
            merges_file = os.path.join(os.path.dirname(vocab_file), "merges.txt")
            with open(vocab_file, "r", encoding="utf-8") as f:
                vocab = json.load(f)
            tokens = sorted(vocab, key=vocab.get)
            if [vocab[t] for t in tokens] != list(range(len(tokens))) or any("\n" in t for t in tokens):
                return
            with open(merges_file, "r", encoding="utf-8") as f:
                merges = [l.rstrip("\n") for l in f if l.strip() and not l.startswith("#version")]
            from ..ops import native_host
            host = native_host.load_or_none()
            if host is not None:
                self.native = native_host.BpeEncoder(host, tokens, merges, _bpe_categories())
        except Exception:
            self.native = None

    def encode_batch(self, texts: List[str], fallback) -> List[List[int]]:
        out: List[Optional[List[int]]] = [None] * len(texts)
        if self.native is not None and texts:
            # lower-casing happens here (str.lower == the normaliser of the `tokenizers` package except for the
            # context dependent capital sigma / dotted capital I: such lines take the fallback)
            take = [i for i, t in enumerate(texts) if not (self.lowercase and ("\u03a3" in t or "\u0130" in t))]
            prepared = [texts[i].lower() if self.lowercase else texts[i] for i in take]
            if self.add_prefix_space:          # ByteLevel(add_prefix_space=True): a space in front unless there is one
                prepared = [t if (not t or t.startswith(" ")) else " " + t for t in prepared]
            for i, ids in zip(take, self.native.encode_batch(prepared)):
                if ids is not None:
                    out[i] = ids.tolist()
        for i, t in enumerate(texts):
            if out[i] is None:
                out[i] = list(fallback(t))
        return out  # type: ignore[return-value]
