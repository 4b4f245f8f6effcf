"""Loader for the host-side native helper ``_host.so`` (ops/csrc/host.cpp): multi-threaded
HDF5 chunk inflate + vectorised dynamic masking.  Pure C++ (no CUDA), built in-tree by
:mod:`bert_pytorch_b200.ops.build`.  Everything that uses it has a NumPy fallback, so a
missing build only costs speed on the CPU side."""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import numpy as np

_LIB = None
_TRIED = False
_HERE = os.path.dirname(os.path.abspath(__file__))


def lib_path() -> str:
    return os.path.join(_HERE, "_host.so")


class _Host:
    def __init__(self, lib: ctypes.CDLL):
        self.lib = lib
        i64p = ctypes.POINTER(ctypes.c_int64)
        lib.h5_inflate_rows.restype = ctypes.c_int
        lib.h5_inflate_rows.argtypes = [ctypes.c_void_p, ctypes.c_int64, i64p, i64p, i64p, ctypes.c_int64,
                                        ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p,
                                        ctypes.c_int]
        lib.mask_batch_i32.restype = None
        lib.mask_batch_i32.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int32,
                                       ctypes.c_int32, ctypes.c_double, ctypes.c_int32, ctypes.c_double,
                                       ctypes.c_double, ctypes.c_uint64, ctypes.c_int]

        lib.wp_create.restype = ctypes.c_void_p
        lib.wp_create.argtypes = [ctypes.c_char_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                  ctypes.c_void_p, ctypes.c_int64]
        lib.wp_destroy.restype = None
        lib.wp_destroy.argtypes = [ctypes.c_void_p]
        lib.wp_encode_batch.restype = ctypes.c_int64
        lib.wp_encode_batch.argtypes = [ctypes.c_void_p, ctypes.c_char_p, i64p, ctypes.c_int64, ctypes.c_void_p,
                                        ctypes.c_int64, i64p, ctypes.c_void_p, ctypes.c_int]

    def inflate_rows(self, raw: bytes, offs: np.ndarray, sizes: np.ndarray, rows0: np.ndarray,
                     chunk_rows: int, total_rows: int, row_bytes: int, out: np.ndarray,
                     threads: int = 0) -> None:
        if not isinstance(raw, bytes):
            raw = bytes(raw)
        src = ctypes.cast(ctypes.c_char_p(raw), ctypes.c_void_p)   # no copy
        i64p = ctypes.POINTER(ctypes.c_int64)
        rc = self.lib.h5_inflate_rows(src, len(raw), offs.ctypes.data_as(i64p), sizes.ctypes.data_as(i64p),
                                      rows0.ctypes.data_as(i64p), len(offs), chunk_rows, total_rows,
                                      row_bytes, out.ctypes.data_as(ctypes.c_void_p), threads)
        if rc != 0:
            raise IOError(f"native HDF5 chunk inflate failed (zlib rc={rc})")


def load():
    global _LIB, _TRIED
    if _LIB is None:
        if not os.path.exists(lib_path()):
            raise ImportError(f"{lib_path()} is not built (python -m bert_pytorch_b200.ops.build)")
        _LIB = _Host(ctypes.CDLL(lib_path()))
    return _LIB


def load_or_none() -> Optional[_Host]:
    global _TRIED
    if _LIB is not None:
        return _LIB
    if _TRIED:
        return None
    _TRIED = True
    try:
        return load()
    except Exception:
        return None


def mask_batch(host: _Host, ids: np.ndarray, sp: np.ndarray, *, seed: int, mask_token_index: int,
               max_pred_per_seq: int, masked_lm_prob: float, vocab_size: int,
               original_token_prob: float, random_token_prob: float, threads: int = 0):
    ids = np.ascontiguousarray(ids, dtype=np.int32)
    sp = np.ascontiguousarray(sp, dtype=np.int32)
    out_ids = np.empty_like(ids)
    labels = np.empty_like(ids)
    B, S = ids.shape
    host.lib.mask_batch_i32(ids.ctypes.data, sp.ctypes.data, out_ids.ctypes.data, labels.ctypes.data,
                            B, S, sp.shape[1], mask_token_index, max_pred_per_seq, masked_lm_prob,
                            vocab_size, original_token_prob, random_token_prob, seed, threads)
    return out_ids, labels


class WordPieceEncoder:
    """Native WordPiece (ops/csrc/host.cpp: wp_*): clean-up + whitespace / punctuation splitting + greedy
    longest-match sub-words, batched and threaded.  ``vocab_tokens``: tokens in id order.  ``char_table``:
    ``(cls uint8[N], map_off int32[N+1], map_blob uint8[...])`` -- the caller's character rules for code points
    below N (see ``data/tokenization.py: FastWordPiece``)."""

    def __init__(self, host: _Host, vocab_tokens, char_table):
        self._host = host
        blob = "\n".join(vocab_tokens).encode("utf-8")
        cls, off, mp = (np.ascontiguousarray(char_table[0], dtype=np.uint8),
                        np.ascontiguousarray(char_table[1], dtype=np.int32),
                        np.ascontiguousarray(char_table[2], dtype=np.uint8))
        assert off.size == cls.size + 1
        self._h = host.lib.wp_create(blob, len(blob), cls.ctypes.data, off.ctypes.data, cls.size, mp.ctypes.data, mp.size)
        if not self._h:
            raise RuntimeError("wp_create failed")

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                self._host.lib.wp_destroy(h)
            except Exception:
                pass

    def encode_batch(self, texts, threads: int = 0):
        """-> list with an int32 id array per text, or ``None`` where the text is outside the character table."""
        n = len(texts)
        if n == 0:
            return []
        enc = [t.encode("utf-8", errors="surrogatepass") for t in texts]
        offs = np.zeros(n + 1, dtype=np.int64)
        np.cumsum([len(e) for e in enc], out=offs[1:])
        buf = b"".join(enc)
        out_offs = np.empty(n + 1, dtype=np.int64)
        ok = np.empty(n, dtype=np.uint8)
        cap = max(16, 2 * len(buf))
        i64p = ctypes.POINTER(ctypes.c_int64)
        for _ in range(2):
            out = np.empty(cap, dtype=np.int32)
            total = self._host.lib.wp_encode_batch(self._h, buf, offs.ctypes.data_as(i64p), n, out.ctypes.data, cap,
                                                   out_offs.ctypes.data_as(i64p), ok.ctypes.data, threads)
            if total <= cap:
                break
            cap = int(total)
        return [out[out_offs[i]:out_offs[i + 1]] if ok[i] else None for i in range(n)]


class BpeEncoder:
    """Native byte-level BPE (ops/csrc/host.cpp: bpe_*): GPT-2 pre-tokenisation over a category table for the BMP,
    byte -> printable alphabet, rank-ordered merges; batched and threaded.  ``vocab_tokens`` in id order, ``merges`` as
    ``"left right"`` lines in rank order, ``categories`` uint8[65536] (0 other, 1 letter, 2 number, 3 white space)."""

    def __init__(self, host: _Host, vocab_tokens, merges, categories):
        self._host = host
        vb = "\n".join(vocab_tokens).encode("utf-8")
        mb = "\n".join(merges).encode("utf-8")
        cat = np.ascontiguousarray(categories, dtype=np.uint8)
        lib = host.lib
        lib.bpe_create.restype = ctypes.c_void_p
        lib.bpe_create.argtypes = [ctypes.c_char_p, ctypes.c_int64, ctypes.c_char_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64]
        lib.bpe_destroy.argtypes = [ctypes.c_void_p]
        lib.bpe_encode_batch.restype = ctypes.c_int64
        lib.bpe_encode_batch.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_int64), ctypes.c_int64,
                                         ctypes.c_void_p, ctypes.c_int64, ctypes.POINTER(ctypes.c_int64), ctypes.c_void_p,
                                         ctypes.c_int]
        self._h = lib.bpe_create(vb, len(vb), mb, len(mb), cat.ctypes.data, cat.size)
        if not self._h:
            raise RuntimeError("bpe_create failed")

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                self._host.lib.bpe_destroy(h)
            except Exception:
                pass

    def encode_batch(self, texts, threads: int = 0):
        """-> list with an int32 id array per text, or ``None`` where the text is not covered (non-BMP characters,
        a symbol missing from the vocabulary)."""
        n = len(texts)
        if n == 0:
            return []
        try:
            enc = [t.encode("utf-8") for t in texts]
        except UnicodeEncodeError:
            enc = [t.encode("utf-8", errors="replace") for t in texts]
        offs = np.zeros(n + 1, dtype=np.int64)
        np.cumsum([len(e) for e in enc], out=offs[1:])
        buf = b"".join(enc)
        out_offs = np.empty(n + 1, dtype=np.int64)
        ok = np.empty(n, dtype=np.uint8)
        cap = max(16, len(buf) + 16)
        i64p = ctypes.POINTER(ctypes.c_int64)
        for _ in range(2):
            out = np.empty(cap, dtype=np.int32)
            total = self._host.lib.bpe_encode_batch(self._h, buf, offs.ctypes.data_as(i64p), n, out.ctypes.data, cap,
                                                    out_offs.ctypes.data_as(i64p), ok.ctypes.data, threads)
            if total <= cap:
                break
            cap = int(total)
        return [out[out_offs[i]:out_offs[i + 1]] if ok[i] else None for i in range(n)]
