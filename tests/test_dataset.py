import warnings

import numpy as np
import pytest
import torch
from hypothesis import given, settings, strategies as st

from bert_pytorch_b200.data import (BatchedPretrainingLoader, DistributedSampler, ShardedPretrainingDataset,
                                    hdf5, mask_batch, segment_ids_and_input_mask, synthetic)
from bert_pytorch_b200.ops import native_host


@settings(max_examples=40, deadline=None)
@given(seed=st.integers(0, 10_000), max_pred=st.integers(1, 40), frac=st.floats(0.05, 0.5), nsp=st.booleans())
def test_masking_invariants(seed, max_pred, frac, nsp):
    rng = np.random.default_rng(seed)
    ids, sp, _ = synthetic.make_samples(16, 64, 1000, nsp, rng)
    out, labels = mask_batch(ids, sp, mask_token_index=4, max_pred_per_seq=max_pred, masked_lm_prob=frac,
                             vocab_size=1000, rng=rng)
    assert out.shape == ids.shape and out is not ids
    changed = out != ids
    for b in range(16):
        last = sp[b, -1]
        specials = set(sp[b].tolist())
        lab_pos = np.nonzero(labels[b] >= 0)[0]
        assert 1 <= len(lab_pos) <= max_pred
        assert all(p < last and p not in specials for p in lab_pos)          # never CLS/SEP/padding
        assert (labels[b, lab_pos] == ids[b, lab_pos]).all()                   # label = original token
        assert not changed[b, [p for p in range(64) if p not in set(lab_pos.tolist())]].any()
        n_cand = last - (sp.shape[1] - 1)
        assert len(lab_pos) <= min(max_pred, max(1, int(n_cand * frac)))


def test_masking_proportions_and_no_mutation():
    rng = np.random.default_rng(0)
    ids, sp, _ = synthetic.make_samples(512, 128, 30522, True, rng)
    keep = ids.copy()
    out, labels = mask_batch(ids, sp, mask_token_index=4, max_pred_per_seq=20, masked_lm_prob=0.15, vocab_size=30522, rng=rng)
    assert np.array_equal(ids, keep)                     # quirk Q3 fixed: the shard cache is never mutated
    sel = labels >= 0
    n = sel.sum()
    frac_mask = (out[sel] == 4).sum() / n
    frac_keep = (out[sel] == ids[sel]).sum() / n
    assert 0.74 < frac_mask < 0.86 and 0.06 < frac_keep < 0.14


def test_native_masking_same_invariants():
    h = native_host.load_or_none()
    if h is None:
        pytest.skip("host helper not built")
    rng = np.random.default_rng(0)
    ids, sp, _ = synthetic.make_samples(256, 128, 30522, True, rng)
    out, labels = native_host.mask_batch(h, ids, sp, seed=5, mask_token_index=4, max_pred_per_seq=20, masked_lm_prob=0.2,
                                         vocab_size=30522, original_token_prob=0.1, random_token_prob=0.1)
    out2, labels2 = native_host.mask_batch(h, ids, sp, seed=5, mask_token_index=4, max_pred_per_seq=20, masked_lm_prob=0.2,
                                           vocab_size=30522, original_token_prob=0.1, random_token_prob=0.1, threads=1)
    assert np.array_equal(out, out2) and np.array_equal(labels, labels2)     # independent of thread count
    sel = labels >= 0
    assert (labels[sel] == ids[sel]).all() and (out[~sel] == ids[~sel]).all()
    cnt = sel.sum(1)
    assert cnt.max() <= 20 and cnt.min() >= 1
    pos = np.arange(128)[None, :]
    assert not (sel & (pos >= sp[:, -1:])).any()
    for j in range(3):
        assert not sel[np.arange(256), sp[:, j]].any()
    assert 0.74 < (out[sel] == 4).mean() < 0.86


def test_segment_and_input_mask():
    ids = np.zeros((2, 10), dtype=np.int32)
    seg, im = segment_ids_and_input_mask(ids, np.array([[0, 3, 7], [0, 2, 9]], dtype=np.int32))
    assert seg[0].tolist() == [0, 0, 0, 0, 1, 1, 1, 1, 0, 0] and im[0].tolist() == [1] * 8 + [0] * 2
    assert seg[1].tolist() == [0, 0, 0] + [1] * 7 and im[1].tolist() == [1] * 10
    seg2, im2 = segment_ids_and_input_mask(ids, np.array([[0, 5], [0, 9]], dtype=np.int32))
    assert seg2.sum() == 0 and im2[0].tolist() == [1] * 6 + [0] * 4


@pytest.fixture
def shards(tmp_path):
    return synthetic.write_shards(str(tmp_path), 4, 25, 32, 500, next_sentence=True, seed=1)


def _dataset(paths):
    return ShardedPretrainingDataset(list(paths), mask_token_index=4, max_pred_per_seq=5, masked_lm_prob=0.2,
                                     vocab_size=500, seed=0)


def test_dataset_index_and_samples(shards):
    ds = _dataset(shards)
    assert len(ds) == 100 and ds.file_idxs == [(0, 25), (25, 50), (50, 75), (75, 100)]
    s = ds[0]
    assert len(s) == 5 and all(a.dtype == np.int64 for a in s) and s[0].shape == (32,)
    ds[24]; ds[25]; ds[49]
    with pytest.raises(RuntimeError):
        ds[99]                                              # sequential-access contract
    with pytest.raises(ValueError):
        ShardedPretrainingDataset(list(shards), 4, 5, 1.5, 500)
    with pytest.raises(ValueError):
        ShardedPretrainingDataset(list(shards), 4, 5, 0.2, 500, shuffle=True)


def test_dataset_skips_bad_files(shards, tmp_path):
    bad = tmp_path / "zz_bad.hdf5"
    bad.write_bytes(b"garbage")
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        ds = _dataset(list(shards) + [str(bad), str(tmp_path / "missing.hdf5")])
    assert len(ds) == 100 and len(w) >= 2
    with pytest.raises(RuntimeError):
        _dataset([str(bad)])


def test_legacy_premasked_schema(tmp_path):
    p = str(tmp_path / "legacy.hdf5")
    N, S = 6, 16
    ids = np.random.default_rng(0).integers(5, 100, size=(N, S), dtype=np.int32)
    pos = np.zeros((N, 4), dtype=np.int32); pos[:, 0] = 3; pos[:, 1] = 7
    mids = np.zeros((N, 4), dtype=np.int32); mids[:, 0] = 11; mids[:, 1] = 12
    with hdf5.File(p, "w") as f:
        for k, v in dict(input_ids=ids, segment_ids=np.zeros_like(ids), input_mask=np.ones_like(ids),
                         masked_lm_positions=pos, masked_lm_ids=mids,
                         next_sentence_labels=np.zeros(N, dtype=np.int8)).items():
            f.create_dataset(k, data=v, compression="gzip")
    ds = ShardedPretrainingDataset([p], None, 4, 0.15, 100)
    s = ds[2]
    assert s[3][3] == 11 and s[3][7] == 12 and (np.delete(s[3], [3, 7]) == -1).all()
    assert np.array_equal(s[0], ids[2])


def test_sampler_contiguous_chunks_and_state(shards):
    ds = _dataset(shards)
    samplers = [DistributedSampler(ds, num_replicas=4, rank=r) for r in range(4)]
    got = [list(s) for s in samplers]
    assert got[0] == list(range(0, 25)) and got[3] == list(range(75, 100))     # contiguous, not round robin
    s = DistributedSampler(ds, num_replicas=3, rank=2)
    assert len(s) == 34 and s.total_size == 102
    idx = list(s)
    assert idx[:3] == [68, 69, 70] and idx[-2:] == [0, 1]                        # padding wraps to the start
    s = DistributedSampler(ds, num_replicas=2, rank=0)
    for _ in range(7):
        next(s)
    sd = s.state_dict()
    assert sd == {"epoch": 0, "seed": 0, "num_replicas": 2, "total_size": 100, "index": 7}
    s2 = DistributedSampler(ds, num_replicas=2, rank=0)
    s2.load_state_dict(sd)
    assert next(s2) == 7
    with pytest.warns(UserWarning):
        DistributedSampler(ds, num_replicas=4, rank=0).load_state_dict(sd)        # world size changed
    with pytest.warns(UserWarning):
        s2.load_state_dict(dict(sd, total_size=999))                              # dataset changed


def test_batched_loader_covers_epoch_and_resumes(shards):
    ds = _dataset(shards)
    sampler = DistributedSampler(ds, num_replicas=1, rank=0)
    loader = BatchedPretrainingLoader(ds, sampler, batch_size=8, pin_memory=False)
    assert len(loader) == 13
    batches = list(loader)
    assert len(batches) == 13 and sum(b[0].size(0) for b in batches) == 100
    assert batches[0][0].dtype == torch.int32 and batches[0][0].shape == (8, 32) and batches[-1][0].size(0) == 4
    # batches that straddle shard boundaries keep sample order
    with hdf5.File(shards[0], "r") as f0, hdf5.File(shards[1], "r") as f1:
        raw = np.concatenate([f0["input_ids"][:], f1["input_ids"][:]])[24:32]
    lab = batches[3][3].numpy()
    assert (np.where(lab >= 0, lab, raw) == raw).all()
    # resume: consume 5 batches, checkpoint, rebuild, continue -> no repeats, no gaps (quirk Q6 fixed)
    ds2 = _dataset(shards)
    sam2 = DistributedSampler(ds2, num_replicas=1, rank=0)
    ld2 = BatchedPretrainingLoader(ds2, sam2, batch_size=8, pin_memory=False, depth=4)
    it = iter(ld2)
    for _ in range(5):
        next(it)
    state = ld2.state_dict()
    assert state["index"] == 40
    ld2.close()
    ds3 = _dataset(shards)
    sam3 = DistributedSampler(ds3, num_replicas=1, rank=0)
    sam3.load_state_dict(state)
    rest = list(BatchedPretrainingLoader(ds3, sam3, batch_size=8, pin_memory=False))
    assert sum(b[0].size(0) for b in rest) == 60


def test_native_wordpiece_matches_python_and_hf(tmp_path):
    """ops/csrc/host.cpp wp_*: the C++ WordPiece gives the ids of the pure-Python BasicTokenizer + WordpieceTokenizer and
    of the `tokenizers` package on ASCII text; non-ASCII lines take the fallback; read_documents keeps document
    boundaries when it tokenises in bulk."""
    import random
    from bert_pytorch_b200.data import encode
    from bert_pytorch_b200.data.tokenization import FastWordPiece, get_wordpiece_tokenizer
    words = ["the", "quick", "brown", "fox", "jump", "##s", "##ed", "over", "lazy", "dog", ".", ",", "!", "?", "'", "-", "a", "b",
             "c", "##a", "##b", "##c", "un", "##able", "##ing", "run", "##n", "1", "2", "##3", "(", ")", "hello", "world", "##ld",
             "wor", "caf", "##e"]
    vocab = ["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"] + words
    vf = tmp_path / "vocab.txt"
    vf.write_text("\n".join(vocab) + "\n")
    fw = FastWordPiece(str(vf), True)
    if fw.native is None:
        pytest.skip("native host helper not built")
    hf = get_wordpiece_tokenizer(str(vf))
    rng = random.Random(3)
    alphabet = "abcABC thequickbrownfoxjumpsedoverlazydogworldhello .,!?'-()123\t"
    texts = ["".join(rng.choice(alphabet) for _ in range(rng.randint(0, 70))) for _ in range(1500)]
    texts += ["[MASK] hello [CLS] WORLD worlds unable running", "", "   ", "a" * 150, "ctrl\x01char\x7f here"]
    nat = fw.encode_batch(texts)
    for t, ids in zip(texts, nat):
        assert ids == fw.py.convert_tokens_to_ids(fw.py.tokenize(t)), t
        if t.strip() and "[" not in t and "\x01" not in t:
            assert ids == hf.encode(t, add_special_tokens=False).ids, t
    # Latin text with accents / typographic punctuation is native too (table driven by the Python rules) ...
    uni = ["Café hello “world” – naïve…", "ÜBER Straße", "ǅ İ x", "\u0301combining", "tab\tsep\u00a0nbsp"]
    uni += ["".join(rng.choice(alphabet + "éèüÜßçñœæ“”–…«»") for _ in range(rng.randint(0, 50))) for _ in range(1500)]
    res = fw.native.encode_batch(uni)
    assert all(r is not None for r in res)
    for t, r in zip(uni, res):
        assert r.tolist() == fw.py.convert_tokens_to_ids(fw.py.tokenize(t)), t
    # ... while characters outside the table (CJK, astral planes) send the line to the fallback
    seen = []
    out = fw.encode_batch(["hello", "hello 中文", "smile \U0001F600"], fallback=lambda t: (seen.append(t), [1])[1])
    assert seen == ["hello 中文", "smile \U0001F600"] and out[1] == [1] and out[0] == [vocab.index("hello")]
    # bulk tokenisation inside read_documents: same documents as the line-by-line path
    src = tmp_path / "corpus.txt"
    src.write_text("the quick brown fox.\njumps over\n\n\nhello world!\ncafé hello\n中文 hello\n\nthe dog\n", encoding="utf-8")
    a = encode.read_documents(str(src), hf)
    b = encode.read_documents(str(src), hf, fast=fw, batch_lines=3)
    assert a == b and len(a) == 3


def test_loader_smoke_main(tmp_path, capsys):
    """``python -m bert_pytorch_b200.data.dataset`` (the reference's loader smoke loop, src/dataset.py:431-505)."""
    import numpy as np
    from bert_pytorch_b200.data import dataset as D, synthetic
    synthetic.write_shards(str(tmp_path), 2, 24, 32, 1000, True, seed=0)
    assert D._smoke(["--input_dir", str(tmp_path), "--batch_size", "8", "--epochs", "2", "--max_predictions_per_seq", "8",
                     "--vocab_size", "1000"]) == 0
    out = capsys.readouterr().out
    assert "Dataset size = 48" in out and "epoch 1: 48 samples" in out


def test_native_bpe_matches_tokenizers_package(tmp_path):
    """The C++ byte-level BPE (GPT-2 pre-tokenisation pattern, byte alphabet, rank-ordered merges) returns the ids of
    ``tokenizers.ByteLevelBPETokenizer`` on mixed-script BMP text, cased and lower-cased; characters outside the BMP and the
    context-dependent lower-casing cases take the fallback."""
    import random
    tokenizers = pytest.importorskip("tokenizers")
    from bert_pytorch_b200.data.tokenization import FastBPE, get_bpe_tokenizer
    rnd = random.Random(0)
    words = ("the quick brown fox jumps over lazy dog while rain keeps falling don't it's we'll they've I'm he'd 1234 56.7 "
             "café naïve über straße").split()
    corpus = tmp_path / "c.txt"
    corpus.write_text("\n".join(" ".join(rnd.choice(words) for _ in range(rnd.randint(3, 14))) for _ in range(300)), encoding="utf-8")
    tok = tokenizers.ByteLevelBPETokenizer(add_prefix_space=False, lowercase=True)
    tok.train([str(corpus)], vocab_size=380, show_progress=False, special_tokens=["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"])
    tok.save_model(str(tmp_path))
    vocab = str(tmp_path / "vocab.json")
    C = lambda *cps: "".join(chr(c) for c in cps)
    pools = ["abcdefghijklmnopqrstuvwxyz", "ABCDEFGHIJKLMNOPQRSTUVWXYZ", "0123456789", C(0x20, 0x09, 0x0A, 0x0D, 0xA0, 0x2003, 0x3000, 0x85, 0x2028),
             ".,;:!?'\"()[]{}-_/\\@#$%^&*+=<>|~`", C(0xE9, 0xFC, 0xF1, 0xDF, 0x142, 0x15F, 0x131), C(0x4E2D, 0x6587, 0x6771, 0x4EAC),
             C(0x3042, 0x30A2), C(0xAC00), C(0x430, 0x414), C(0x3B1, 0x394), C(0x627, 0x644), C(0x00, 0x07, 0xFFFD, 0x200B, 0xAD, 0xFEFF, 0x1C),
             C(0x2014, 0x2018, 0x201D, 0x2026, 0x20AC), C(0x301, 0x308), C(0xB2, 0xBC, 0x663, 0x2167), "'s 't 're 've 'm 'll 'd 'S 'LL",
             C(0x3A3, 0x130, 0x1F600)]
    texts = ["".join(rnd.choice(rnd.choice(pools)) for _ in range(rnd.randint(1, 40))) for _ in range(800)]
    texts += ["Don't stop, it's 12:30pm!!  ok\n\n next\tline ", "  leading", "trailing   ", "a  b   c", ""]
    for lower in (True, False):
        hf = get_bpe_tokenizer(vocab, uppercase=not lower)
        fast = FastBPE(vocab, lowercase=lower)
        if fast.native is None:
            pytest.skip("native host helper not built")
        fell_back = []
        got = fast.encode_batch(texts, fallback=lambda t: fell_back.append(t) or hf.encode(t).ids)
        assert 0 < len(fell_back) < 0.8 * len(texts)                  # both paths are exercised (the last pool forces fallbacks)
        for t, ids in zip(texts, got):
            assert list(ids) == hf.encode(t).ids, (lower, ascii(t))
