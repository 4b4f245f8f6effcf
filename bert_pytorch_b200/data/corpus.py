"""Offline corpus tooling (library behind the ``utils/*.py`` CLIs): download, format to one sentence per line,
shard by size, sample + shard, build a vocabulary.

Parity targets: utils/download.py (datasets: wikicorpus, bookscorpus, squad, sst-2, mrpc, google weights),
utils/format.py (wikiextractor ``<doc>`` parsing, BooksCorpus one-book-per-file, sentence splitting,
round-robin files -> shards, blank line between articles), utils/shard.py (split at article boundaries once
a shard exceeds a byte budget; sizes accept K/M/B suffixes), utils/sample_and_shard.py (random articles up
to a sentence budget per input file), utils/build_vocab.py (HF tokenizers trainer, special tokens first,
``[PAD]`` forced to index 0).  Sentence splitting uses nltk's punkt when it is installed and a built-in
rule-based splitter otherwise (nltk is not in the image).
"""
from __future__ import annotations

import bz2
import hashlib
import os
import random
import re
import subprocess
import time
import urllib.request
from pathlib import Path
from typing import Iterable, Iterator, List, Optional, Sequence

# ---------------------------------------------------------------------------
# helpers
# ---------------------------------------------------------------------------
_SUFFIX = {"K": 10 ** 3, "M": 10 ** 6, "B": 10 ** 9, "G": 10 ** 9}


def parse_value_as_int(value) -> int:
    """'100M' -> 100000000, '2.5K' -> 2500, 42 -> 42."""
    if isinstance(value, (int, float)):
        return int(value)
    v = str(value).strip()
    if re.fullmatch(r"\d+", v):
        return int(v)
    if len(v) > 1 and v[-1].upper() in _SUFFIX:
        return int(float(v[:-1]) * _SUFFIX[v[-1].upper()])
    raise ValueError(f'Unable to parse "{value}" as integer')


def find_txt_files(path: str) -> List[str]:
    if os.path.isfile(path):
        return [path]
    if os.path.isdir(path):
        return sorted(str(p) for p in Path(path).rglob("*.txt") if p.is_file())
    raise ValueError(f"{path} is not a valid path")


_ABBREV = {"mr", "mrs", "ms", "dr", "prof", "sr", "jr", "st", "vs", "etc", "e.g", "i.e", "inc", "ltd", "co", "u.s",
           "no", "fig", "gen", "col", "lt", "sgt", "capt", "jan", "feb", "mar", "apr", "jun", "jul", "aug", "sep",
           "sept", "oct", "nov", "dec", "mt", "rev", "hon", "pres", "gov", "sen", "rep", "approx", "est", "vol"}
_SPLIT = re.compile(r"(?<=[.!?])[\"')\]]*\s+(?=[\"'(\[]*[A-Z0-9])")


def split_sentences(text: str) -> List[str]:
    try:
        from nltk.tokenize import sent_tokenize  # type: ignore
        return sent_tokenize(text)
    except Exception:  # noqa: BLE001 - nltk or its punkt model is missing: rule based fallback
        pass
    parts = _SPLIT.split(text.strip())
    out: List[str] = []
    for part in parts:
        if out:
            last = out[-1].rstrip("\"')]").split()[-1].rstrip(".").lower() if out[-1].split() else ""
            if out[-1].rstrip("\"')]").endswith(".") and (last in _ABBREV or (len(last) == 1 and last.isalpha())):
                out[-1] = out[-1] + " " + part       # "Dr. Smith", "J. R. R. Tolkien"
                continue
        if part:
            out.append(part)
    return [s.strip() for s in out if s.strip()]


# ---------------------------------------------------------------------------
# formatting
# ---------------------------------------------------------------------------
def wiki_articles(path: str) -> Iterator[List[str]]:
    """Articles of a wikiextractor output file as lists of raw lines (title line dropped)."""
    lines: List[str] = []
    inside = False
    with open(path, "r", encoding="utf-8", errors="ignore") as f:
        for line in f:
            if line.startswith("<doc id="):
                inside, lines = True, []
            elif line.startswith("</doc>"):
                if inside and len(lines) > 1:
                    yield lines[1:]
                inside = False
            elif inside:
                lines.append(line)


def book_article(path: str) -> List[str]:
    with open(path, "r", encoding="ISO-8859-1") as f:
        text = " ".join(l.encode("utf-8", "ignore").decode("utf-8").strip() for l in f)
    return [text] if text.strip() else []


def format_files(dataset: str, input_files: Sequence[str], output_file: str) -> None:
    t0 = time.time()
    with open(output_file, "w", encoding="utf-8") as out:
        for path in input_files:
            articles: Iterable[List[str]] = wiki_articles(path) if dataset == "wikicorpus" else [book_article(path)]
            for lines in articles:
                wrote = False
                for raw in lines:
                    for s in split_sentences(raw.strip()):
                        out.write(s.strip() + "\n")
                        wrote = True
                if wrote:
                    out.write("\n")
    print(f"[{dataset}] Finished shard: {output_file} (time={time.time() - t0:.1f}s)", flush=True)


def format_corpus(dataset: str, input_dir: str, output_dir: str, processes: int = 1, shards: int = 64) -> List[str]:
    """Round-robin the input files over ``shards`` output files ``{dataset}_{i}.txt``."""
    if dataset == "wikicorpus":
        files = sorted(str(p) for p in Path(input_dir).rglob("wiki_*") if p.is_file())
    elif dataset == "bookscorpus":
        files = sorted(str(p) for p in Path(input_dir).rglob("*.txt") if p.is_file())
    else:
        raise ValueError(f"unknown dataset {dataset}")
    if not files:
        raise ValueError(f"no input files found under {input_dir}")
    os.makedirs(output_dir, exist_ok=True)
    shards = len(files) if shards <= 0 else min(shards, len(files))
    jobs = [(dataset, files[i::shards], os.path.join(output_dir, f"{dataset}_{i}.txt")) for i in range(shards)]
    if processes > 1:
        import multiprocessing as mp
        with mp.Pool(processes) as pool:
            pool.starmap(format_files, jobs)
    else:
        for j in jobs:
            format_files(*j)
    return [j[2] for j in jobs]


# ---------------------------------------------------------------------------
# sharding / sampling
# ---------------------------------------------------------------------------
def shard_text(input_file: str, output_format: str, bytes_per_shard: int, max_shards: Optional[int] = None) -> int:
    """Split at article boundaries (blank lines); returns the number of shards written."""
    if not os.path.exists(input_file):
        raise ValueError(f"Could not find input file {input_file}")
    if "{index}" not in output_format:
        raise ValueError('output_file_format must contain "{index}"')
    os.makedirs(os.path.dirname(os.path.abspath(output_format)), exist_ok=True)
    index, written = 1, 0
    out = open(output_format.format(index=index), "w", encoding="utf-8")
    with open(input_file, "r", encoding="utf-8") as f:
        for line in f:
            out.write(line)
            written += len(line.encode("utf-8"))
            if line == "\n" and written > bytes_per_shard:
                out.close()
                index += 1
                written = 0
                if max_shards is not None and index > max_shards:
                    return index - 1
                out = open(output_format.format(index=index), "w", encoding="utf-8")
    out.close()
    return index


def file_to_articles(path: str) -> List[List[str]]:
    arts: List[List[str]] = [[]]
    with open(path, "r", encoding="utf-8") as f:
        for line in f:
            line = line.rstrip()
            if line == "":
                if arts[-1]:
                    arts.append([])
            else:
                arts[-1].append(line)
    return [a for a in arts if a]


def sample_and_shard(input_files: Sequence[str], output_format: str, shard_size: int, sentences: int,
                     rng: Optional[random.Random] = None) -> int:
    rng = rng or random.Random()
    per_input = sentences // max(len(input_files), 1)
    os.makedirs(os.path.dirname(os.path.abspath(output_format)), exist_ok=True)
    idx, written = 0, 0
    out = open(output_format.format(index=idx), "w", encoding="utf-8")
    for n, path in enumerate(input_files):
        t0 = time.time()
        arts = file_to_articles(path)
        order = list(range(len(arts)))
        rng.shuffle(order)
        count = 0
        for a in reversed(order):          # the reference pops from the end of the shuffled list
            if count >= per_input:
                break
            if written > shard_size:
                out.close()
                idx += 1
                written = 0
                out = open(output_format.format(index=idx), "w", encoding="utf-8")
            for line in arts[a]:
                out.write(line + "\n")
                written += len(line.encode("utf-8")) + 1
            out.write("\n")
            count += len(arts[a])
        print(f"[sampler] Finished sampling from input file {n + 1}/{len(input_files)} (time={time.time() - t0:.1f})")
    out.close()
    return idx + 1


# ---------------------------------------------------------------------------
# vocabulary
# ---------------------------------------------------------------------------
def build_vocab(input_files: Sequence[str], output: str, size: int = 30000, tokenizer: str = "wordpiece",
                uppercase: bool = False, special_tokens: Sequence[str] = ("[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"),
                pad_token: str = "[PAD]") -> List[str]:
    import tokenizers
    if tokenizer == "wordpiece":
        tok = tokenizers.BertWordPieceTokenizer(clean_text=True, handle_chinese_chars=True, lowercase=not uppercase)
    elif tokenizer == "bpe":
        tok = tokenizers.ByteLevelBPETokenizer(add_prefix_space=True, lowercase=not uppercase, trim_offsets=True)
    else:
        raise ValueError(f"unknown tokenizer {tokenizer}")
    tok.train(list(input_files), vocab_size=size, show_progress=False, special_tokens=list(special_tokens))
    vocab = [w for w, _ in sorted(tok.get_vocab().items(), key=lambda kv: kv[1])]
    for t in special_tokens:                       # specials to the front ...
        if t in vocab:
            vocab.insert(0, vocab.pop(vocab.index(t)))
    if pad_token in vocab:                         # ... and the padding token at index 0
        vocab.remove(pad_token)
    vocab.insert(0, pad_token)
    os.makedirs(os.path.dirname(os.path.abspath(output)) or ".", exist_ok=True)
    with open(output, "w", encoding="utf-8") as f:
        for w in vocab:
            f.write(w + "\n")
    if tokenizer == "bpe":
        # a byte-level BPE model is vocab.json + merges.txt (the word list alone cannot tokenise): written next to the
        # list; pass .../vocab.json as ``vocab_file`` to the encoder / runners
        tok.save_model(os.path.dirname(os.path.abspath(output)))
    return vocab


# ---------------------------------------------------------------------------
# downloads (need network; kept for feature parity with utils/download.py)
# ---------------------------------------------------------------------------
SQUAD_URLS = {
    "https://rajpurkar.github.io/SQuAD-explorer/dataset/train-v1.1.json": "v1.1/train-v1.1.json",
    "https://rajpurkar.github.io/SQuAD-explorer/dataset/dev-v1.1.json": "v1.1/dev-v1.1.json",
    "https://worksheets.codalab.org/rest/bundles/0xbcd57bee090b421c982906709c8c27e1/contents/blob/": "v1.1/evaluate-v1.1.py",
    "https://rajpurkar.github.io/SQuAD-explorer/dataset/train-v2.0.json": "v2.0/train-v2.0.json",
    "https://rajpurkar.github.io/SQuAD-explorer/dataset/dev-v2.0.json": "v2.0/dev-v2.0.json",
    "https://worksheets.codalab.org/rest/bundles/0x6b567e1cf2e041ec80d7098f031c5c9e/contents/blob/": "v2.0/evaluate-v2.0.py",
}
WIKI_URLS = {"https://dumps.wikimedia.org/enwiki/latest/enwiki-latest-pages-articles.xml.bz2": "wikicorpus_en.xml.bz2"}
WEIGHT_URLS = {
    "bert_base_uncased": "https://storage.googleapis.com/bert_models/2018_10_18/uncased_L-12_H-768_A-12.zip",
    "bert_large_uncased": "https://storage.googleapis.com/bert_models/2018_10_18/uncased_L-24_H-1024_A-16.zip",
    "bert_base_cased": "https://storage.googleapis.com/bert_models/2018_10_18/cased_L-12_H-768_A-12.zip",
    "bert_large_cased": "https://storage.googleapis.com/bert_models/2018_10_18/cased_L-24_H-1024_A-16.zip",
}
GLUE_TASKS = {"sst-2": "SST", "mrpc": "MRPC"}
# SHA-256 of the files inside Google's 2018_10_18 archives, in the order (bert_config.json, ckpt.data, ckpt.index,
# ckpt.meta, vocab.txt): integrity check of the download and a tripwire for upstream changes
# (reference: utils/download.py:136-175)
_WEIGHT_FILES = ("bert_config.json", "bert_model.ckpt.data-00000-of-00001", "bert_model.ckpt.index",
                 "bert_model.ckpt.meta", "vocab.txt")
_UNCASED_VOCAB = "07eced375cec144d27c900241f3e339478dec958f92fddbc551f295c992038a3"
_CASED_VOCAB = "eeaa9875b23b04b4c54ef759d03db9d1ba1554838f8fb26c5d96fa551df93d02"
WEIGHT_SHA256 = {
    "bert_base_uncased": ("7b4e5f53efbd058c67cda0aacfafb340113ea1b5797d9ce6ee411704ba21fcbc",
                          "58580dc5e0bf0ae0d2efd51d0e8272b2f808857f0a43a88aaf7549da6d7a8a84",
                          "04c1323086e2f1c5b7c0759d8d3e484afbb0ab45f51793daab9f647113a0117b",
                          "dd5682170a10c3ea0280c2e9b9a45fee894eb62da649bbdea37b38b0ded5f60e", _UNCASED_VOCAB),
    "bert_large_uncased": ("bfa42236d269e2aeb3a6d30412a33d15dbe8ea597e2b01dc9518c63cc6efafcb",
                           "bc6b3363e3be458c99ecf64b7f472d2b7c67534fd8f564c0556a678f90f4eea1",
                           "68b52f2205ffc64dc627d1120cf399c1ef1cbc35ea5021d1afc889ffe2ce2093",
                           "6fcce8ff7628f229a885a593625e3d5ff9687542d5ef128d9beb1b0c05edc4a1", _UNCASED_VOCAB),
    "bert_base_cased": ("f11dfb757bea16339a33e1bf327b0aade6e57fd9c29dc6b84f7ddb20682f48bc",
                        "734d5a1b68bf98d4e9cb6b6692725d00842a1937af73902e51776905d8f760ea",
                        "517d6ef5c41fc2ca1f595276d6fccf5521810d57f5a74e32616151557790f7b1",
                        "5f8a9771ff25dadd61582abb4e3a748215a10a6b55947cbb66d0f0ba1694be98", _CASED_VOCAB),
    "bert_large_cased": ("7adb2125c8225da495656c982fd1c5f64ba8f20ad020838571a3f8a954c2df57",
                         "6ff33640f40d472f7a16af0c17b1179ca9dcc0373155fb05335b6a4dd1657ef0",
                         "ef42a53f577fbe07381f4161b13c7cab4f4fc3b167cec6a9ae382c53d18049cf",
                         "d2ddff3ed33b80091eac95171e94149736ea74eb645e575d942ec4a5e01a40a1", _CASED_VOCAB),
}


def verify_weights(name: str, model_dir: str) -> List[str]:
    """Names of the files under ``model_dir`` whose SHA-256 differs from the recorded one (missing files count)."""
    bad = []
    for fname, want in zip(_WEIGHT_FILES, WEIGHT_SHA256[name]):
        path = os.path.join(model_dir, fname)
        if not os.path.isfile(path) or sha256sum(path) != want:
            bad.append(fname)
    return bad


def sha256sum(path: str) -> str:
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for block in iter(lambda: f.read(1 << 20), b""):
            h.update(block)
    return h.hexdigest()


def fetch(url: str, dst: str) -> str:
    if os.path.isfile(dst):
        print(f"** Download file already exists, skipping download: {dst}")
        return dst
    os.makedirs(os.path.dirname(os.path.abspath(dst)), exist_ok=True)
    print(f"Downloading {url} -> {dst}")
    tmp = dst + ".part"
    with urllib.request.urlopen(url, timeout=60) as r, open(tmp, "wb") as f:
        while True:
            block = r.read(1 << 20)
            if not block:
                break
            f.write(block)
    os.replace(tmp, dst)
    return dst


def download(dataset: str, root: str) -> None:
    dataset = dataset.lower()
    base = os.path.join(root, {"sst-2": "glue", "mrpc": "glue", "mprc": "glue"}.get(dataset, dataset))
    os.makedirs(base, exist_ok=True)
    if dataset == "squad":
        for url, rel in SQUAD_URLS.items():
            fetch(url, os.path.join(base, rel))
    elif dataset == "wikicorpus":
        for url, rel in WIKI_URLS.items():
            cfile = fetch(url, os.path.join(base, rel))
            xml = cfile.rsplit(".", 1)[0]
            if os.path.isfile(xml):
                print(f"[wikicorpus] ** Extracted file already exists, skipping extraction: {xml}")
            else:
                with bz2.open(cfile, "rb") as src, open(xml, "wb") as dst:
                    for block in iter(lambda: src.read(1 << 22), b""):
                        dst.write(block)
    elif dataset == "bookscorpus":
        repo = os.path.join(base, "bookcorpus")
        if not os.path.isdir(repo):
            subprocess.run(["git", "clone", "https://github.com/soskek/bookcorpus.git", repo], check=True)
        subprocess.run(["python", os.path.join(repo, "download_files.py"), "--list", os.path.join(repo, "url_list.jsonl"),
                        "--out", os.path.join(base, "download"), "--trash-bad-count"], check=True)
    elif dataset in ("sst-2", "mrpc", "mprc"):
        script = fetch("https://gist.githubusercontent.com/W4ngatang/60c2bdb54d156a41194446737ce03e2e/raw/"
                       "17b8dd0d724281ed7c3b2aeeda662b92809aadd5/download_glue_data.py",
                       os.path.join(base, "download_glue_data.py"))
        subprocess.run(["python", script, "--data_dir", base, "--tasks", GLUE_TASKS.get(dataset, "MRPC")], check=True)
    elif dataset == "weights":
        import zipfile
        for name, url in WEIGHT_URLS.items():
            z = fetch(url, os.path.join(base, os.path.basename(url)))
            with zipfile.ZipFile(z) as zf:
                zf.extractall(base)
            bad = verify_weights(name, z[:-4])
            for fname in bad:
                print(f"[weights] SHA256sum does not match on file: {fname} from download url: {url}")
            if not bad:
                print(f"[weights] {name}: {len(_WEIGHT_FILES)} files verified")
    else:
        raise ValueError(f"unknown dataset {dataset}")
