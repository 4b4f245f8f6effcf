"""Resolve a model location to a local path (src/file_utils.py:1-263): local files/dirs are
returned as is, http(s) URLs are downloaded once into a cache keyed by sha256(url[+etag]),
``s3://`` needs boto3 (optional).  There is no network in the build image, so the remote
branches are only reachable on user machines."""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import tempfile
from pathlib import Path
from typing import Optional, Tuple
from urllib.parse import urlparse

PYTORCH_PRETRAINED_BERT_CACHE = Path(os.getenv("PYTORCH_PRETRAINED_BERT_CACHE",
                                               Path.home() / ".pytorch_pretrained_bert"))


def url_to_filename(url: str, etag: Optional[str] = None) -> str:
    name = hashlib.sha256(url.encode("utf-8")).hexdigest()
    if etag:
        name += "." + hashlib.sha256(etag.encode("utf-8")).hexdigest()
    return name


def filename_to_url(filename: str, cache_dir=None) -> Tuple[str, Optional[str]]:
    cache_dir = Path(cache_dir or PYTORCH_PRETRAINED_BERT_CACHE)
    path = cache_dir / filename
    meta = Path(str(path) + ".json")
    if not path.exists() or not meta.exists():
        raise FileNotFoundError(f"{path} (or its .json metadata) not found")
    with open(meta, encoding="utf-8") as f:
        m = json.load(f)
    return m["url"], m.get("etag")


def _http_head_etag(url: str) -> Optional[str]:
    import urllib.request
    req = urllib.request.Request(url, method="HEAD")
    with urllib.request.urlopen(req, timeout=30) as r:
        if r.status != 200:
            raise IOError(f"HEAD request failed for {url} with status {r.status}")
        return r.headers.get("ETag")


def _http_get(url: str, fh) -> None:
    import urllib.request
    with urllib.request.urlopen(url, timeout=60) as r:
        shutil.copyfileobj(r, fh)


def _s3_split(url: str) -> Tuple[str, str]:
    p = urlparse(url)
    if not p.netloc or not p.path:
        raise ValueError(f"bad s3 path {url}")
    return p.netloc, p.path.lstrip("/")


def get_from_cache(url: str, cache_dir=None) -> str:
    cache_dir = Path(cache_dir or PYTORCH_PRETRAINED_BERT_CACHE)
    cache_dir.mkdir(parents=True, exist_ok=True)
    if url.startswith("s3://"):
        try:
            import boto3  # type: ignore
        except ImportError as e:
            raise ImportError("s3:// locations need boto3") from e
        bucket, key = _s3_split(url)
        obj = boto3.resource("s3").Object(bucket, key)
        etag = obj.e_tag
        fetch = lambda fh: obj.download_fileobj(fh)  # noqa: E731
    else:
        etag = _http_head_etag(url)
        fetch = lambda fh: _http_get(url, fh)  # noqa: E731
    path = cache_dir / url_to_filename(url, etag)
    if not path.exists():
        with tempfile.NamedTemporaryFile() as tmp:
            fetch(tmp)
            tmp.flush(); tmp.seek(0)
            with open(path, "wb") as out:
                shutil.copyfileobj(tmp, out)
        with open(str(path) + ".json", "w", encoding="utf-8") as m:
            json.dump({"url": url, "etag": etag}, m)
    return str(path)


def cached_path(url_or_filename, cache_dir=None) -> str:
    s = os.fspath(url_or_filename)
    scheme = urlparse(s).scheme
    if scheme in ("http", "https", "s3"):
        return get_from_cache(s, cache_dir)
    if os.path.exists(s):
        return s
    if scheme == "":
        raise FileNotFoundError(f"file {s} not found")
    raise ValueError(f"unable to parse {s} as a URL or as a local path")


# -- remaining public helpers of the reference module (src/file_utils.py:127-263) ------------------------------
def split_s3_path(url: str) -> Tuple[str, str]:
    """``s3://bucket/key`` -> ``(bucket, key)``."""
    return _s3_split(url)


def s3_request(func):
    """Decorator turning botocore's 404 ``ClientError`` into ``FileNotFoundError`` (src/file_utils.py:140-157)."""
    import functools

    @functools.wraps(func)
    def wrapper(url, *args, **kwargs):
        try:
            return func(url, *args, **kwargs)
        except Exception as e:                      # botocore may be absent: match on the response payload only
            code = getattr(e, "response", {}).get("Error", {}).get("Code") if hasattr(e, "response") else None
            if code is not None and int(code) == 404:
                raise FileNotFoundError(f"file {url} not found") from e
            raise
    return wrapper


@s3_request
def s3_etag(url: str) -> Optional[str]:
    import boto3  # type: ignore
    bucket, key = split_s3_path(url)
    return boto3.resource("s3").Object(bucket, key).e_tag


@s3_request
def s3_get(url: str, temp_file) -> None:
    import boto3  # type: ignore
    bucket, key = split_s3_path(url)
    boto3.resource("s3").Bucket(bucket).download_fileobj(key, temp_file)


def http_get(url: str, temp_file) -> None:
    """Stream ``url`` into the open binary file ``temp_file``."""
    _http_get(url, temp_file)


def read_set_from_file(filename: str) -> set:
    """One item per line -> set of the stripped lines."""
    with open(filename, "r", encoding="utf-8") as f:
        return {line.rstrip() for line in f}


def get_file_extension(path: str, dot: bool = True, lower: bool = True) -> str:
    ext = os.path.splitext(path)[1]
    ext = ext if dot else ext[1:]
    return ext.lower() if lower else ext
