import os
import struct

import numpy as np
import pytest

from bert_pytorch_b200.data import hdf5, synthetic


@pytest.mark.parametrize("compression", [None, "gzip"])
@pytest.mark.parametrize("dtype", ["i4", "i1", "u2", "i8", "f4", "f8"])
def test_roundtrip_dtypes(tmp_path, compression, dtype):
    rng = np.random.default_rng(0)
    a = (rng.standard_normal((257, 33)) * 50).astype(dtype)
    b = (rng.standard_normal(1000) * 100).astype(dtype)
    p = str(tmp_path / "x.hdf5")
    with hdf5.File(p, "w") as f:
        f.create_dataset("a", data=a, compression=compression)
        f.create_dataset("b", data=b, compression=compression, chunks=(100,) if compression else None)
    with hdf5.File(p, "r") as f:
        assert sorted(f.keys()) == ["a", "b"]
        assert f["a"].shape == (257, 33) and f["a"].dtype == np.dtype(dtype) and len(f["b"]) == 1000
        assert np.array_equal(f["a"][:], a) and np.array_equal(f["b"][:], b)
        assert np.array_equal(f["a"][10:20], a[10:20])
        assert f["a"].compression == ("gzip" if compression else None)
        assert "a" in f and "zzz" not in f


def test_pretraining_shard_schema(tmp_path):
    paths = synthetic.write_shards(str(tmp_path), 2, 50, 128, 30522, next_sentence=True, seed=3)
    with hdf5.File(paths[0], "r") as f:
        assert set(f.keys()) == {"input_ids", "special_token_positions", "next_sentence_labels"}
        ids, sp, nsl = f["input_ids"][:], f["special_token_positions"][:], f["next_sentence_labels"][:]
    assert ids.dtype == np.int32 and ids.shape == (50, 128)
    assert sp.dtype == np.int32 and sp.shape == (50, 3)
    assert nsl.dtype == np.int8 and nsl.shape == (50,)
    assert (ids[np.arange(50), sp[:, 0]] == 2).all() and (ids[np.arange(50), sp[:, 2]] == 3).all()


def test_shuffle_filter_and_many_chunks(tmp_path):
    a = np.arange(64 * 7, dtype=np.int32).reshape(64, 7)
    p = str(tmp_path / "s.hdf5")
    with hdf5.File(p, "w") as f:
        f.create_dataset("a", data=a, compression="gzip", shuffle=True, chunks=(1, 7))   # 64 chunks = node capacity
    with hdf5.File(p, "r") as f:
        assert np.array_equal(f["a"][:], a)
    with pytest.raises(hdf5.HDF5Error):
        with hdf5.File(str(tmp_path / "t.hdf5"), "w") as f:
            f.create_dataset("a", data=np.zeros((65, 2), np.int32), compression="gzip", chunks=(1, 2))


def test_native_and_python_decoders_agree(tmp_path, monkeypatch):
    a = np.random.default_rng(1).integers(0, 30000, size=(999, 128), dtype=np.int32)
    p = str(tmp_path / "n.hdf5")
    with hdf5.File(p, "w") as f:
        f.create_dataset("a", data=a, compression="gzip", chunks=(100, 128))
    with hdf5.File(p, "r") as f:
        fast = f["a"][:]
    monkeypatch.setattr(hdf5, "_NATIVE", None)
    monkeypatch.setattr(hdf5, "_NATIVE_TRIED", True)
    with hdf5.File(p, "r") as f:
        slow = f["a"][:]
    assert np.array_equal(fast, a) and np.array_equal(slow, a)


def test_superblock_fields_and_not_hdf5(tmp_path):
    p = str(tmp_path / "x.hdf5")
    with hdf5.File(p, "w") as f:
        f.create_dataset("a", data=np.arange(10, dtype=np.int32))
    raw = open(p, "rb").read()
    assert raw[:8] == b"\x89HDF\r\n\x1a\n" and raw[8] == 0 and raw[13] == 8 and raw[14] == 8
    eof = struct.unpack_from("<Q", raw, 40)[0]
    assert eof == len(raw)
    bad = tmp_path / "bad.hdf5"
    bad.write_bytes(b"not an hdf5 file at all" * 10)
    with pytest.raises(hdf5.HDF5Error):
        hdf5.File(str(bad), "r")


def _v2_file_with_compact_links(path, arr):
    """Hand-built 'new style' file: superblock v2, v2 object headers (OHDR), link messages, a contiguous
    dataset with a v2 dataspace -- the structures h5py emits with libver='latest'."""
    img = bytearray(b"\0" * 48)
    def put(b):
        off = len(img); img.extend(b); img.extend(b"\0" * ((-len(img)) % 8)); return off
    data_addr = put(arr.tobytes())
    def ohdr(msgs):
        body = b"".join(struct.pack("<BHB", t, len(d), 0) + d for t, d in msgs)
        return b"OHDR" + bytes([2, 0x00]) + struct.pack("<B", len(body)) + body + b"\0\0\0\0"
    dspace = bytes([2, arr.ndim, 0, 1]) + b"".join(struct.pack("<Q", s) for s in arr.shape)
    dtype = bytes([0x10, 0x08, 0, 0]) + struct.pack("<I", 4) + struct.pack("<HH", 0, 32)
    layout = bytes([3, 1]) + struct.pack("<QQ", data_addr, arr.nbytes)
    ds_addr = put(ohdr([(1, dspace), (3, dtype), (8, layout)]))
    name = b"input_ids"
    link = bytes([1, 0x00, len(name)]) + name + struct.pack("<Q", ds_addr)
    linkinfo = bytes([0, 0]) + struct.pack("<QQ", 0xFFFFFFFFFFFFFFFF, 0xFFFFFFFFFFFFFFFF)
    root = put(ohdr([(2, linkinfo), (6, link)]))
    sb = b"\x89HDF\r\n\x1a\n" + bytes([2, 8, 8, 0]) + struct.pack("<QQQQ", 0, 0xFFFFFFFFFFFFFFFF, len(img), root) + b"\0\0\0\0"
    img[:len(sb)] = sb
    open(path, "wb").write(bytes(img))


def test_reader_handles_v2_superblock_ohdr_and_link_messages(tmp_path):
    arr = np.arange(60, dtype=np.int32).reshape(12, 5)
    p = str(tmp_path / "v2.hdf5")
    _v2_file_with_compact_links(p, arr)
    with hdf5.File(p, "r") as f:
        assert f.keys() == ["input_ids"]
        assert np.array_equal(f["input_ids"][:], arr)
