"""A self-contained HDF5 reader/writer (no libhdf5, no h5py -- neither exists in the
image, SURVEY.md 0.2 / hard part #1).

Scope: what the pre-training shard schema needs (written by
utils/encode_data.py:204-210 and read by src/dataset.py:217-222 of the reference):
a root group holding N-dimensional fixed-point / IEEE-float datasets, stored
contiguous, compact, or chunked with the deflate (+ shuffle, fletcher32) filters.

Reader: superblock v0/v1/v2/v3; object headers v1 and v2 (with continuation
blocks); "old style" groups (symbol table: v1 B-tree + SNOD + local heap) and "new
style" groups with compact link messages; data layout v3 (compact / contiguous /
chunked via v1 B-tree) and v4 (single-chunk, implicit and fixed-array chunk
indices); nested groups via ``file["a/b"]``.
Writer: superblock v0, symbol-table root group, v1 object headers, contiguous or
chunked+gzip datasets (one leaf B-tree node, <= 64 chunks) -- the same on-disk
structures h5py's default ``libver='earliest'`` produces, so files are meant to be
readable by stock HDF5 tools.

The API mirrors the h5py subset the reference touches: ``File(path, 'r'|'w')`` as a
context manager, ``f.keys()``, ``f[name][:]``, ``f[name].shape/.dtype/len()``,
``f.create_dataset(name, data=, dtype=, compression='gzip')``.
The bulk chunk decode runs in the native helper ``ops/csrc/host.cpp`` when built
(multi-threaded inflate + scatter); the pure-Python path below is the fallback and
the oracle for its tests.
"""
from __future__ import annotations

import os
import struct
import zlib
from typing import Any, Dict, Iterator, List, Optional, Sequence, Tuple, Union

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class HDF5Error(IOError):
    pass


# ---------------------------------------------------------------------------
# low level reader helpers
# ---------------------------------------------------------------------------


class _Buf:
    """Random access byte source with little-endian field readers."""

    def __init__(self, data: Union[bytes, memoryview], base: int = 0):
        self.d = memoryview(data)
        self.base = base

    def u(self, off: int, n: int) -> int:
        return int.from_bytes(self.d[off:off + n], "little")

    def bytes(self, off: int, n: int) -> bytes:
        return bytes(self.d[off:off + n])


def _pad8(n: int) -> int:
    return (n + 7) & ~7


class _Datatype:
    def __init__(self, np_dtype: np.dtype, size: int):
        self.np_dtype, self.size = np_dtype, size


def _parse_datatype(b: bytes) -> _Datatype:
    cls_ver = b[0]
    cls = cls_ver & 0x0F
    bits0 = b[1]
    size = struct.unpack_from("<I", b, 4)[0]
    order = ">" if (bits0 & 1) else "<"
    if cls == 0:  # fixed point
        signed = bool(bits0 & 0x08)
        if size not in (1, 2, 4, 8):
            raise HDF5Error(f"unsupported integer size {size}")
        return _Datatype(np.dtype(f"{order}{'i' if signed else 'u'}{size}"), size)
    if cls == 1:  # floating point
        if size not in (2, 4, 8):
            raise HDF5Error(f"unsupported float size {size}")
        return _Datatype(np.dtype(f"{order}f{size}"), size)
    raise HDF5Error(f"unsupported HDF5 datatype class {cls}")


def _parse_dataspace(b: bytes, L: int) -> Tuple[int, ...]:
    ver, rank, flags = b[0], b[1], b[2]
    if ver == 1:
        off = 8
    elif ver == 2:
        off = 4
        if b[3] == 2:  # null dataspace
            return (0,)
    else:
        raise HDF5Error(f"unsupported dataspace version {ver}")
    return tuple(int.from_bytes(b[off + i * L: off + (i + 1) * L], "little") for i in range(rank))


def _parse_filters(b: bytes) -> List[Tuple[int, List[int]]]:
    ver, n = b[0], b[1]
    out: List[Tuple[int, List[int]]] = []
    off = 8 if ver == 1 else 2
    for _ in range(n):
        fid = struct.unpack_from("<H", b, off)[0]
        off += 2
        if ver == 1 or fid >= 256:
            name_len = struct.unpack_from("<H", b, off)[0]
            off += 2
        else:
            name_len = 0
        _flags, nvals = struct.unpack_from("<HH", b, off)
        off += 4
        off += _pad8(name_len) if ver == 1 else name_len
        vals = list(struct.unpack_from(f"<{nvals}I", b, off))
        off += 4 * nvals
        if ver == 1 and nvals % 2:
            off += 4
        out.append((fid, vals))
    return out


def _unshuffle(buf: bytes, elem: int) -> bytes:
    if elem <= 1:
        return buf
    n = len(buf) // elem
    a = np.frombuffer(buf, dtype=np.uint8, count=n * elem).reshape(elem, n)
    return a.T.tobytes() + buf[n * elem:]


def _shuffle(buf: bytes, elem: int) -> bytes:
    if elem <= 1:
        return buf
    n = len(buf) // elem
    a = np.frombuffer(buf, dtype=np.uint8, count=n * elem).reshape(n, elem)
    return a.T.tobytes() + buf[n * elem:]


def _apply_filters_decode(raw: bytes, filters, mask: int, elem: int) -> bytes:
    for i in range(len(filters) - 1, -1, -1):
        if mask & (1 << i):
            continue
        fid, _vals = filters[i]
        if fid == 1:
            raw = zlib.decompress(raw)
        elif fid == 2:
            raw = _unshuffle(raw, elem)
        elif fid == 3:
            raw = raw[:-4]
        else:
            raise HDF5Error(f"unsupported HDF5 filter id {fid}")
    return raw


# ---------------------------------------------------------------------------
# object header parsing
# ---------------------------------------------------------------------------

MSG_DATASPACE, MSG_LINKINFO, MSG_DATATYPE, MSG_FILL_OLD, MSG_FILL, MSG_LINK = 1, 2, 3, 4, 5, 6
MSG_LAYOUT, MSG_FILTERS, MSG_CONT, MSG_SYMTAB = 8, 0x0B, 0x10, 0x11


class _Reader:
    def __init__(self, path: str):
        self.path = path
        with open(path, "rb") as f:
            self.raw = f.read()
        self.buf = _Buf(self.raw)
        if self.raw[:8] != SIGNATURE:
            # superblock may sit at 512, 1024, ... (user block)
            off, found = 512, False
            while off < len(self.raw):
                if self.raw[off:off + 8] == SIGNATURE:
                    found = True
                    break
                off *= 2
            if not found:
                raise HDF5Error(f"{path}: not an HDF5 file")
            self.sb = off
        else:
            self.sb = 0
        self._parse_superblock()

    def _parse_superblock(self) -> None:
        b, s = self.buf, self.sb
        ver = b.u(s + 8, 1)
        self.sb_version = ver
        if ver in (0, 1):
            self.O, self.L = b.u(s + 13, 1), b.u(s + 14, 1)
            p = s + 24 + (4 if ver == 1 else 0)
            self.base = b.u(p, self.O)
            p += 4 * self.O
            # root symbol table entry
            self.root_header = b.u(p + self.O, self.O)
            cache_type = b.u(p + 2 * self.O, 4)
            self.root_scratch = None
            if cache_type == 1:
                sp = p + 2 * self.O + 8
                self.root_scratch = (b.u(sp, self.O), b.u(sp + self.O, self.O))
        elif ver in (2, 3):
            self.O, self.L = b.u(s + 9, 1), b.u(s + 10, 1)
            p = s + 12
            self.base = b.u(p, self.O)
            self.root_header = b.u(p + 3 * self.O, self.O)
            self.root_scratch = None
        else:
            raise HDF5Error(f"unsupported superblock version {ver}")
        if self.base == UNDEF:
            self.base = 0
        self.base += 0  # addresses are relative to base (user block aware)

    def addr(self, a: int) -> int:
        return a + self.base

    # -- messages ---------------------------------------------------------
    def messages(self, header_addr: int) -> List[Tuple[int, bytes]]:
        b = self.buf
        a = self.addr(header_addr)
        out: List[Tuple[int, bytes]] = []
        if b.bytes(a, 4) == b"OHDR":
            flags = b.u(a + 5, 1)
            p = a + 6
            if flags & 0x20:
                p += 16
            if flags & 0x10:
                p += 4
            szlen = 1 << (flags & 3)
            chunk0 = b.u(p, szlen)
            p += szlen
            blocks = [(p, chunk0)]
            track_order = bool(flags & 0x04)
            while blocks:
                start, size = blocks.pop(0)
                q, end = start, start + size
                while q + 4 <= end:
                    mtype = b.u(q, 1)
                    msize = b.u(q + 1, 2)
                    q += 4 + (2 if track_order else 0)
                    if q + msize > end:
                        break
                    data = b.bytes(q, msize)
                    q += msize
                    if mtype == MSG_CONT:
                        ca, cl = int.from_bytes(data[:self.O], "little"), int.from_bytes(
                            data[self.O:self.O + self.L], "little")
                        ca = self.addr(ca)
                        if b.bytes(ca, 4) != b"OCHK":
                            raise HDF5Error("bad object header continuation")
                        blocks.append((ca + 4, cl - 8))
                    elif mtype != 0:
                        out.append((mtype, data))
            return out
        # version 1
        if b.u(a, 1) != 1:
            raise HDF5Error(f"unsupported object header at {a:#x}")
        nmsgs = b.u(a + 2, 2)
        hsize = b.u(a + 8, 4)
        blocks = [(a + 16, hsize)]
        seen = 0
        while blocks and seen < nmsgs:
            start, size = blocks.pop(0)
            q, end = start, start + size
            while q + 8 <= end and seen < nmsgs:
                mtype, msize = b.u(q, 2), b.u(q + 2, 2)
                data = b.bytes(q + 8, msize)
                q += 8 + msize
                seen += 1
                if mtype == MSG_CONT:
                    ca = int.from_bytes(data[:self.O], "little")
                    cl = int.from_bytes(data[self.O:self.O + self.L], "little")
                    blocks.append((self.addr(ca), cl))
                elif mtype != 0:
                    out.append((mtype, data))
        return out

    # -- groups -------------------------------------------------------------
    def group_links(self, header_addr: int) -> Dict[str, int]:
        links: Dict[str, int] = {}
        for mtype, data in self.messages(header_addr):
            if mtype == MSG_SYMTAB:
                btree = int.from_bytes(data[:self.O], "little")
                heap = int.from_bytes(data[self.O:2 * self.O], "little")
                links.update(self._walk_group_btree(btree, heap))
            elif mtype == MSG_LINK:
                name, target = self._parse_link(data)
                if target is not None:
                    links[name] = target
            elif mtype == MSG_LINKINFO:
                flags = data[1]
                p = 2 + (8 if flags & 1 else 0)
                fheap = int.from_bytes(data[p:p + self.O], "little")
                if fheap != UNDEF and fheap != (1 << (8 * self.O)) - 1:
                    raise HDF5Error("dense link storage (fractal heap) is not supported; "
                                    "the shard schema never has more than 8 links per group")
        return links

    def _parse_link(self, d: bytes) -> Tuple[str, Optional[int]]:
        flags = d[1]
        p = 2
        ltype = 0
        if flags & 0x08:
            ltype = d[p]; p += 1
        if flags & 0x04:
            p += 8
        if flags & 0x10:
            p += 1
        n = 1 << (flags & 3)
        ln = int.from_bytes(d[p:p + n], "little"); p += n
        name = d[p:p + ln].decode("utf-8"); p += ln
        if ltype != 0:
            return name, None  # soft/external links are ignored
        return name, int.from_bytes(d[p:p + self.O], "little")

    def _heap_data_addr(self, heap_addr: int) -> int:
        a = self.addr(heap_addr)
        if self.buf.bytes(a, 4) != b"HEAP":
            raise HDF5Error("bad local heap signature")
        return self.addr(self.buf.u(a + 8 + 2 * self.L, self.O))

    def _cstr(self, a: int) -> str:
        end = self.raw.index(b"\0", a)
        return self.raw[a:end].decode("utf-8")

    def _walk_group_btree(self, btree_addr: int, heap_addr: int) -> Dict[str, int]:
        heap_data = self._heap_data_addr(heap_addr)
        out: Dict[str, int] = {}
        stack = [btree_addr]
        b = self.buf
        while stack:
            a = self.addr(stack.pop())
            sig = b.bytes(a, 4)
            if sig == b"TREE":
                level, n = b.u(a + 5, 1), b.u(a + 6, 2)
                p = a + 8 + 2 * self.O
                for i in range(n):
                    p += self.L  # key i
                    stack.append(b.u(p, self.O))
                    p += self.O
                _ = level
            elif sig == b"SNOD":
                n = b.u(a + 6, 2)
                p = a + 8
                for _i in range(n):
                    name_off = b.u(p, self.O)
                    hdr = b.u(p + self.O, self.O)
                    out[self._cstr(heap_data + name_off)] = hdr
                    p += 2 * self.O + 24
            else:
                raise HDF5Error(f"unexpected group node signature {sig!r}")
        return out

    # -- datasets -----------------------------------------------------------
    def dataset_info(self, header_addr: int) -> Dict[str, Any]:
        info: Dict[str, Any] = {"filters": []}
        for mtype, data in self.messages(header_addr):
            if mtype == MSG_DATASPACE:
                info["shape"] = _parse_dataspace(data, self.L)
                if data[0] == 1 and data[1] == 0:
                    info["shape"] = ()
                elif data[0] == 2 and data[1] == 0:
                    info["shape"] = ()
            elif mtype == MSG_DATATYPE:
                info["dtype"] = _parse_datatype(data)
            elif mtype == MSG_FILTERS:
                info["filters"] = _parse_filters(data)
            elif mtype == MSG_LAYOUT:
                info["layout"] = data
        return info

    def is_dataset(self, header_addr: int) -> bool:
        types = {t for t, _ in self.messages(header_addr)}
        return MSG_LAYOUT in types and MSG_DATATYPE in types

    def read_dataset(self, info: Dict[str, Any]) -> np.ndarray:
        shape: Tuple[int, ...] = info["shape"]
        dt: _Datatype = info["dtype"]
        lay: bytes = info["layout"]
        n_elem = int(np.prod(shape)) if shape else 1
        ver, cls = lay[0], lay[1]
        O, L = self.O, self.L
        if ver == 3 or ver == 4:
            if cls == 0:  # compact
                sz = struct.unpack_from("<H", lay, 2)[0]
                arr = np.frombuffer(lay[4:4 + sz], dtype=dt.np_dtype, count=n_elem)
                return arr.reshape(shape).astype(dt.np_dtype.newbyteorder("="))
            if cls == 1:  # contiguous
                a = int.from_bytes(lay[2:2 + O], "little")
                if a == UNDEF or n_elem == 0:
                    return np.zeros(shape, dtype=dt.np_dtype.newbyteorder("="))
                arr = np.frombuffer(self.raw, dtype=dt.np_dtype, count=n_elem, offset=self.addr(a))
                return arr.reshape(shape).astype(dt.np_dtype.newbyteorder("="))
            if cls == 2:
                return self._read_chunked(info, lay, ver)
        elif ver in (1, 2):
            rank = lay[1]
            lcls = lay[2]
            p = 8
            a = None
            if lcls != 0:
                a = int.from_bytes(lay[p:p + O], "little"); p += O
            dims = [struct.unpack_from("<I", lay, p + 4 * i)[0] for i in range(rank)]
            if lcls == 1:
                arr = np.frombuffer(self.raw, dtype=dt.np_dtype, count=n_elem, offset=self.addr(a))
                return arr.reshape(shape).astype(dt.np_dtype.newbyteorder("="))
            if lcls == 2:
                return self._assemble(info, self._chunks_btree_v1(a, rank), tuple(dims[:-1]))
        raise HDF5Error(f"unsupported data layout version {ver} class {cls}")

    def _read_chunked(self, info, lay: bytes, ver: int) -> np.ndarray:
        O, L = self.O, self.L
        dt: _Datatype = info["dtype"]
        shape = info["shape"]
        if ver == 3:
            ndim = lay[2]
            a = int.from_bytes(lay[3:3 + O], "little")
            dims = [struct.unpack_from("<I", lay, 3 + O + 4 * i)[0] for i in range(ndim)]
            chunk = tuple(dims[:-1])
            if a == UNDEF:
                return np.zeros(shape, dtype=dt.np_dtype.newbyteorder("="))
            return self._assemble(info, self._chunks_btree_v1(a, ndim), chunk)
        # version 4
        flags, ndim, enc = lay[2], lay[3], lay[4]
        p = 5
        dims = [int.from_bytes(lay[p + i * enc: p + (i + 1) * enc], "little") for i in range(ndim)]
        p += ndim * enc
        chunk = tuple(dims[:-1])
        idx_type = lay[p]; p += 1
        filtered = bool(info["filters"])
        chunk_bytes = int(np.prod(chunk)) * dt.size
        grid = [-(-s // c) for s, c in zip(shape, chunk)]
        n_chunks = int(np.prod(grid))
        if idx_type == 1:  # single chunk
            size, mask = chunk_bytes, 0
            if flags & 0x02:
                size = int.from_bytes(lay[p:p + L], "little"); p += L
                mask = struct.unpack_from("<I", lay, p)[0]; p += 4
            a = int.from_bytes(lay[p:p + O], "little")
            return self._assemble(info, [((0,) * len(chunk), a, size, mask)], chunk)
        if idx_type == 2:  # implicit: chunks laid out back to back, unfiltered
            a = int.from_bytes(lay[p:p + O], "little")
            ents = []
            for i in range(n_chunks):
                ents.append((self._grid_to_offset(i, grid, chunk), a + i * chunk_bytes, chunk_bytes, 0))
            return self._assemble(info, ents, chunk)
        if idx_type == 3:  # fixed array
            _page_bits = lay[p]; p += 1
            a = int.from_bytes(lay[p:p + O], "little")
            return self._assemble(info, self._chunks_fixed_array(a, grid, chunk, chunk_bytes, filtered), chunk)
        raise HDF5Error(f"chunk index type {idx_type} (extensible array / v2 B-tree) is not supported; "
                        "rewrite the file with libver='earliest' or a fixed-size dataset")

    @staticmethod
    def _grid_to_offset(i: int, grid: Sequence[int], chunk: Sequence[int]) -> Tuple[int, ...]:
        idx = []
        for g in reversed(grid):
            idx.append(i % g)
            i //= g
        return tuple(k * c for k, c in zip(reversed(idx), chunk))

    def _chunks_fixed_array(self, addr: int, grid, chunk, chunk_bytes: int, filtered: bool):
        b = self.buf
        a = self.addr(addr)
        if b.bytes(a, 4) != b"FAHD":
            raise HDF5Error("bad fixed array header")
        entry_size = b.u(a + 6, 1)
        page_bits = b.u(a + 7, 1)
        n = b.u(a + 8, self.L)
        dblk = self.addr(b.u(a + 8 + self.L, self.O))
        if b.bytes(dblk, 4) != b"FADB":
            raise HDF5Error("bad fixed array data block")
        p = dblk + 6 + self.O
        page_elems = 1 << page_bits
        paged = n > page_elems
        if paged:
            npages = -(-n // page_elems)
            p += -(-npages // 8)
            p += 4  # data block checksum precedes the pages
        ents = []
        for i in range(n):
            if paged and i > 0 and i % page_elems == 0:
                p += 4  # page checksum
            if filtered:
                ca = b.u(p, self.O)
                szlen = entry_size - self.O - 4
                size = b.u(p + self.O, szlen)
                mask = b.u(p + self.O + szlen, 4)
            else:
                ca, size, mask = b.u(p, self.O), chunk_bytes, 0
            p += entry_size
            if ca != UNDEF:
                ents.append((self._grid_to_offset(i, grid, chunk), ca, size, mask))
        return ents

    def _chunks_btree_v1(self, root: int, ndim: int):
        b = self.buf
        out = []
        stack = [root]
        key_size = 8 + 8 * ndim
        while stack:
            a = self.addr(stack.pop())
            if b.bytes(a, 4) != b"TREE":
                raise HDF5Error("bad chunk B-tree node")
            level, n = b.u(a + 5, 1), b.u(a + 6, 2)
            p = a + 8 + 2 * self.O
            for _ in range(n):
                size, mask = b.u(p, 4), b.u(p + 4, 4)
                offs = tuple(b.u(p + 8 + 8 * i, 8) for i in range(ndim - 1))
                child = b.u(p + key_size, self.O)
                p += key_size + self.O
                if level == 0:
                    out.append((offs, child, size, mask))
                else:
                    stack.append(child)
        return out

    def _assemble(self, info, entries, chunk: Tuple[int, ...]) -> np.ndarray:
        dt: _Datatype = info["dtype"]
        shape = info["shape"]
        native = dt.np_dtype.newbyteorder("=")
        out = np.zeros(shape, dtype=native)
        filters = info["filters"]
        fast = _native_chunks()
        if (fast is not None and filters and all(f[0] == 1 for f in filters)
                and dt.np_dtype.byteorder in ("<", "=", "|") and len(shape) in (1, 2)
                and (len(shape) == 1 or chunk[1] == shape[1])):
            # rows-only chunking + pure deflate: hand the whole job to the native decoder
            ok = all(mask == 0 for (_o, _a, _s, mask) in entries)
            if ok:
                row_bytes = dt.size * (shape[1] if len(shape) == 2 else 1)
                offs = np.asarray([self.addr(a) for (_o, a, _s, _m) in entries], dtype=np.int64)
                sizes = np.asarray([s for (_o, _a, s, _m) in entries], dtype=np.int64)
                rows0 = np.asarray([o[0] for (o, _a, _s, _m) in entries], dtype=np.int64)
                fast.inflate_rows(self.raw, offs, sizes, rows0, int(chunk[0]), int(shape[0]),
                                  int(row_bytes), out)
                return out
        for offs, a, size, mask in entries:
            raw = self.raw[self.addr(a): self.addr(a) + size]
            if filters:
                raw = _apply_filters_decode(raw, filters, mask, dt.size)
            block = np.frombuffer(raw, dtype=dt.np_dtype, count=int(np.prod(chunk))).reshape(chunk)
            sl_out, sl_in = [], []
            for o, c, s in zip(offs, chunk, shape):
                hi = min(o + c, s)
                sl_out.append(slice(o, hi))
                sl_in.append(slice(0, hi - o))
            out[tuple(sl_out)] = block[tuple(sl_in)]
        return out


_NATIVE = None
_NATIVE_TRIED = False


def _native_chunks():
    """The C++ chunk decoder (ops/csrc/h5chunks.cpp) if it has been built."""
    global _NATIVE, _NATIVE_TRIED
    if not _NATIVE_TRIED:
        _NATIVE_TRIED = True
        try:
            from ..ops import native_host
            _NATIVE = native_host.load()
        except Exception:
            _NATIVE = None
    return _NATIVE


# ---------------------------------------------------------------------------
# public read API
# ---------------------------------------------------------------------------


class Dataset:
    def __init__(self, reader: _Reader, name: str, header_addr: int):
        self._r, self.name, self._h = reader, name, header_addr
        self._info = reader.dataset_info(header_addr)
        if "shape" not in self._info or "dtype" not in self._info or "layout" not in self._info:
            raise HDF5Error(f"{name}: incomplete dataset header")

    @property
    def shape(self) -> Tuple[int, ...]:
        return tuple(self._info["shape"])

    @property
    def dtype(self) -> np.dtype:
        return self._info["dtype"].np_dtype.newbyteorder("=")

    @property
    def compression(self) -> Optional[str]:
        return "gzip" if any(f[0] == 1 for f in self._info["filters"]) else None

    def __len__(self) -> int:
        if not self.shape:
            raise TypeError("scalar dataset has no len()")
        return self.shape[0]

    def read(self) -> np.ndarray:
        return self._r.read_dataset(self._info)

    def __getitem__(self, key) -> np.ndarray:
        arr = self.read()
        if key is Ellipsis or (isinstance(key, slice) and key == slice(None)) or key == ():
            return arr
        return arr[key]

    def __array__(self, dtype=None, copy=None):
        a = self.read()
        return a.astype(dtype) if dtype is not None else a


class Group:
    def __init__(self, reader: _Reader, name: str, header_addr: int):
        self._r, self.name, self._h = reader, name, header_addr
        self._links = reader.group_links(header_addr)

    def keys(self):
        return list(self._links.keys())

    def __iter__(self) -> Iterator[str]:
        return iter(self._links)

    def __len__(self) -> int:
        return len(self._links)

    def __contains__(self, name: str) -> bool:
        try:
            self[name]
            return True
        except KeyError:
            return False

    def __getitem__(self, name: str):
        node: Any = self
        for part in [p for p in name.split("/") if p]:
            if not isinstance(node, Group) or part not in node._links:
                raise KeyError(name)
            h = node._links[part]
            node = (Dataset(self._r, part, h) if self._r.is_dataset(h) else Group(self._r, part, h))
        return node

    def items(self):
        return [(k, self[k]) for k in self._links]


# ---------------------------------------------------------------------------
# writer
# ---------------------------------------------------------------------------


def _dtype_message(dt: np.dtype) -> bytes:
    dt = np.dtype(dt)
    if dt.byteorder == ">":
        raise HDF5Error("big-endian output is not supported")
    if dt.kind in "iu":
        bits0 = 0x08 if dt.kind == "i" else 0x00
        return bytes([0x10 | 0, bits0, 0, 0]) + struct.pack("<I", dt.itemsize) + struct.pack(
            "<HH", 0, dt.itemsize * 8)
    if dt.kind == "f":
        if dt.itemsize == 4:
            props = struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
            bits = bytes([0x20, 31, 0])
        elif dt.itemsize == 8:
            props = struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
            bits = bytes([0x20, 63, 0])
        elif dt.itemsize == 2:
            props = struct.pack("<HHBBBBI", 0, 16, 10, 5, 0, 10, 15)
            bits = bytes([0x20, 15, 0])
        else:
            raise HDF5Error(f"unsupported float size {dt.itemsize}")
        return bytes([0x10 | 1]) + bits + struct.pack("<I", dt.itemsize) + props
    raise HDF5Error(f"unsupported dtype {dt}")


def _msg_v1(mtype: int, data: bytes, flags: int = 0) -> bytes:
    data = data + b"\0" * (_pad8(len(data)) - len(data))
    return struct.pack("<HHB3x", mtype, len(data), flags) + data


class _Writer:
    """Append-only file image; everything is laid out in memory and flushed on close."""

    GROUP_LEAF_K = 16      # up to 32 links per symbol-table node
    GROUP_INTERNAL_K = 16
    CHUNK_BTREE_K = 32     # indexed-storage K (HDF5 default) -> 64 entries per node

    def __init__(self, path: str):
        self.path = path
        self.img = bytearray(b"\0" * 96)  # superblock v0 placeholder
        self.datasets: List[Tuple[str, int]] = []

    def _align(self, n: int = 8) -> int:
        pad = (-len(self.img)) % n
        self.img += b"\0" * pad
        return len(self.img)

    def _append(self, data: bytes, align: int = 8) -> int:
        a = self._align(align)
        self.img += data
        return a

    def add_dataset(self, name: str, arr: np.ndarray, compression: Optional[str],
                    compression_opts: int, chunks: Optional[Tuple[int, ...]], shuffle: bool) -> None:
        if any(n == name for n, _ in self.datasets):
            raise ValueError(f"dataset {name!r} already exists")
        if "/" in name.strip("/"):
            raise HDF5Error("the writer only supports datasets in the root group")
        name = name.strip("/")
        arr = np.ascontiguousarray(arr)
        if arr.dtype.byteorder == ">":
            arr = arr.astype(arr.dtype.newbyteorder("<"))
        rank = arr.ndim
        msgs = []
        # dataspace v1 (+ max dims = dims)
        ds = struct.pack("<BBB5x", 1, rank, 1)
        ds += b"".join(struct.pack("<Q", s) for s in arr.shape) * 2
        msgs.append(_msg_v1(MSG_DATASPACE, ds))
        msgs.append(_msg_v1(MSG_DATATYPE, _dtype_message(arr.dtype), flags=1))
        msgs.append(_msg_v1(MSG_FILL, struct.pack("<BBBB", 2, 2, 2, 0)))  # v2, alloc late, write if set, undefined
        if compression in (None, False) or arr.size == 0 or rank == 0:
            data_addr = self._append(arr.tobytes()) if arr.size else UNDEF
            lay = struct.pack("<BB", 3, 1) + struct.pack("<QQ", data_addr, arr.nbytes)
            msgs.append(_msg_v1(MSG_LAYOUT, lay))
        else:
            if compression not in ("gzip", True):
                raise HDF5Error(f"unsupported compression {compression!r}")
            level = 4 if compression_opts is None else int(compression_opts)
            max_chunks = 2 * self.CHUNK_BTREE_K
            if chunks is None:
                rows = max(1, -(-arr.shape[0] // max_chunks))
                # aim at ~1 MiB chunks like h5py's guesser but never more than 64 chunks
                row_bytes = max(1, arr.nbytes // max(arr.shape[0], 1))
                rows = max(rows, min(arr.shape[0], max(1, (1 << 20) // row_bytes)))
                chunks = (rows,) + tuple(arr.shape[1:])
            if tuple(chunks[1:]) != tuple(arr.shape[1:]):
                raise HDF5Error("the writer chunks along the first axis only")
            n_chunks = -(-arr.shape[0] // chunks[0])
            if n_chunks > max_chunks:
                raise HDF5Error(f"{n_chunks} chunks exceed the single-node B-tree capacity {max_chunks}")
            entries = []
            chunk_elems = int(np.prod(chunks))
            for i in range(n_chunks):
                block = arr[i * chunks[0]:(i + 1) * chunks[0]]
                if block.shape[0] < chunks[0]:  # edge chunks are stored full size
                    pad = np.zeros((chunks[0] - block.shape[0],) + block.shape[1:], dtype=arr.dtype)
                    block = np.concatenate([block, pad], axis=0)
                raw = block.tobytes()
                assert len(raw) == chunk_elems * arr.dtype.itemsize
                if shuffle:
                    raw = _shuffle(raw, arr.dtype.itemsize)
                comp = zlib.compress(raw, level)
                entries.append((i * chunks[0], self._append(comp), len(comp)))
            # one leaf node of the v1 chunk B-tree, sized for 2K entries
            ndim = rank + 1
            key_size = 8 + 8 * ndim
            node = bytearray(b"TREE" + struct.pack("<BBH", 1, 0, len(entries)) + struct.pack("<QQ", UNDEF, UNDEF))
            for row0, a, sz in entries:
                node += struct.pack("<II", sz, 0) + struct.pack("<Q", row0) + b"\0" * (8 * (ndim - 1))
                node += struct.pack("<Q", a)
            # final key: one past the last chunk
            node += struct.pack("<II", 0, 0) + struct.pack("<Q", n_chunks * chunks[0]) + b"\0" * (8 * (ndim - 1))
            full = 24 + (2 * self.CHUNK_BTREE_K) * (key_size + 8) + key_size
            node += b"\0" * (full - len(node))
            btree_addr = self._append(bytes(node))
            lay = struct.pack("<BBB", 3, 2, ndim) + struct.pack("<Q", btree_addr)
            lay += b"".join(struct.pack("<I", c) for c in chunks) + struct.pack("<I", arr.dtype.itemsize)
            msgs.append(_msg_v1(MSG_LAYOUT, lay))
            filt = struct.pack("<BB6x", 1, 2 if shuffle else 1)
            if shuffle:
                filt += struct.pack("<HHHH", 2, 8, 1, 1) + b"shuffle\0" + struct.pack("<I", arr.dtype.itemsize) + b"\0" * 4
            filt += struct.pack("<HHHH", 1, 8, 1, 1) + b"deflate\0" + struct.pack("<I", level) + b"\0" * 4
            msgs.append(_msg_v1(MSG_FILTERS, filt))
        body = b"".join(msgs)
        header = struct.pack("<BBHII4x", 1, 0, len(msgs), 1, len(body)) + body
        self.datasets.append((name, self._append(header)))

    def finish(self) -> None:
        names = sorted(self.datasets, key=lambda t: t[0])
        # local heap data: offset 0 is the empty string
        heap = bytearray(b"\0" * 8)
        offs = {}
        for n, _ in names:
            offs[n] = len(heap)
            enc = n.encode("utf-8") + b"\0"
            heap += enc + b"\0" * (_pad8(len(enc)) - len(enc))
        free_off = len(heap)
        heap += struct.pack("<QQ", 1, 32) + b"\0" * 16  # one free block (next=NULL(1), size 32)
        heap_data_addr = self._append(bytes(heap))
        heap_addr = self._append(b"HEAP" + struct.pack("<B3x", 0) + struct.pack("<QQQ", len(heap), free_off,
                                                                                 heap_data_addr))
        cap = 2 * self.GROUP_LEAF_K
        if len(names) > cap:
            raise HDF5Error(f"more than {cap} datasets in the root group")
        snod = bytearray(b"SNOD" + struct.pack("<BBH", 1, 0, len(names)))
        for n, h in names:
            snod += struct.pack("<QQII16x", offs[n], h, 0, 0)
        snod += b"\0" * (8 + cap * 40 - len(snod))
        snod_addr = self._append(bytes(snod))
        tree = bytearray(b"TREE" + struct.pack("<BBH", 0, 0, 1) + struct.pack("<QQ", UNDEF, UNDEF))
        tree += struct.pack("<Q", 0) + struct.pack("<Q", snod_addr)
        tree += struct.pack("<Q", offs[names[-1][0]] if names else 0)
        full = 24 + (2 * self.GROUP_INTERNAL_K) * 16 + 8
        tree += b"\0" * (full - len(tree))
        btree_addr = self._append(bytes(tree))
        symtab = _msg_v1(MSG_SYMTAB, struct.pack("<QQ", btree_addr, heap_addr))
        root_hdr = struct.pack("<BBHII4x", 1, 0, 1, 1, len(symtab)) + symtab
        root_addr = self._append(root_hdr)
        eof = self._align(8)
        sb = bytearray(SIGNATURE)
        sb += struct.pack("<BBBBBBBB", 0, 0, 0, 0, 0, 8, 8, 0)
        sb += struct.pack("<HHI", self.GROUP_LEAF_K, self.GROUP_INTERNAL_K, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
        sb += struct.pack("<QQII", 0, root_addr, 1, 0) + struct.pack("<QQ", btree_addr, heap_addr)
        assert len(sb) == 96
        self.img[:96] = sb
        tmp = self.path + ".tmp"
        with open(tmp, "wb") as f:
            f.write(self.img)
        os.replace(tmp, self.path)


class File(Group):
    """``with File(path, 'r') as f: f['input_ids'][:]`` / ``File(path, 'w').create_dataset``."""

    def __init__(self, path: str, mode: str = "r"):
        self.filename = os.fspath(path)
        self.mode = mode
        self._w: Optional[_Writer] = None
        if mode == "r":
            r = _Reader(self.filename)
            Group.__init__(self, r, "/", r.root_header)
        elif mode in ("w", "x"):
            if mode == "x" and os.path.exists(self.filename):
                raise FileExistsError(self.filename)
            self._w = _Writer(self.filename)
            self._links = {}
        else:
            raise ValueError("mode must be 'r', 'w' or 'x'")

    def create_dataset(self, name: str, shape=None, dtype=None, data=None, compression=None,
                       compression_opts=None, chunks=None, shuffle: bool = False, **_ignored) -> None:
        if self._w is None:
            raise HDF5Error("file is not open for writing")
        if data is None:
            if shape is None:
                raise ValueError("either data or shape is required")
            data = np.zeros(shape, dtype=dtype or np.float32)
        arr = np.asarray(data)
        if dtype is not None:
            arr = arr.astype(np.dtype(dtype))
        if chunks is True:
            chunks = None
        self._w.add_dataset(name, arr, compression, compression_opts, chunks, shuffle)
        self._links[name.strip("/")] = -1

    def close(self) -> None:
        if self._w is not None:
            self._w.finish()
            self._w = None

    def __enter__(self) -> "File":
        return self

    def __exit__(self, *exc) -> None:
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
