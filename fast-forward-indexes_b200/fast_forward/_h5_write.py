"""The write side of `OnDiskIndex` without h5py: creates and extends an index file in the
reference's HDF5 layout (src/fast_forward/index/disk.py:83-85 root attributes, :123-136 the
`quantizer` group, :138-165 `_create_ds`, :243-301 `_add`).

`OnDiskIndex` writes through h5py when it is installed; this module is used when it is not, so
that an index can be created, extended and re-opened on a box that has neither h5py nor libhdf5.
What it writes is plain "earliest"-format HDF5 (HDF5 File Format Specification 3.0), the same
encodings h5py's defaults produce and `fast_forward._h5` / csrc/ffx_h5.cpp read back:

  bytes 0..4095, fixed skeleton (rewritten in place as the index grows):
    superblock v0 -> root group (version-1 object header, symbol-table message; attributes
    `num_vectors` int64, `ff_version` variable-length string in a small global heap) with its
    v1 B-tree, local heap and one symbol-table node; the object headers of `vectors`,
    `doc_ids`, `psg_ids` (dataspace v1 with unlimited first axis, datatype, fill value, layout v3
    chunked, modification time)
  bytes 4096.., appended: raw chunks (stored whole, allocated on first write), the chunk
    B-trees (a fresh copy per growth step; the layout message is repointed), and the
    `quantizer/{meta,attributes,data}` subtree (attributes: int64 / float64 / bool as the int8
    enum h5py uses / variable-length UTF-8 strings; datasets contiguous).

Growth is O(rows added): new chunks and a few KB of B-tree per `add`, dims and `num_vectors`
patched in place (`num_vectors` last, so a torn write leaves the old, consistent extent).

PARITY UNPINNED against libhdf5 (none in the image): checked against this package's own reader
only (tests/test_disk.py); h5py remains the writer wherever it is available.
"""

from __future__ import annotations

import os
import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
SIGNATURE = b"\x89HDF\r\n\x1a\n"

# ---- fixed skeleton ---------------------------------------------------------------------------
ROOT_HEADER, ROOT_HEADER_BYTES = 96, 608       # 16-byte prefix + 592 bytes of messages
ROOT_TREE = 704                                # 24 + 32 * 16 + 8 = 544 bytes (K = 16)
ROOT_HEAP, ROOT_HEAP_DATA, ROOT_HEAP_BYTES = 1248, 1280, 256
ROOT_SNOD = 1536                               # 8 + 8 * 40 = 328 bytes (leaf K = 4)
DATASET_HEADER = {"vectors": 2048, "doc_ids": 2560, "psg_ids": 3072}
DATASET_HEADER_BYTES = 512
ROOT_GCOL, ROOT_GCOL_BYTES = 3584, 512
DATA_START = 4096
NAME_OFFSET = {"doc_ids": 8, "psg_ids": 16, "quantizer": 24, "vectors": 40}  # inside the root heap
CHUNK_K = 32                                   # chunk B-tree: up to 64 entries per node


def _pad8(b: bytes) -> bytes:
    return b + b"\0" * (-len(b) % 8)


def _message(kind: int, body: bytes) -> bytes:
    body = _pad8(body)
    return struct.pack("<HHBBBB", kind, len(body), 0, 0, 0, 0) + body


def _datatype(dt: np.dtype) -> bytes:
    dt = np.dtype(dt)
    if dt.kind in "iu":
        return struct.pack("<BBBBIHH", 0x10, 0x08 if dt.kind == "i" else 0, 0, 0, dt.itemsize, 0, 8 * dt.itemsize)
    if dt.kind == "f":
        exp_bits, mantissa = {2: (5, 10), 4: (8, 23), 8: (11, 52)}[dt.itemsize]
        return struct.pack("<BBBBIHHBBBBI", 0x11, 0x20, 8 * dt.itemsize - 1, 0, dt.itemsize, 0, 8 * dt.itemsize,
                           mantissa, exp_bits, 0, mantissa, (1 << (exp_bits - 1)) - 1)
    if dt.kind == "S":
        return struct.pack("<BBBBI", 0x13, 0x01, 0, 0, dt.itemsize)  # null-padded
    if dt.kind == "b":  # numpy bool: enum {FALSE = 0, TRUE = 1} over int8, as h5py stores it
        base = struct.pack("<BBBBIHH", 0x10, 0x08, 0, 0, 1, 0, 8)
        return struct.pack("<BBBBI", 0x18, 2, 0, 0, 1) + base + _pad8(b"FALSE\0") + _pad8(b"TRUE\0") + b"\0\1"
    raise TypeError(f"vectors of dtype {dt} cannot be stored")


_VLEN_STRING = struct.pack("<BBBBI", 0x19, 0x01, 0x01, 0, 16) + struct.pack("<BBBBI", 0x13, 0x00, 0, 0, 1)


def _dataspace(shape: tuple, unlimited_first: bool) -> bytes:
    body = struct.pack("<BBBBI", 1, len(shape), 1 if unlimited_first else 0, 0, 0)
    body += b"".join(struct.pack("<Q", d) for d in shape)
    if unlimited_first:
        body += struct.pack("<Q", UNDEF) + b"".join(struct.pack("<Q", d) for d in shape[1:])
    return body


def _attribute(name: str, datatype: bytes, shape: tuple, value: bytes) -> bytes:
    nm = name.encode("utf-8") + b"\0"
    space = _dataspace(shape, False)
    return (struct.pack("<BBHHH", 1, 0, len(nm), len(datatype), len(space)) + _pad8(nm) + _pad8(datatype)
            + _pad8(space) + value)


def _object_header(messages: list[bytes], capacity: int | None = None) -> bytes:
    """Version-1 object header; with `capacity` the message area is padded by a NIL message."""
    body = b"".join(messages)
    count = len(messages)
    if capacity is not None:
        room = capacity - 16 - len(body)
        if room < 8:
            raise ValueError("object header does not fit its reserved space")
        body += struct.pack("<HHBBBB", 0, room - 8, 0, 0, 0, 0) + b"\0" * (room - 8)
        count += 1
    return struct.pack("<BBHII", 1, 0, count, 1, len(body)) + b"\0" * 4 + body


def _global_heap(strings: list[bytes], size: int | None = None) -> bytes:
    body = b""
    for i, raw in enumerate(strings, 1):
        body += struct.pack("<HHIQ", i, 1, 0, len(raw)) + _pad8(raw)
    size = size or max(4096, 16 + len(body) + 16)
    size += -size % 8
    free = size - 16 - len(body)
    if free < 16:
        raise ValueError("strings do not fit the global heap collection")
    return b"GCOL" + struct.pack("<BBBBQ", 1, 0, 0, 0, size) + body + struct.pack("<HHIQ", 0, 0, 0, free) + b"\0" * (free - 16)


class _Blob:
    """A self-contained run of structures to be appended at a known file address."""

    def __init__(self, base: int) -> None:
        self.base = base
        self.buf = bytearray()
        self.strings: list[bytes] = []
        self.string_refs: list[tuple[int, int, int]] = []  # (offset in buf, length, heap index)

    def emit(self, data: bytes) -> int:
        self.buf += b"\0" * (-len(self.buf) % 8)
        at = self.base + len(self.buf)
        self.buf += data
        return at

    def attribute(self, name: str, value) -> tuple[bytes, list[tuple[int, int, int]]]:
        """-> (message body, [(offset of a heap reference inside the body, length, index)])"""
        if isinstance(value, (str, np.str_)):
            raw = str(value).encode("utf-8")
            self.strings.append(raw)
            body = _attribute(name, _VLEN_STRING, (), struct.pack("<IQI", len(raw), 0, len(self.strings)))
            return body, [(len(body) - 16, len(raw), len(self.strings))]
        arr = np.asarray(value)
        if arr.dtype.kind == "U":
            arr = np.char.encode(arr, "utf-8")
        if arr.dtype.kind in "iu" and arr.dtype.itemsize < 8 and arr.ndim == 0:
            arr = arr.astype(np.int64)
        return _attribute(name, _datatype(arr.dtype), arr.shape, np.ascontiguousarray(arr).tobytes()), []

    def header(self, messages: list[bytes], attrs: dict) -> int:
        refs = []
        for name, value in attrs.items():
            if value is None:  # (h5py has no encoding for None either; the key is simply absent)
                continue
            body, found = self.attribute(name, value)
            at = sum(len(m) for m in messages) + 8  # body starts behind its 8-byte message header
            refs += [(at + off, n, idx) for off, n, idx in found]
            messages = messages + [_message(0x0C, body)]
        addr = self.emit(_object_header(messages))
        self.string_refs += [(addr - self.base + 16 + off, n, idx) for off, n, idx in refs]
        return addr

    def group(self, links: dict[str, int], attrs: dict) -> int:
        """A symbol-table group with at most 8 links (one leaf node)."""
        assert len(links) <= 8
        heap = bytearray(b"\0" * 8)
        entries = []
        for name in sorted(links):
            entries.append((len(heap), links[name]))
            heap += _pad8(name.encode("utf-8") + b"\0")
        free = len(heap)
        heap += struct.pack("<QQ", 1, 32) + b"\0" * 16
        heap_data = self.emit(bytes(heap))
        heap_addr = self.emit(b"HEAP" + struct.pack("<BBBBQQQ", 0, 0, 0, 0, len(heap), free, heap_data))
        body = b"".join(struct.pack("<QQII", off, addr, 0, 0) + b"\0" * 16 for off, addr in entries)
        snod = self.emit(b"SNOD" + struct.pack("<BBH", 1, 0, len(entries)) + body + b"\0" * (40 * (8 - len(entries))))
        keys = struct.pack("<QQQ", 0, snod, entries[-1][0] if entries else 0)
        tree = self.emit(b"TREE" + struct.pack("<BBHQQ", 0, 0, 1, UNDEF, UNDEF) + keys + b"\0" * (544 - 24 - len(keys)))
        return self.header([_message(0x11, struct.pack("<QQ", tree, heap_addr))], attrs)

    def dataset(self, data: np.ndarray) -> int:
        data = np.asarray(data, order="C")
        addr = self.emit(data.tobytes()) if data.size else UNDEF
        return self.header([_message(0x01, _dataspace(data.shape, False)), _message(0x03, _datatype(data.dtype)),
                            _message(0x05, struct.pack("<BBBB", 2, 2, 0, 0)),
                            _message(0x08, struct.pack("<BBQQ", 3, 1, addr, data.nbytes))], {})

    def finish(self) -> bytes:
        if self.strings:
            heap = self.emit(_global_heap(self.strings))
            for at, n, idx in self.string_refs:
                self.buf[at:at + 16] = struct.pack("<IQI", n, heap, idx)
        return bytes(self.buf)


class _Dataset:
    def __init__(self, name: str, dtype: np.dtype, inner: tuple, chunk_rows: int) -> None:
        self.name, self.dtype, self.inner, self.chunk_rows = name, np.dtype(dtype), tuple(inner), int(chunk_rows)
        self.rows = 0
        self.tree = UNDEF
        self.chunks: dict[int, int] = {}  # chunk number -> file address
        self.row_bytes = self.dtype.itemsize * int(np.prod(self.inner, dtype=np.int64))

    @property
    def chunk_bytes(self) -> int:
        return self.chunk_rows * self.row_bytes

    def header(self) -> bytes:
        rank = 1 + len(self.inner)
        layout = struct.pack("<BBBQ", 3, 2, rank + 1, self.tree)
        layout += b"".join(struct.pack("<I", c) for c in (self.chunk_rows,) + self.inner) + struct.pack("<I", self.dtype.itemsize)
        return _object_header([
            _message(0x01, _dataspace((self.rows,) + self.inner, True)), _message(0x03, _datatype(self.dtype)),
            _message(0x05, struct.pack("<BBBB", 2, 3, 0, 0)), _message(0x08, layout),
            _message(0x12, struct.pack("<BBBBI", 1, 0, 0, 0, 0))], DATASET_HEADER_BYTES)


class IndexFile:
    """One index file, open for writing.  Use `create` or `open_existing`, then `close`."""

    def __init__(self, path, fh) -> None:
        self.path, self._fh = path, fh
        self.num_vectors = 0
        self.version = ""
        self.datasets: dict[str, _Dataset] = {}
        self.quantizer_header = UNDEF
        self._end = DATA_START

    # ---- opening --------------------------------------------------------------------------------
    @classmethod
    def create(cls, path, version: str) -> "IndexFile":
        self = cls(path, open(path, "w+b"))
        self.version = version
        self._write_skeleton()
        return self

    @classmethod
    def open_existing(cls, path) -> "IndexFile":
        """Re-open a file this class wrote.  ValueError for any other HDF5 file (extending a file
        written by libhdf5 needs h5py)."""
        self = cls(path, open(path, "r+b"))
        try:
            self._read_skeleton()
        except Exception:
            self._fh.close()
            raise
        return self

    def close(self) -> None:
        if self._fh is not None:
            self._fh.flush()
            self._fh.close()
            self._fh = None

    def __enter__(self) -> "IndexFile":
        return self

    def __exit__(self, *exc) -> None:
        self.close()

    # ---- skeleton -------------------------------------------------------------------------------
    def _put(self, at: int, data: bytes) -> None:
        self._fh.seek(at)
        self._fh.write(data)

    def _root_header(self) -> tuple[bytes, int]:
        version = self.version.encode("utf-8")
        messages = [
            _message(0x11, struct.pack("<QQ", ROOT_TREE, ROOT_HEAP)),
            _message(0x0C, _attribute("num_vectors", _datatype(np.int64), (), struct.pack("<q", self.num_vectors))),
            _message(0x0C, _attribute("ff_version", _VLEN_STRING, (), struct.pack("<IQI", len(version), ROOT_GCOL, 1))),
            _message(0x0C, _attribute("ffx_writer", _datatype(np.int64), (), struct.pack("<q", 1))),
        ]
        count_at = ROOT_HEADER + 16 + len(messages[0]) + len(messages[1]) - 8  # the int64 of num_vectors
        return _object_header(messages, ROOT_HEADER_BYTES), count_at

    def _links(self) -> dict[str, int]:
        links = {name: DATASET_HEADER[name] for name in self.datasets}
        if self.quantizer_header != UNDEF:
            links["quantizer"] = self.quantizer_header
        return links

    def _write_skeleton(self) -> None:
        for ds in self.datasets.values():
            if ds.tree is None:  # chunks were allocated since the last flush
                ds.tree = self._chunk_tree(ds)
        header, self._count_at = self._root_header()
        links = self._links()
        heap = bytearray(ROOT_HEAP_BYTES)
        for name, off in NAME_OFFSET.items():
            heap[off:off + len(name)] = name.encode()
        heap[48:64] = struct.pack("<QQ", 1, ROOT_HEAP_BYTES - 48)
        names = sorted(links)
        snod = b"SNOD" + struct.pack("<BBH", 1, 0, len(names))
        snod += b"".join(struct.pack("<QQII", NAME_OFFSET[n], links[n], 0, 0) + b"\0" * 16 for n in names)
        snod += b"\0" * (40 * (8 - len(names)))
        keys = struct.pack("<QQQ", 0, ROOT_SNOD, NAME_OFFSET[names[-1]] if names else 0)
        tree = b"TREE" + struct.pack("<BBHQQ", 0, 0, 1, UNDEF, UNDEF) + keys
        superblock = SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, 4, 16, 0)
        superblock += struct.pack("<QQQQ", 0, UNDEF, max(self._end, DATA_START), UNDEF)
        superblock += struct.pack("<QQII", 0, ROOT_HEADER, 1, 0) + struct.pack("<QQ", ROOT_TREE, ROOT_HEAP)
        image = bytearray(DATA_START)
        image[0:len(superblock)] = superblock
        image[ROOT_HEADER:ROOT_HEADER + len(header)] = header
        image[ROOT_TREE:ROOT_TREE + len(tree)] = tree
        image[ROOT_HEAP:ROOT_HEAP + 32] = b"HEAP" + struct.pack("<BBBBQQQ", 0, 0, 0, 0, ROOT_HEAP_BYTES, 48, ROOT_HEAP_DATA)
        image[ROOT_HEAP_DATA:ROOT_HEAP_DATA + ROOT_HEAP_BYTES] = heap
        image[ROOT_SNOD:ROOT_SNOD + len(snod)] = snod
        for name, ds in self.datasets.items():
            block = ds.header()
            image[DATASET_HEADER[name]:DATASET_HEADER[name] + len(block)] = block
        heap_block = _global_heap([self.version.encode("utf-8")], ROOT_GCOL_BYTES)
        image[ROOT_GCOL:ROOT_GCOL + len(heap_block)] = heap_block
        self._put(0, bytes(image))
        if self._end == DATA_START:
            self._fh.truncate(DATA_START)

    def _read_skeleton(self) -> None:
        fh = self._fh
        fh.seek(0)
        image = fh.read(DATA_START)
        _, count_at = self._root_header()
        marker = _pad8(b"ffx_writer\0")
        if len(image) < DATA_START or image[:8] != SIGNATURE or marker not in image[ROOT_HEADER:ROOT_HEADER + ROOT_HEADER_BYTES]:
            raise ValueError(f"{self.path} was not written by this package's HDF5 writer; extending it needs h5py")
        self._count_at = count_at
        self.num_vectors = struct.unpack_from("<q", image, count_at)[0]
        heap_at = ROOT_GCOL + 16
        n = struct.unpack_from("<Q", image, heap_at + 8)[0]
        self.version = image[heap_at + 16:heap_at + 16 + n].decode("utf-8")
        fh.seek(0, os.SEEK_END)
        self._end = max(fh.tell(), DATA_START)
        count = struct.unpack_from("<H", image, ROOT_SNOD + 6)[0]
        by_offset = {off: name for name, off in NAME_OFFSET.items()}
        for i in range(count):
            name_off, addr = struct.unpack_from("<QQ", image, ROOT_SNOD + 8 + 40 * i)
            name = by_offset[name_off]
            if name == "quantizer":
                self.quantizer_header = addr
                continue
            self.datasets[name] = self._parse_dataset(name, image[addr:addr + DATASET_HEADER_BYTES])

    def _parse_dataset(self, name: str, block: bytes) -> _Dataset:
        at, shape, dtype, chunk, tree = 16, None, None, None, UNDEF
        for _ in range(struct.unpack_from("<H", block, 2)[0]):
            kind, size = struct.unpack_from("<HH", block, at)
            body = block[at + 8:at + 8 + size]
            if kind == 0x01:
                shape = struct.unpack_from(f"<{body[1]}Q", body, 8)
            elif kind == 0x03:
                cls, bits, size_of = body[0] & 15, body[1], struct.unpack_from("<I", body, 4)[0]
                dtype = np.dtype(f"S{size_of}") if cls == 3 else np.dtype(
                    f"f{size_of}" if cls == 1 else f"{'i' if bits & 8 else 'u'}{size_of}")
            elif kind == 0x08:
                rank1 = body[2]
                tree = struct.unpack_from("<Q", body, 3)[0]
                chunk = struct.unpack_from(f"<{rank1}I", body, 11)
            at += 8 + size
        ds = _Dataset(name, dtype, shape[1:], chunk[0])
        ds.rows, ds.tree = shape[0], tree
        if tree != UNDEF:
            self._walk_chunks(ds, tree)
        return ds

    def _walk_chunks(self, ds: _Dataset, node: int) -> None:
        self._fh.seek(node)
        head = self._fh.read(24)
        if head[:4] != b"TREE":
            raise ValueError("corrupt chunk B-tree")
        level, used = head[5], struct.unpack_from("<H", head, 6)[0]
        key = 8 + 8 * (len(ds.inner) + 2)
        body = self._fh.read(used * (key + 8) + key)
        for i in range(used):
            row = struct.unpack_from("<Q", body, i * (key + 8) + 8)[0]
            child = struct.unpack_from("<Q", body, i * (key + 8) + key)[0]
            if level:
                self._walk_chunks(ds, child)
            else:
                ds.chunks[row // ds.chunk_rows] = child

    # ---- datasets -------------------------------------------------------------------------------
    def create_datasets(self, dim: int, dtype, init_size: int, chunk_size: int, id_width: int) -> None:
        """`_create_ds` (disk.py:138-165): vectors (init_size, dim) chunked (chunk_size, dim),
        doc_ids / psg_ids (init_size,) of fixed-width bytes."""
        self.datasets["vectors"] = _Dataset("vectors", dtype, (dim,), chunk_size)
        self.datasets["doc_ids"] = _Dataset("doc_ids", f"S{id_width}", (), chunk_size)
        self.datasets["psg_ids"] = _Dataset("psg_ids", f"S{id_width}", (), chunk_size)
        for ds in self.datasets.values():
            ds.rows = init_size
        self._write_skeleton()

    @property
    def capacity(self) -> int:
        return self.datasets["vectors"].rows if self.datasets else 0

    def resize(self, rows: int) -> None:
        for ds in self.datasets.values():
            ds.rows = rows
        self._flush_headers()

    def _append(self, data: bytes) -> int:
        self._end += -self._end % 8
        at = self._end
        self._put(at, data)
        self._end += len(data)
        return at

    def _chunk_address(self, ds: _Dataset, number: int) -> int:
        if number not in ds.chunks:
            self._end += -self._end % 8
            ds.chunks[number] = self._end
            self._end += ds.chunk_bytes
            self._fh.truncate(self._end)  # zero-filled (sparse) until written
            ds.tree = None  # rebuilt on the next flush
        return ds.chunks[number]

    def write_rows(self, name: str, row0: int, values: np.ndarray) -> None:
        """Rows [row0, row0 + len(values)) of a dataset (`fp[name][a:b] = values`)."""
        ds = self.datasets[name]
        values = np.ascontiguousarray(values, dtype=ds.dtype)
        if row0 + len(values) > ds.rows:
            raise ValueError("rows beyond the dataset's extent")
        done = 0
        while done < len(values):
            number, inside = divmod(row0 + done, ds.chunk_rows)
            take = min(len(values) - done, ds.chunk_rows - inside)
            self._put(self._chunk_address(ds, number) + inside * ds.row_bytes, values[done:done + take].tobytes())
            done += take

    def write_at(self, name: str, positions, values) -> None:
        """Scattered single rows (`fp["doc_ids"][positions] = values`), positions increasing."""
        positions = np.asarray(positions, dtype=np.int64)
        if len(positions) == 0:
            return
        ds = self.datasets[name]
        values = np.asarray([v.encode("utf-8") if isinstance(v, str) else v for v in values], dtype=ds.dtype)
        breaks = np.flatnonzero(np.diff(positions) != 1) + 1  # runs of consecutive rows
        for lo, hi in zip(np.concatenate([[0], breaks]), np.concatenate([breaks, [len(positions)]])):
            self.write_rows(name, int(positions[lo]), values[lo:hi])

    def _chunk_tree(self, ds: _Dataset) -> int:
        if not ds.chunks:
            return UNDEF
        rank = 1 + len(ds.inner)

        def key(row: int, nbytes: int) -> bytes:
            return struct.pack("<IIQ", nbytes, 0, row) + b"\0" * (8 * rank)

        entries = [(n * ds.chunk_rows, key(n * ds.chunk_rows, ds.chunk_bytes), addr) for n, addr in sorted(ds.chunks.items())]
        end_key = key((max(ds.chunks) + 1) * ds.chunk_rows, 0)
        level, width = 0, 2 * CHUNK_K
        node_bytes = 24 + width * (len(end_key) + 8) + len(end_key)
        while True:
            groups = [entries[i:i + width] for i in range(0, len(entries), width)]
            self._end += -self._end % 8
            addrs = [self._end + i * node_bytes for i in range(len(groups))]
            image = bytearray(node_bytes * len(groups))
            parents = []
            for g, (grp, addr) in enumerate(zip(groups, addrs)):
                last = groups[g + 1][0][1] if g + 1 < len(groups) else end_key
                body = b"".join(k + struct.pack("<Q", child) for _, k, child in grp) + last
                node = b"TREE" + struct.pack("<BBHQQ", 1, level, len(grp), addrs[g - 1] if g else UNDEF,
                                             addrs[g + 1] if g + 1 < len(groups) else UNDEF) + body
                image[g * node_bytes:g * node_bytes + len(node)] = node
                parents.append((grp[0][0], grp[0][1], addr))
            self._append(bytes(image))
            if len(groups) == 1:
                return addrs[0]
            entries, level = parents, level + 1

    def _flush_headers(self) -> None:
        for name, ds in self.datasets.items():
            if ds.tree is None:
                ds.tree = self._chunk_tree(ds)
            self._put(DATASET_HEADER[name], ds.header())
        self._put(40, struct.pack("<Q", self._end))  # superblock: end-of-file address
        self._fh.truncate(max(self._end, DATA_START))

    def set_num_vectors(self, count: int) -> None:
        """Publishes the rows written so far: headers and B-trees first, the counter last."""
        self._flush_headers()
        self._fh.flush()
        self.num_vectors = int(count)
        self._put(self._count_at, struct.pack("<q", self.num_vectors))
        self._fh.flush()

    # ---- quantizer ------------------------------------------------------------------------------
    def set_quantizer(self, meta: dict, attributes: dict, data: dict) -> None:
        """`quantizer/{meta,attributes,data}` (disk.py:123-136); replaces an earlier one."""
        self._end += -self._end % 8
        blob = _Blob(self._end)
        arrays = blob.group({key: blob.dataset(value) for key, value in data.items()}, {})
        subtree = {"meta": blob.group({}, dict(meta)), "attributes": blob.group({}, dict(attributes)), "data": arrays}
        self.quantizer_header = blob.group(subtree, {})
        self._append(blob.finish())
        self._write_skeleton()
        self._flush_headers()
        self._fh.flush()
