"""TEST INFRASTRUCTURE — a small HDF5 *writer*, the counterpart the native reader
(`fast-forward-indexes_b200/csrc/ffx_h5.cpp`) is exercised against.

Neither libhdf5 nor h5py exists in the build image or on the GPU box, so no file written by
the real library is available; this module restates the on-disk structures h5py's default
settings (libver 'earliest') produce for the objects `OnDiskIndex` creates, following the
HDF5 File Format Specification 3.0:

  superblock v0 with the root symbol-table entry; version-1 object headers (optionally with a
  continuation block and NIL padding); groups as symbol table message + v1 B-tree (type 0) +
  SNOD leaves + local heap; datasets with dataspace v1, datatype v1, fill value v2,
  modification time, layout v3 (compact / contiguous / chunked through a v1 B-tree of type 1,
  edge chunks stored whole, unwritten chunks absent); attributes v1 (integers, floats, numpy
  bool as an int8 enum, fixed strings, variable-length UTF-8 strings through a global heap).

`modern=True` writes the newer encodings the reader also accepts: superblock v2, version-2
object headers ("OHDR"/"OCHK"), compact link messages, dataspace v2, attributes v3.
Checksums in those structures are written as zeros (the reader does not verify them).

PARITY UNPINNED against libhdf5: writer and reader are two restatements of the same
specification by the same author; a box with h5py should re-run tests/test_h5.py's
`h5py`-gated cases to pin both.
"""

from __future__ import annotations

import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


class Group:
    def __init__(self):
        self.attrs: dict = {}
        self.children: dict[str, "Group | Dataset"] = {}


class Dataset:
    def __init__(self, data: np.ndarray, chunks: tuple | None = None, maxshape: tuple | None = None,
                 compact: bool = False, missing_chunks: set | None = None):
        self.attrs: dict = {}
        self.data = np.asarray(data, order="C")  # (ascontiguousarray would turn a scalar into shape (1,))
        self.chunks = chunks
        self.maxshape = maxshape
        self.compact = compact
        self.missing_chunks = missing_chunks or set()


def pad8(b: bytes) -> bytes:
    return b + b"\0" * (-len(b) % 8)


# ---- datatype / dataspace encodings ----------------------------------------------------------
def dtype_message(dt: np.dtype) -> bytes:
    dt = np.dtype(dt)
    if dt.kind in "iu":
        bits = 0x08 if dt.kind == "i" else 0
        return struct.pack("<BBBBIHH", 0x10, bits, 0, 0, dt.itemsize, 0, 8 * dt.itemsize)
    if dt.kind == "f":
        exp_size, mant = {2: (5, 10), 4: (8, 23), 8: (11, 52)}[dt.itemsize]
        sign_loc = 8 * dt.itemsize - 1
        return struct.pack("<BBBBIHHBBBBI", 0x11, 0x20, sign_loc, 0, dt.itemsize, 0, 8 * dt.itemsize,
                           mant, exp_size, 0, mant, (1 << (exp_size - 1)) - 1)
    if dt.kind == "S":
        return struct.pack("<BBBBI", 0x13, 0x01, 0, 0, dt.itemsize)  # null-padded, ASCII
    if dt.kind == "b":  # h5py: enum {FALSE=0, TRUE=1} over int8
        base = struct.pack("<BBBBIHH", 0x10, 0x08, 0, 0, 1, 0, 8)
        return struct.pack("<BBBBI", 0x18, 2, 0, 0, 1) + base + pad8(b"FALSE\0") + pad8(b"TRUE\0") + b"\0\1"
    raise TypeError(f"no HDF5 encoding for {dt}")


VLEN_STR = struct.pack("<BBBBI", 0x19, 0x01, 0x01, 0, 16) + struct.pack("<BBBBI", 0x13, 0x00, 0, 0, 1)


def dataspace_message(shape: tuple, maxshape: tuple | None, modern: bool) -> bytes:
    flags = 1 if maxshape is not None else 0
    if modern:
        head = struct.pack("<BBBB", 2, len(shape), flags, 1 if shape else 0)
    else:
        head = struct.pack("<BBBBI", 1, len(shape), flags, 0, 0)
    body = b"".join(struct.pack("<Q", d) for d in shape)
    if maxshape is not None:
        body += b"".join(struct.pack("<Q", UNDEF if d is None else d) for d in maxshape)
    return head + body


class Writer:
    def __init__(self, modern: bool = False, group_leaf_k: int = 4, group_node_k: int = 16, chunk_k: int = 32,
                 split_headers: bool = False, user_block: int = 0):
        self.modern = modern
        self.leaf_k, self.node_k, self.chunk_k = group_leaf_k, group_node_k, chunk_k
        self.split_headers = split_headers
        self.base = user_block
        self.buf = bytearray(b"\0" * user_block)
        self.heap_objects: list[bytes] = []  # global heap: variable-length strings
        self.heap_refs: list[tuple[int, int, int]] = []  # (position of the reference, length, object index)

    # ---- raw allocation (addresses are relative to the base address) -------------------------
    def alloc(self, n: int, align: int = 8) -> int:
        self.buf += b"\0" * (-(len(self.buf) - self.base) % align)
        addr = len(self.buf) - self.base
        self.buf += b"\0" * n
        return addr

    def put(self, addr: int, data: bytes) -> None:
        self.buf[self.base + addr:self.base + addr + len(data)] = data

    def emit(self, data: bytes, align: int = 8) -> int:
        addr = self.alloc(len(data), align)
        self.put(addr, data)
        return addr

    # ---- attributes ----------------------------------------------------------------------------
    def attr_message(self, name: str, value) -> bytes:
        refs = []
        if isinstance(value, str):
            raw = value.encode("utf-8")
            self.heap_objects.append(raw)
            dt, shape = VLEN_STR, ()
            data = struct.pack("<IQI", len(raw), 0, len(self.heap_objects))
            refs.append((0, len(raw), len(self.heap_objects)))
        else:
            arr = np.asarray(value)
            if arr.dtype.kind == "U":
                arr = np.char.encode(arr, "utf-8")
            if arr.dtype.kind == "O":
                raise TypeError("object arrays are not written")
            dt, shape, data = dtype_message(arr.dtype), arr.shape, np.ascontiguousarray(arr).tobytes()
        space = dataspace_message(shape, None, self.modern)
        nm = name.encode("utf-8") + b"\0"
        if self.modern:
            head = struct.pack("<BBHHHB", 3, 0, len(nm), len(dt), len(space), 1)
            body = head + nm + dt + space
        else:
            head = struct.pack("<BBHHH", 1, 0, len(nm), len(dt), len(space))
            body = head + pad8(nm) + pad8(dt) + pad8(space)
        self._pending_refs = [(len(body) + pos, n, idx) for pos, n, idx in refs]
        return body + data

    # ---- object headers ------------------------------------------------------------------------
    def object_header(self, messages: list[tuple[int, bytes]], attrs: dict) -> int:
        msgs = list(messages)
        ref_marks = []
        for name, value in attrs.items():
            body = self.attr_message(name, value)
            ref_marks.append((len(msgs), self._pending_refs))
            msgs.append((0x0C, body))
        return self._header_v2(msgs, ref_marks) if self.modern else self._header_v1(msgs, ref_marks)

    def _header_v1(self, msgs, ref_marks) -> int:
        def block(items):
            out, where = b"", {}
            for i, (typ, body) in items:
                body = pad8(body)
                where[i] = len(out) + 8
                out += struct.pack("<HHBBBB", typ, len(body), 0, 0, 0, 0) + body
            return out, where

        items = list(enumerate(msgs))
        first, rest = items, []
        if self.split_headers and len(items) > 2:
            first, rest = items[:2], items[2:]
        n_msgs = len(items)
        first_bytes, first_pos = block(first)
        first_bytes += struct.pack("<HHBBBB", 0, 8, 0, 0, 0, 0) + b"\0" * 8  # a NIL message, as left by deletions
        n_msgs += 1
        rest_addr, rest_pos = None, {}
        if rest:
            rest_bytes, rest_pos = block(rest)
            rest_addr = self.emit(rest_bytes)
            first_bytes += struct.pack("<HHBBBB", 0x10, 16, 0, 0, 0, 0) + struct.pack("<QQ", rest_addr, len(rest_bytes))
            n_msgs += 1
        head = struct.pack("<BBHII", 1, 0, n_msgs, 1, len(first_bytes)) + b"\0" * 4
        addr = self.emit(head + first_bytes)
        for i, refs in ref_marks:
            at = addr + 16 + first_pos[i] if i in first_pos else rest_addr + rest_pos[i]
            self.heap_refs += [(at + pos, n, idx) for pos, n, idx in refs]
        return addr

    def _header_v2(self, msgs, ref_marks) -> int:
        def block(items):
            out, where = b"", {}
            for i, (typ, body) in items:
                where[i] = len(out) + 4
                out += struct.pack("<BHB", typ, len(body), 0) + body
            return out, where

        items = list(enumerate(msgs))
        first, rest = items, []
        if self.split_headers and len(items) > 2:
            first, rest = items[:2], items[2:]
        first_bytes, first_pos = block(first)
        rest_addr, rest_pos = None, {}
        if rest:
            rest_bytes, rest_pos = block(rest)
            chunk = b"OCHK" + rest_bytes + b"\0" * 4
            rest_addr = self.emit(chunk)
            first_bytes += struct.pack("<BHB", 0x10, 16, 0) + struct.pack("<QQ", rest_addr, len(chunk))
        first_bytes += b"\0" * 3  # a gap too small for a message header
        prefix = b"OHDR" + struct.pack("<BB", 2, 0x02) + struct.pack("<I", len(first_bytes))  # 4-byte chunk size
        addr = self.emit(prefix + first_bytes + b"\0" * 4)
        for i, refs in ref_marks:
            at = addr + len(prefix) + first_pos[i] if i in first_pos else rest_addr + 4 + rest_pos[i]
            self.heap_refs += [(at + pos, n, idx) for pos, n, idx in refs]
        return addr

    # ---- datasets ------------------------------------------------------------------------------
    def chunk_tree(self, ds: Dataset) -> int:
        data, rank = ds.data, ds.data.ndim
        rows = ds.chunks[0]
        assert tuple(ds.chunks[1:]) == data.shape[1:], "the index layout chunks whole rows only"
        row_bytes = data.dtype.itemsize * int(np.prod(data.shape[1:], dtype=np.int64))
        n_chunks = -(-data.shape[0] // rows)

        def key(row, nbytes):
            return struct.pack("<II", nbytes, 0) + struct.pack("<Q", row) + b"\0" * (8 * rank)

        entries = []  # (first row, key bytes, child address)
        for c in range(n_chunks):
            if c in ds.missing_chunks:
                continue
            block = np.zeros((rows,) + data.shape[1:], data.dtype)  # edge chunks are stored whole
            part = data[c * rows:(c + 1) * rows]
            block[:len(part)] = part
            entries.append((c * rows, key(c * rows, rows * row_bytes), self.emit(block.tobytes())))
        if not entries:
            return UNDEF
        level, width = 0, 2 * self.chunk_k
        end_key = key(n_chunks * rows, 0)
        while True:
            groups = [entries[i:i + width] for i in range(0, len(entries), width)]
            node_size = 24 + width * (len(end_key) + 8) + len(end_key)
            addrs = [self.alloc(node_size) for _ in groups]
            parents = []
            for g, (grp, addr) in enumerate(zip(groups, addrs)):
                last = groups[g + 1][0][1] if g + 1 < len(groups) else end_key
                body = b"".join(k + struct.pack("<Q", child) for _, k, child in grp) + last
                left = addrs[g - 1] if g else UNDEF
                right = addrs[g + 1] if g + 1 < len(groups) else UNDEF
                self.put(addr, b"TREE" + struct.pack("<BBHQQ", 1, level, len(grp), left, right) + body)
                parents.append((grp[0][0], grp[0][1], addr))
            if len(groups) == 1:
                return addrs[0]
            entries, level = parents, level + 1

    def dataset(self, ds: Dataset) -> int:
        data = ds.data
        msgs = [(0x01, dataspace_message(data.shape, ds.maxshape, self.modern)),
                (0x03, dtype_message(data.dtype)),
                (0x05, struct.pack("<BBBB", 2, 3 if ds.chunks else 2, 0, 0)),
                (0x12, struct.pack("<BBBBI", 1, 0, 0, 0, 1_700_000_000))]
        if ds.compact:
            raw = data.tobytes()
            layout = struct.pack("<BBH", 3, 0, len(raw)) + raw
        elif ds.chunks:
            tree = self.chunk_tree(ds)
            layout = struct.pack("<BBBQ", 3, 2, data.ndim + 1, tree)
            layout += b"".join(struct.pack("<I", c) for c in ds.chunks) + struct.pack("<I", data.dtype.itemsize)
        else:
            addr = self.emit(data.tobytes()) if data.size else UNDEF
            layout = struct.pack("<BBQQ", 3, 1, addr, data.nbytes)
        msgs.append((0x08, layout))
        return self.object_header(msgs, ds.attrs)

    # ---- groups --------------------------------------------------------------------------------
    def group(self, g: Group) -> tuple[int, int, int]:
        """-> (object header address, B-tree address, local heap address); the last two are
        UNDEF for link-message groups."""
        links = sorted((name, self.node(child)) for name, child in g.children.items())
        if self.modern:
            msgs = [(0x02, struct.pack("<BBQQ", 0, 0, UNDEF, UNDEF)), (0x0A, struct.pack("<BB", 0, 0))]
            for name, addr in links:
                nm = name.encode("utf-8")
                msgs.append((0x06, struct.pack("<BBB", 1, 0x10, 1) + struct.pack("<B", len(nm)) + nm
                             + struct.pack("<Q", addr)))
            return self.object_header(msgs, g.attrs), UNDEF, UNDEF

        heap = bytearray(b"\0" * 8)  # offset 0: the empty name
        offsets = []
        for name, _ in links:
            offsets.append(len(heap))
            heap += pad8(name.encode("utf-8") + b"\0")
        free = len(heap)
        heap += struct.pack("<QQ", 1, 16) + b"\0" * 16  # one free block: next (1 = none), size
        heap_data = self.emit(bytes(heap))
        heap_addr = self.emit(b"HEAP" + struct.pack("<BBBBQQQ", 0, 0, 0, 0, len(heap), free, heap_data))

        per_leaf = 2 * self.leaf_k
        entries = []  # (largest name offset in the subtree, child address)
        for i in range(0, max(len(links), 1), per_leaf):
            part = list(zip(offsets[i:i + per_leaf], links[i:i + per_leaf]))
            body = b"".join(struct.pack("<QQII", off, addr, 0, 0) + b"\0" * 16 for off, (_, addr) in part)
            body += b"\0" * (40 * (per_leaf - len(part)))
            snod = self.emit(b"SNOD" + struct.pack("<BBH", 1, 0, len(part)) + body)
            entries.append((part[-1][0] if part else 0, snod))
        level, width = 0, 2 * self.node_k
        while True:
            groups = [entries[i:i + width] for i in range(0, len(entries), width)]
            addrs = [self.alloc(24 + width * 16 + 8) for _ in groups]
            parents = []
            for gi, (grp, addr) in enumerate(zip(groups, addrs)):
                first_key = 0 if gi == 0 else groups[gi - 1][-1][0]
                body = struct.pack("<Q", first_key) + b"".join(struct.pack("<QQ", child, k) for k, child in grp)
                left = addrs[gi - 1] if gi else UNDEF
                right = addrs[gi + 1] if gi + 1 < len(groups) else UNDEF
                self.put(addr, b"TREE" + struct.pack("<BBHQQ", 0, level, len(grp), left, right) + body)
                parents.append((grp[-1][0], addr))
            if len(groups) == 1:
                tree = addrs[0]
                break
            entries, level = parents, level + 1
        header = self.object_header([(0x11, struct.pack("<QQ", tree, heap_addr))], g.attrs)
        return header, tree, heap_addr

    def node(self, n) -> int:
        return self.dataset(n) if isinstance(n, Dataset) else self.group(n)[0]

    # ---- the file ------------------------------------------------------------------------------
    def global_heap(self) -> None:
        if not self.heap_objects:
            return
        body = b""
        for i, raw in enumerate(self.heap_objects, 1):
            body += struct.pack("<HHIQ", i, 1, 0, len(raw)) + pad8(raw)
        size = max(4096, 16 + len(body) + 16)
        size += -size % 8
        free = size - 16 - len(body)
        body += struct.pack("<HHIQ", 0, 0, 0, free) + b"\0" * (free - 16)
        addr = self.emit(b"GCOL" + struct.pack("<BBBBQ", 1, 0, 0, 0, size) + body)
        for at, n, idx in self.heap_refs:
            self.put(at, struct.pack("<IQI", n, addr, idx))

    def write(self, root: Group, path) -> None:
        sb_size = 48 if self.modern else 96
        assert self.alloc(sb_size) == 0
        header, tree, heap = self.group(root)
        self.global_heap()
        eof = len(self.buf) - self.base
        sig = b"\x89HDF\r\n\x1a\n"
        if self.modern:
            sb = sig + struct.pack("<BBBBQQQQI", 2, 8, 8, 0, self.base, UNDEF, eof, header, 0)
        else:
            sb = sig + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, self.leaf_k, self.node_k, 0)
            sb += struct.pack("<QQQQ", self.base, UNDEF, eof, UNDEF)
            sb += struct.pack("<QQII", 0, header, 1, 0) + struct.pack("<QQ", tree, heap)
        self.put(0, sb)
        with open(path, "wb") as fh:
            fh.write(self.buf)


def write_hdf5(root: Group, path, **options) -> None:
    Writer(**options).write(root, path)
