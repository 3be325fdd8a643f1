"""`OnDiskIndex` — drop-in for src/fast_forward/index/disk.py:25-418.

The HDF5 file keeps the reference's layout (root attrs `num_vectors`, `ff_version`; datasets
`vectors` (capacity, dim) chunked (chunk_size, dim), `doc_ids` / `psg_ids` as fixed-width
bytes with "" = absent; group `quantizer/{meta,attributes,data}`), so files are
interchangeable.  The difference is the read path: instead of fancy-indexing the file on
every call, the rows are staged ONCE — HDF5 chunk by chunk through the pinned double buffer —
into an HBM row store, and all scoring runs there.

`load` (and everything a loaded index does afterwards except `add`) reads the file with the
library's own HDF5 reader (`fast_forward._h5`, csrc/ffx_h5.cpp): the file is mapped, every chunk
of `vectors` is handed to the staging buffers as a pointer into the mapping, and the two id
columns are coded by the C++ dictionaries straight from their fixed-width bytes.  Creating a
file and adding to it writes through `h5py` when it is installed, else through the package's
own writer of the same layout (`fast_forward._h5_write`).
"""

from __future__ import annotations

import logging
from collections.abc import Iterable, Iterator
from pathlib import Path

import numpy as np

import fast_forward
from fast_forward import _ffx, _h5, _h5_write
from fast_forward.encoder.base import Encoder
from fast_forward.index._store import RowStore, make_store
from fast_forward.index.base import IDSequence, Index, Mode
from fast_forward.index.memory import InMemoryIndex
from fast_forward.quantizer import Quantizer

LOGGER = logging.getLogger(__name__)


def _h5py():
    """The h5py module, or None when it is not installed (the native writer is used then)."""
    try:
        import h5py
    except ImportError:
        return None
    return h5py


def _text_ids(raw: np.ndarray):
    """A column of fixed-width byte ids (b"" = no id, disk.py:414-417) as an Arrow string array
    with nulls, built from the bytes without creating one Python object per id; without
    pyarrow, an object array of `str | None`."""
    raw = np.ascontiguousarray(raw, dtype=f"S{max(np.asarray(raw).dtype.itemsize, 1)}")
    try:
        import pyarrow as pa
    except ImportError:  # pragma: no cover - depends on the environment
        out = np.char.decode(raw, "utf-8").astype(object)
        out[out == ""] = None
        return out
    import ctypes as C

    n, width = len(raw), raw.dtype.itemsize
    offsets = np.empty(n + 1, np.int64)
    data = np.empty(max(n * width, 1), np.uint8)
    validity = np.empty(max((n + 7) // 8, 1), np.uint8)
    present = C.c_int64()
    _ffx.check(_ffx.lib().ffx_fixed_width_to_arrow(
        C.c_void_p(raw.ctypes.data), n, width, C.c_void_p(offsets.ctypes.data), C.c_void_p(data.ctypes.data),
        C.c_void_p(validity.ctypes.data), C.byref(present)))
    buffers = [pa.py_buffer(validity), pa.py_buffer(offsets), pa.py_buffer(data)]
    return pa.Array.from_buffers(pa.large_string(), n, buffers, null_count=n - present.value)


class OnDiskIndex(Index):
    """Fast-Forward index persisted in an HDF5 file and served from GPU memory.

    `memory_mapped` and `max_indexing_size` are accepted for compatibility; they tuned the
    reference's per-call file reads, which no longer happen.
    """

    def __init__(self, index_file: Path, query_encoder: Encoder | None = None,
                 quantizer: Quantizer | None = None, mode: Mode = Mode.MAXP,
                 encoder_batch_size: int = 32, init_size: int = 2**16, chunk_size: int = 2**16,
                 max_id_length: int = 8, overwrite: bool = False, memory_mapped: bool = False,
                 max_indexing_size: int = 2**10, device: int = 0, devices=None, shard: str = "query") -> None:
        """Create (or overwrite) an index file.  ValueError if it exists and `overwrite=False`.
        `devices` / `shard`: as `InMemoryIndex` (several GPUs driven by this process)."""
        if index_file.exists() and not overwrite:
            raise ValueError(f"File {index_file} exists.")
        self._index_file = index_file.absolute()
        self._store = make_store(device, devices, shard)
        self._init_size = init_size
        self._chunk_size = chunk_size
        self._max_id_length = max_id_length
        self._memory_mapped = memory_mapped
        self._max_indexing_size = max_indexing_size
        LOGGER.debug("creating file %s", self._index_file)
        h5py = _h5py()
        if h5py is None:
            _h5_write.IndexFile.create(self._index_file, fast_forward.__version__).close()
        else:
            with h5py.File(self._index_file, "w") as fp:
                fp.attrs["num_vectors"] = 0
                fp.attrs["ff_version"] = fast_forward.__version__
        super().__init__(query_encoder=query_encoder, quantizer=quantizer, mode=mode,
                         encoder_batch_size=encoder_batch_size)

    # ---- persistence ----------------------------------------------------------------------
    def _on_quantizer_set(self) -> None:
        meta, attributes, data = self.quantizer.serialize()
        h5py = _h5py()
        if h5py is None:
            with _h5_write.IndexFile.open_existing(self._index_file) as out:
                out.set_quantizer(meta, attributes, data)
            return
        with h5py.File(self._index_file, "a") as fp:
            if "quantizer" in fp:
                del fp["quantizer"]
            fp.create_group("quantizer/meta").attrs.update(meta)
            fp.create_group("quantizer/attributes").attrs.update(attributes)
            arrays = fp.create_group("quantizer/data")
            for key, value in data.items():
                arrays.create_dataset(key, data=value)

    def _create_datasets(self, fp, dim: int, dtype) -> None:
        id_type = f"S{self._max_id_length}"
        fp.create_dataset("vectors", (self._init_size, dim), dtype, maxshape=(None, dim),
                          chunks=(self._chunk_size, dim))
        for name in ("doc_ids", "psg_ids"):
            fp.create_dataset(name, (self._init_size,), id_type, maxshape=(None,), chunks=True)

    def _validate_ids(self, doc_ids: IDSequence, psg_ids: IDSequence, doc_width: int, psg_width: int) -> None:
        for d in doc_ids:
            if d is not None and len(d) > doc_width:
                raise RuntimeError(f"Document ID {d} is longer than the maximum ({doc_width} characters).")
        for p in psg_ids:
            if p is not None and len(p) > psg_width:
                raise RuntimeError(f"Passage ID {p} is longer than the maximum ({psg_width} characters).")
        self._store.check_new_passages(psg_ids)

    def _grown(self, have: int, extra: int) -> int:
        """Capacity after growing in whole HDF5 chunks (disk.py:268-276)."""
        return max(int((have + extra) / self._chunk_size + 0.5) * self._chunk_size, have + extra)

    def _add_native(self, vectors: np.ndarray, doc_ids: IDSequence, psg_ids: IDSequence) -> None:
        """`_add` through the package's own HDF5 writer (no h5py)."""
        with _h5_write.IndexFile.open_existing(self._index_file) as out:
            if not out.datasets:
                out.create_datasets(vectors.shape[-1], vectors.dtype, self._init_size, self._chunk_size,
                                    self._max_id_length)
            self._validate_ids(doc_ids, psg_ids, out.datasets["doc_ids"].dtype.itemsize,
                               out.datasets["psg_ids"].dtype.itemsize)
            have, extra = out.num_vectors, vectors.shape[0]
            assert have == self._store.count, "index file and device store are out of step"
            if extra > out.capacity - have:
                LOGGER.debug("resizing index from %s to %s", out.capacity, self._grown(have, extra))
                out.resize(self._grown(have, extra))
            for name, ids in (("doc_ids", doc_ids), ("psg_ids", psg_ids)):
                out.write_at(name, [have + i for i, v in enumerate(ids) if v is not None],
                             [v for v in ids if v is not None])
            out.write_rows("vectors", have, vectors)
            out.set_num_vectors(have + extra)  # bumped last: the file stays consistent

    def _add(self, vectors: np.ndarray, doc_ids: IDSequence, psg_ids: IDSequence) -> None:
        if self.quantizer is not None and vectors.dtype != np.uint8:
            raise NotImplementedError(
                f"Quantizer codes of type {vectors.dtype.name} (Ks > 256) cannot be scored on the device: only uint8 "
                "codes (Ks <= 256) are supported.")
        # device memory first: if the store cannot grow (out of HBM) nothing has been written, and the
        # file and the row store cannot drift apart
        rows_kind = np.uint8 if vectors.dtype == np.uint8 and self.quantizer is not None else np.float32
        self._store.reserve_for(vectors.shape[0], vectors.shape[-1], rows_kind, self._init_size, self._chunk_size)
        h5py = _h5py()
        if h5py is None:
            self._add_native(vectors, doc_ids, psg_ids)
            self._stage_added(vectors, doc_ids, psg_ids)
            return
        with h5py.File(self._index_file, "a") as fp:
            if "vectors" not in fp:
                self._create_datasets(fp, vectors.shape[-1], vectors.dtype)
            self._validate_ids(doc_ids, psg_ids, fp["doc_ids"].dtype.itemsize, fp["psg_ids"].dtype.itemsize)

            have = int(fp.attrs["num_vectors"])
            assert have == self._store.count, "index file and device store are out of step"
            extra = vectors.shape[0]
            if extra > fp["vectors"].shape[0] - have:
                grown = self._grown(have, extra)
                LOGGER.debug("resizing index from %s to %s", fp["vectors"].shape[0], grown)
                for name in ("vectors", "doc_ids", "psg_ids"):
                    fp[name].resize(grown, axis=0)

            for name, ids in (("doc_ids", doc_ids), ("psg_ids", psg_ids)):
                where = [have + i for i, v in enumerate(ids) if v is not None]
                if where:
                    fp[name][where] = [v for v in ids if v is not None]
            fp["vectors"][have:have + extra] = vectors
            fp.attrs["num_vectors"] = have + extra  # bumped last: the file stays consistent
        self._stage_added(vectors, doc_ids, psg_ids)

    def _stage_added(self, vectors: np.ndarray, doc_ids: IDSequence, psg_ids: IDSequence) -> None:
        rows = np.ascontiguousarray(vectors) if vectors.dtype == np.uint8 and self.quantizer is not None \
            else np.ascontiguousarray(vectors, dtype=np.float32)
        self._store.append(rows, doc_ids, psg_ids, self._init_size, self._chunk_size)

    # ---- Index contract -------------------------------------------------------------------
    def _get_num_vectors(self) -> int:
        return self._store.count  # == the file's `num_vectors`: this object is its only writer

    def _get_internal_dim(self) -> int | None:
        return self._store.width

    def _get_doc_ids(self) -> set[str]:
        return self._store.doc_id_set()

    def _get_psg_ids(self) -> set[str]:
        return self._store.psg_id_set()

    def _get_vectors(self, ids: Iterable[str]) -> tuple[np.ndarray, list[str]]:
        rows, owners = self._store.rows_for(ids, self.mode.name)
        return self._store.read(rows), owners

    def _batch_iter(self, batch_size: int) -> Iterator[tuple[np.ndarray, IDSequence, IDSequence]]:
        total = self._store.count
        for lo in range(0, total, batch_size):
            hi = min(lo + batch_size, total)
            doc_ids, psg_ids = self._store.id_columns(lo, hi)
            yield self._store.read(np.arange(lo, hi)), doc_ids, psg_ids

    def _device(self) -> _ffx.DeviceIndex:
        return self._store.device_index(self.quantizer)

    def _score(self, mode: Mode, qv, q_off, cand, lex=None, alpha=0.0, k=0, want_ff=True, out=None) -> dict:
        return self._store.score(self.quantizer, mode.value, qv, q_off, cand, lex, alpha, k, want_ff, out)

    def _early_stop(self, mode: Mode, qv, q_off, cand, lex, alpha, cutoff, depths) -> dict:
        return self._store.early_stop(self.quantizer, mode.value, qv, q_off, cand, lex, alpha, cutoff, depths)

    def _candidates(self, cols, mode: Mode) -> np.ndarray:
        return cols.candidates(self._store, mode == Mode.PASSAGE)

    def _resolve(self, ids, mode: Mode, missing_ok: bool = False) -> np.ndarray:
        return self._store.resolve(ids, mode == Mode.PASSAGE, missing_ok)

    # ---- conversions ------------------------------------------------------------------------
    def to_memory(self, batch_size: int | None = None) -> InMemoryIndex:
        """An `InMemoryIndex` with the same contents (disk.py:177-205)."""
        index = InMemoryIndex(query_encoder=self._query_encoder, quantizer=self._quantizer,
                              mode=self.mode, encoder_batch_size=self._encoder_batch_size,
                              init_size=max(len(self), 1), device=self._store.device,
                              devices=getattr(self._store, "devices", None),
                              shard="doc" if hasattr(self._store, "shards") else "query")
        copy = self._store.clone()
        if copy is not None:
            # the rows are in HBM already: one device-to-device copy and two dictionary copies instead of
            # reading every row back and adding it again
            index._store = copy
            return index
        for rows, doc_ids, psg_ids in self._batch_iter(batch_size or max(self._store.count, 1)):
            index._add(rows, doc_ids=doc_ids, psg_ids=psg_ids)
        return index

    @classmethod
    def load(cls, index_file: Path, query_encoder: Encoder | None = None, mode: Mode = Mode.MAXP,
             encoder_batch_size: int = 32, memory_mapped: bool = False,
             max_indexing_size: int = 2**10, device: int = 0, devices=None, shard: str = "query") -> "OnDiskIndex":
        """Open an existing index file and stage it into GPU memory (disk.py:355-418).
        `devices` / `shard`: as `InMemoryIndex` — `shard="query"` stages a replica on every device,
        `shard="doc"` spreads the documents over the devices (files larger than one GPU)."""
        LOGGER.debug("reading file %s", index_file)
        index = cls.__new__(cls)
        Index.__init__(index, query_encoder=query_encoder, quantizer=None, mode=mode,
                       encoder_batch_size=encoder_batch_size)
        index._index_file = index_file.absolute()
        index._store = make_store(device, devices, shard)
        index._memory_mapped = memory_mapped
        index._max_indexing_size = max_indexing_size

        try:
            cls._stage_native(index, index_file)
        except _ffx.FFXError as error:
            h5py = _h5py()
            if h5py is None:
                raise
            # a file the native reader does not cover (libver="latest" chunk indexes, filters, dense
            # link / attribute storage, ...): h5py can read whatever libhdf5 wrote
            LOGGER.warning("native HDF5 reader: %s; reading %s through h5py", error, index_file)
            index._quantizer = None
            index._store = make_store(device, devices, shard)
            cls._stage_h5py(index, h5py, index_file)
        return index

    @staticmethod
    def _check_code_dtype(index: "OnDiskIndex", stored_dtype) -> None:
        if index._quantizer is not None and index._quantizer.dtype != np.uint8:
            raise NotImplementedError(
                f"Quantizer codes of type {np.dtype(index._quantizer.dtype).name} (Ks > 256) cannot be scored on the "
                "device: only uint8 codes (Ks <= 256) are supported.")

    @classmethod
    def _stage_native(cls, index: "OnDiskIndex", index_file: Path) -> None:
        """File -> HBM through the library's own reader: every HDF5 chunk is a pointer into the
        mapped file, handed to the staging buffers as is."""
        with _h5.H5File(index_file) as fp:
            if "quantizer" in fp:
                index._quantizer = Quantizer.deserialize(
                    fp.attrs("quantizer/meta"), fp.attrs("quantizer/attributes"),
                    {k: fp.read(f"quantizer/data/{k}") for k in fp.keys("quantizer/data")})
            total = int(fp.attr("/", "num_vectors"))
            index._chunk_size = index._init_size = 2**16
            index._max_id_length = 8
            if "vectors" not in fp:
                return
            meta = fp.info("vectors")
            index._chunk_size = index._init_size = meta["chunk_rows"] or 2**16
            index._max_id_length = fp.info("doc_ids")["dtype"].itemsize
            if total == 0:
                return
            cls._check_code_dtype(index, meta["dtype"])

            # one HDF5 chunk (a contiguous byte range of the file) per staging step, read in place
            codes = index._quantizer is not None and meta["dtype"] == np.uint8
            by_ids = hasattr(index._store, "shards")  # doc shards place every row by its ids
            naming = None
            if by_ids:
                doc_col, psg_col = _text_ids(fp.read("doc_ids", 0, total)), _text_ids(fp.read("psg_ids", 0, total))
            else:
                # the O(N) Python loop of disk.py:408-417, as two calls into the C++ id dictionaries — on a
                # thread of their own, next to the row stream (both sides release the GIL)
                from concurrent.futures import ThreadPoolExecutor

                def name_rows():
                    with _h5.H5File(index_file) as ids_fp:  # a handle of its own: the reader's object cache is not shared
                        return RowStore.prepare_id_columns(_text_ids(ids_fp.read("doc_ids", 0, total)),
                                                           _text_ids(ids_fp.read("psg_ids", 0, total)), total)

                pool = ThreadPoolExecutor(1)
                naming = pool.submit(name_rows)
            try:
                for row0, block in fp.spans("vectors", 0, total):
                    rows = block if codes or block.dtype == np.float32 else block.astype(np.float32)
                    ids = (doc_col.slice(row0, len(rows)), psg_col.slice(row0, len(rows))) if by_ids else (None, None)
                    index._store.append(rows, ids[0], ids[1], first_capacity=total, grow_by=index._chunk_size)
                if naming is not None:
                    index._store.adopt_prepared(*naming.result())
            finally:
                if naming is not None:
                    pool.shutdown(wait=True)

    @classmethod
    def _stage_h5py(cls, index: "OnDiskIndex", h5py, index_file: Path) -> None:
        """File -> HBM through h5py, chunk by chunk (disk.py:380-417 of the reference with the
        per-row Python loop replaced by the C++ id dictionaries)."""
        with h5py.File(index_file, "r") as fp:
            if "quantizer" in fp:
                index._quantizer = Quantizer.deserialize(
                    dict(fp["quantizer/meta"].attrs), dict(fp["quantizer/attributes"].attrs),
                    {k: v[:] for k, v in fp["quantizer/data"].items()})
            total = int(fp.attrs["num_vectors"])
            index._chunk_size = index._init_size = 2**16
            index._max_id_length = 8
            if "vectors" not in fp:
                return
            vectors = fp["vectors"]
            index._chunk_size = index._init_size = (vectors.chunks[0] if vectors.chunks else 0) or 2**16
            index._max_id_length = fp["doc_ids"].dtype.itemsize
            if total == 0:
                return
            cls._check_code_dtype(index, vectors.dtype)
            codes = index._quantizer is not None and vectors.dtype == np.uint8
            for lo in range(0, total, index._chunk_size):
                hi = min(total, lo + index._chunk_size)
                block = vectors[lo:hi]
                rows = block if codes or block.dtype == np.float32 else block.astype(np.float32)
                index._store.append(np.ascontiguousarray(rows), _text_ids(fp["doc_ids"][lo:hi]),
                                    _text_ids(fp["psg_ids"][lo:hi]), first_capacity=total, grow_by=index._chunk_size)
