"""Host id maps + the HBM row store shared by the index back-ends.

The id-mapping contract is the reference's (index/util.py:12-42, index/memory.py:84-95): a
document id owns its rows in insertion order, a passage id owns exactly one row, an id that
resolves to no row raises IndexError.  The maps live in two C++ dictionaries of libffx
(`fast_forward._ids.IdDict`: document id -> document ordinal in order of first appearance,
passage id -> row number) plus one integer per row (its document ordinal); id strings never
cross to the device — what goes there is the integer form, a CSR document-ordinal -> rows table
(ffx_index_set_docs), plus the rows themselves.
"""

from __future__ import annotations

import itertools
from collections.abc import Iterable, Sequence

import numpy as np

from fast_forward import _ffx, _ids


def _take(keys, index: np.ndarray) -> list:
    """keys[index] as a Python list, None where index < 0 (`keys`: pyarrow array or list)."""
    if isinstance(keys, list):
        return [None if i < 0 else keys[i] for i in index.tolist()]
    picks = _ids.pa.array(np.where(index < 0, 0, index), mask=index < 0)
    return keys.take(picks).to_pylist()


_TOKENS = itertools.count(1)


class RowStore:
    """Rows in HBM (fp32 vectors or uint8 PQ codes) + the two id maps."""

    def __init__(self, device: int = 0) -> None:
        self.device = device
        self.token = next(_TOKENS)  # identity of this store in the candidate caches of rankings
        self.version = 0            # bumped whenever ids are added (cached candidates go stale)
        self.dev: _ffx.DeviceIndex | None = None
        self.count = 0
        self.docs = _ids.IdDict()   # document id -> ordinal (order of first appearance)
        self.psgs = _ids.IdDict()   # passage id -> row
        self._row_doc_parts: list[np.ndarray] = []  # document ordinal of every row (-1 = none)
        self._maps_stale = True
        self._doc_off: np.ndarray | None = None   # CSR ordinal -> rows (host copy, for _get_vectors)
        self._doc_rows: np.ndarray | None = None
        self._reverse = None  # (count, doc keys, psg keys, row -> psg key)
        self._pq_of: object | None = None  # quantizer whose tables are on the device

    # ---- properties -------------------------------------------------------------------
    @property
    def width(self) -> int | None:
        """Elements per stored row (vector dimension, or M for codes)."""
        return None if self.dev is None else self.dev.dim

    def doc_id_set(self) -> set[str]:
        return set(self.docs.keys())

    def psg_id_set(self) -> set[str]:
        return set(self.psgs.keys())

    def check_new_passages(self, psg_ids: Sequence[str | None]) -> None:
        """RuntimeError if a passage id exists already or repeats in the batch (memory.py:93-94)."""
        psg_ids = _ids.as_id_list(psg_ids)
        dup = self.psgs.insert_unique(psg_ids, self.count, dry_run=True) if psg_ids else -1
        if dup >= 0:
            raise RuntimeError(f"Passage ID {psg_ids[dup]} already exists.")

    # ---- growth -------------------------------------------------------------------------
    def append(self, rows: np.ndarray, doc_ids, psg_ids, first_capacity: int, grow_by: int) -> None:
        """Stage `rows` at the end of the store and record their ids (`None` for both id
        arguments: a bulk loader names the rows afterwards with `adopt_id_columns`)."""
        n_new = rows.shape[0]
        if psg_ids is not None:
            self.check_new_passages(psg_ids)
        if self.dev is None:
            kind = _ffx.ROWS_PQ_U8 if rows.dtype == np.uint8 else _ffx.ROWS_F32
            self.dev = _ffx.DeviceIndex(rows.shape[1], capacity=max(first_capacity, n_new),
                                        row_kind=kind, device=self.device)
        need = self.count + n_new
        if need > self.dev.capacity:
            # whole `grow_by` chunks like the reference, but at least 1.5x so that repeated
            # appends stay linear (each growth copies the store inside HBM)
            chunks = -(-(need - self.dev.capacity) // max(grow_by, 1))
            self.dev.reserve(max(self.dev.capacity + chunks * max(grow_by, 1),
                                 int(self.dev.capacity * 1.5)))
        self.dev.stage(self.count, rows)
        if doc_ids is not None or psg_ids is not None:
            self.record_ids(doc_ids, psg_ids, self.count, n_new)
        self.count = need
        self._maps_stale = True

    def record_ids(self, doc_ids, psg_ids, base: int, n: int) -> None:
        """Register the ids of rows [base, base+n); either sequence may hold None entries."""
        if doc_ids is not None and len(doc_ids):
            self._row_doc_parts.append(self.docs.insert_ordinal(doc_ids))
        else:
            self._row_doc_parts.append(np.full(n, -1, np.int64))
        if psg_ids is not None and len(psg_ids):
            psg_ids = _ids.as_id_list(psg_ids)
            dup = self.psgs.insert_unique(psg_ids, base)
            if dup >= 0:
                raise RuntimeError(f"Passage ID {psg_ids[dup]} already exists.")
        self._maps_stale = True
        self.version += 1

    def adopt_id_columns(self, doc_col, psg_col) -> None:
        """Name all `count` rows at once from two id columns (None = no id): the vectorised
        replacement for the O(N) Python loop of index/disk.py:408-417."""
        self.docs, self.psgs, self._row_doc_parts = _ids.IdDict(), _ids.IdDict(), []
        self.record_ids(doc_col, psg_col, 0, self.count)

    # ---- id mapping ---------------------------------------------------------------------
    def _row_doc(self) -> np.ndarray:
        if len(self._row_doc_parts) != 1:
            merged = np.concatenate(self._row_doc_parts) if self._row_doc_parts else np.zeros(0, np.int64)
            self._row_doc_parts = [merged]
        return self._row_doc_parts[0]

    def _refresh(self) -> None:
        if not self._maps_stale:
            return
        self._doc_off, self._doc_rows = _ids.csr_from_ordinals(self._row_doc(), len(self.docs))
        if self.dev is not None:
            self.dev.set_docs(self._doc_off, self._doc_rows)
        self._maps_stale = False

    def resolve(self, ids, passage_mode: bool) -> np.ndarray:
        """An id column (one entry per pair, repeats welcome) -> int32 candidates for ffx_rerank:
        document ordinals, or row numbers in PASSAGE mode.  IndexError names the first id that
        is not in the index (index/util.py:38-39)."""
        self._refresh()
        codes, missing = (self.psgs if passage_mode else self.docs).lookup(ids)
        if missing >= 0:
            raise IndexError(f"ID {_ids.first_text(ids, missing)} not found in the index.")
        return codes

    def lookup_keys(self, keys, passage_mode: bool) -> np.ndarray:
        """int32 candidate of every DISTINCT id of a ranking (`fast_forward._cols.IdTable`), -1 where
        the index does not hold it — index/util.py:29-41 once per id instead of once per pair."""
        self._refresh()
        codes, _ = (self.psgs if passage_mode else self.docs).lookup(keys)
        return codes

    def rows_for(self, ids: Iterable[str], mode_name: str) -> tuple[np.ndarray, list[str]]:
        """index/util.py:12-42 (`get_indices`): the rows needed to score `ids` in the given mode
        (all rows of a document, its first row, or a passage's row) and the owning id of each."""
        ids = list(ids)
        if not ids:
            return np.zeros(0, np.int64), []
        codes = self.resolve(ids, mode_name == "PASSAGE").astype(np.int64)
        if mode_name == "PASSAGE":
            return codes & 0xffffffff, ids
        first = self._doc_off[codes]
        if mode_name == "FIRSTP":
            return self._doc_rows[first], ids
        counts = self._doc_off[codes + 1] - first
        within = np.arange(int(counts.sum())) - np.repeat(np.cumsum(counts) - counts, counts)
        return self._doc_rows[np.repeat(first, counts) + within], np.repeat(np.asarray(ids, object), counts).tolist()

    # ---- device ---------------------------------------------------------------------------
    def device_index(self, quantizer=None) -> _ffx.DeviceIndex:
        """The libffx index with the doc table (and PQ tables) up to date."""
        if self.dev is None:
            raise IndexError("The index is empty.")
        self._refresh()
        if quantizer is not None and self._pq_of is not quantizer:
            tables = quantizer.adc_tables()
            if tables is None:
                raise RuntimeError(f"{type(quantizer).__name__} cannot be scored from its codes.")
            self.dev.set_pq(tables[0], tables[1])
            self._pq_of = quantizer
        return self.dev

    def read(self, rows) -> np.ndarray:
        if self.dev is None or len(rows) == 0:
            return np.array([])
        return self.dev.read_rows(rows)

    def id_columns(self, lo: int, hi: int):
        """(doc ids, passage ids) of rows [lo, hi), None where a row has no such id."""
        if self._reverse is None or self._reverse[0] != self.count:
            doc_keys, _ = self.docs.export()
            psg_keys, psg_rows = self.psgs.export()
            row_psg = np.full(self.count, -1, np.int64)
            row_psg[psg_rows] = np.arange(len(psg_rows))
            self._reverse = (self.count, doc_keys, psg_keys, row_psg)
        _, doc_keys, psg_keys, row_psg = self._reverse
        return _take(doc_keys, self._row_doc()[lo:hi]), _take(psg_keys, row_psg[lo:hi])
