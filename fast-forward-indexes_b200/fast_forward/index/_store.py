"""Host id maps + the HBM row store shared by the index back-ends.

The id-mapping contract is the reference's (index/util.py:12-42, index/memory.py:84-95):
`doc_rows[id]` lists a document's rows in insertion order, `psg_row[id]` is a passage's single
row, an id that resolves to no row raises IndexError.  The maps stay on the host (id strings
never cross the C ABI); what goes to the device is their integer form — a CSR
document-ordinal -> rows table (ffx_index_set_docs) — plus the rows themselves.
"""

from __future__ import annotations

import itertools
from collections import defaultdict
from collections.abc import Iterable, Sequence

import numpy as np
import pandas as pd

from fast_forward import _ffx


class RowStore:
    """Rows in HBM (fp32 vectors or uint8 PQ codes) + the two id maps."""

    def __init__(self, device: int = 0) -> None:
        self.device = device
        self.dev: _ffx.DeviceIndex | None = None
        self.count = 0
        self.doc_rows: dict[str, list[int]] = defaultdict(list)
        self.psg_row: dict[str, int] = {}
        self._maps_stale = True
        self._doc_lookup: pd.Index | None = None
        self._psg_lookup: pd.Index | None = None
        self._psg_rows: np.ndarray | None = None
        self._pq_of: object | None = None  # quantizer whose tables are on the device

    # ---- properties -------------------------------------------------------------------
    @property
    def width(self) -> int | None:
        """Elements per stored row (vector dimension, or M for codes)."""
        return None if self.dev is None else self.dev.dim

    def check_new_passages(self, psg_ids: Iterable[str | None]) -> None:
        seen = set()
        for p in psg_ids:
            if p is None:
                continue
            if p in self.psg_row or p in seen:
                raise RuntimeError(f"Passage ID {p} already exists.")
            seen.add(p)

    # ---- growth -------------------------------------------------------------------------
    def append(self, rows: np.ndarray, doc_ids, psg_ids, first_capacity: int, grow_by: int) -> None:
        """Stage `rows` at the end of the store and record their ids."""
        n_new = rows.shape[0]
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
        for i, d in enumerate(doc_ids, self.count):
            if d is not None:
                self.doc_rows[d].append(i)
        for i, p in enumerate(psg_ids, self.count):
            if p is not None:
                self.psg_row[p] = i
        self.count = need
        self._maps_stale = True

    def record_ids(self, doc_ids: Sequence[str | None], psg_ids: Sequence[str | None], base: int) -> None:
        """Register the ids of rows [base, base+len) (used by bulk loaders)."""
        for i, d in enumerate(doc_ids, base):
            if d is not None:
                self.doc_rows[d].append(i)
        for i, p in enumerate(psg_ids, base):
            if p is not None:
                self.psg_row[p] = i
        self._maps_stale = True

    # ---- id mapping ---------------------------------------------------------------------
    def _refresh(self) -> None:
        if not self._maps_stale:
            return
        docs = self.doc_rows
        self._doc_lookup = pd.Index(list(docs.keys()), dtype=object)
        if self.dev is not None:
            lengths = np.fromiter((len(v) for v in docs.values()), np.int64, len(docs))
            off = np.zeros(len(docs) + 1, np.int64)
            np.cumsum(lengths, out=off[1:])
            flat = np.fromiter(itertools.chain.from_iterable(docs.values()), np.int64, int(off[-1]))
            self.dev.set_docs(off, flat)
        self._psg_lookup = pd.Index(list(self.psg_row.keys()), dtype=object)
        self._psg_rows = np.fromiter(self.psg_row.values(), np.int64, len(self.psg_row))
        self._maps_stale = False

    def resolve(self, ids: np.ndarray, passage_mode: bool) -> np.ndarray:
        """Unique ids -> int32 document ordinals (or row numbers in PASSAGE mode)."""
        self._refresh()
        lookup = self._psg_lookup if passage_mode else self._doc_lookup
        where = lookup.get_indexer(pd.Index(ids, dtype=object)) if len(ids) else np.zeros(0, np.int64)
        missing = np.flatnonzero(where < 0)
        if len(missing):
            raise IndexError(f"ID {ids[missing[0]]} not found in the index.")
        if passage_mode:
            where = self._psg_rows[where]
        return where.astype(np.int32)

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
        if getattr(self, "_reverse_for", None) != self.count:
            doc_of = np.full(self.count, None, dtype=object)
            for d, rows in self.doc_rows.items():
                doc_of[rows] = d
            psg_of = np.full(self.count, None, dtype=object)
            if self.psg_row:
                psg_of[np.fromiter(self.psg_row.values(), np.int64, len(self.psg_row))] = list(self.psg_row.keys())
            self._doc_of, self._psg_of, self._reverse_for = doc_of, psg_of, self.count
        return self._doc_of[lo:hi].tolist(), self._psg_of[lo:hi].tolist()
