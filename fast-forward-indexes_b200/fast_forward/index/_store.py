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
import os
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

    def clone(self) -> "RowStore | None":
        """A store with the same rows and ids, copied device to device (OnDiskIndex.to_memory: the
        on-disk index is resident in HBM already); None where only the generic row-by-row route
        applies (the multi-device stores)."""
        if type(self) is not RowStore:
            return None
        other = RowStore(self.device)
        if self.dev is not None:
            other.dev = _ffx.DeviceIndex(self.dev.dim, capacity=max(self.count, 1), row_kind=self.dev.row_kind,
                                         device=self.device)
            other.dev.copy_rows_from(self.dev)
        other.count = self.count
        other.docs, other.psgs = self.docs.clone(), self.psgs.clone()
        other._row_doc_parts = [self._row_doc().copy()]
        return other

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
        if not hasattr(psg_ids, "__getitem__"):
            psg_ids = _ids.as_id_list(psg_ids)
        dup = self.psgs.insert_unique(psg_ids, self.count, dry_run=True) if len(psg_ids) else -1
        if dup >= 0:
            raise RuntimeError(f"Passage ID {_ids.first_text(psg_ids, dup)} already exists.")

    # ---- growth -------------------------------------------------------------------------
    def reserve_for(self, n_new: int, width: int, dtype, first_capacity: int, grow_by: int) -> None:
        """Device memory for `n_new` more rows (creates the store on first use).  The only step of
        an append that can fail for lack of HBM — callers that also persist the rows do it first."""
        if self.dev is None:
            kind = _ffx.ROWS_PQ_U8 if np.dtype(dtype) == np.uint8 else _ffx.ROWS_F32
            self.dev = _ffx.DeviceIndex(width, capacity=max(first_capacity, n_new), row_kind=kind, device=self.device)
        need = self.count + n_new
        if need > self.dev.capacity:
            # whole `grow_by` chunks like the reference, but at least 1.5x so that repeated
            # appends stay linear (each growth copies the store inside HBM)
            chunks = -(-(need - self.dev.capacity) // max(grow_by, 1))
            self.dev.reserve(max(self.dev.capacity + chunks * max(grow_by, 1),
                                 int(self.dev.capacity * 1.5)))

    def append(self, rows: np.ndarray, doc_ids, psg_ids, first_capacity: int, grow_by: int) -> None:
        """Stage `rows` at the end of the store and record their ids (`None` for both id
        arguments: a bulk loader names the rows afterwards with `adopt_id_columns`)."""
        n_new = rows.shape[0]
        if psg_ids is not None:
            self.check_new_passages(psg_ids)
        self.reserve_for(n_new, rows.shape[1], rows.dtype, first_capacity, grow_by)
        need = self.count + n_new
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
            if not hasattr(psg_ids, "__getitem__"):  # a generator: the error message indexes it
                psg_ids = _ids.as_id_list(psg_ids)
            dup = self.psgs.insert_unique(psg_ids, base)  # columns stay columns (no Python list of a million ids)
            if dup >= 0:
                raise RuntimeError(f"Passage ID {_ids.first_text(psg_ids, dup)} already exists.")
        self._maps_stale = True
        self.version += 1

    def adopt_id_columns(self, doc_col, psg_col) -> None:
        """Name all `count` rows at once from two id columns (None = no id): the vectorised
        replacement for the O(N) Python loop of index/disk.py:408-417."""
        self.adopt_prepared(*self.prepare_id_columns(doc_col, psg_col, self.count))

    @staticmethod
    def prepare_id_columns(doc_col, psg_col, n: int):
        """The dictionaries and the row -> document table of `n` rows named by two id columns.
        Touches no store: a loader runs it on a thread of its own while the rows stream to the
        device (the C++ dictionary calls release the GIL)."""
        docs, psgs = _ids.IdDict(), _ids.IdDict()
        row_doc = docs.insert_ordinal(doc_col) if doc_col is not None and len(doc_col) else np.full(n, -1, np.int64)
        if psg_col is not None and len(psg_col):
            dup = psgs.insert_unique(psg_col, 0)
            if dup >= 0:
                raise RuntimeError(f"Passage ID {_ids.first_text(psg_col, dup)} already exists.")
        return docs, psgs, row_doc

    def adopt_prepared(self, docs, psgs, row_doc) -> None:
        self.docs, self.psgs, self._row_doc_parts = docs, psgs, [row_doc]
        self._maps_stale = True
        self.version += 1

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
        self._push_maps()
        self._maps_stale = False

    def _push_maps(self) -> None:
        """The doc -> rows table to the device(s)."""
        if self.dev is not None:
            self.dev.set_docs(self._doc_off, self._doc_rows)

    def _encode(self, codes: np.ndarray, passage_mode: bool) -> np.ndarray:
        """Document ordinals / row numbers (-1 = unknown) -> the candidates ffx_rerank takes; the
        identity on one device, `device * stride + local` on a doc-sharded store."""
        return codes

    def resolve(self, ids, passage_mode: bool, missing_ok: bool = False) -> np.ndarray:
        """An id column (one entry per pair, repeats welcome) -> int32 candidates for ffx_rerank:
        document ordinals, or row numbers in PASSAGE mode.  IndexError names the first id that
        is not in the index (index/util.py:38-39)."""
        return self._encode(self._ordinals(ids, passage_mode, missing_ok), passage_mode)

    def _ordinals(self, ids, passage_mode: bool, missing_ok: bool = False) -> np.ndarray:
        self._refresh()
        codes, missing = (self.psgs if passage_mode else self.docs).lookup(ids)
        if missing >= 0 and not missing_ok:
            raise IndexError(f"ID {_ids.first_text(ids, missing)} not found in the index.")
        return codes

    def lookup_keys(self, keys, passage_mode: bool) -> np.ndarray:
        """int32 candidate of every DISTINCT id of a ranking (`fast_forward._cols.IdTable`), -1 where
        the index does not hold it — index/util.py:29-41 once per id instead of once per pair."""
        self._refresh()
        codes, _ = (self.psgs if passage_mode else self.docs).lookup(keys)
        return self._encode(codes, passage_mode)

    def rows_for(self, ids: Iterable[str], mode_name: str) -> tuple[np.ndarray, list[str]]:
        """index/util.py:12-42 (`get_indices`): the rows needed to score `ids` in the given mode
        (all rows of a document, its first row, or a passage's row) and the owning id of each."""
        ids = list(ids)
        if not ids:
            return np.zeros(0, np.int64), []
        codes = self._ordinals(ids, mode_name == "PASSAGE").astype(np.int64)
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

    # ---- scoring --------------------------------------------------------------------------
    def score(self, quantizer, mode: int, qv, q_off, cand, lex=None, alpha: float = 0.0, k: int = 0,
              want_ff: bool = True, out: dict | None = None) -> dict:
        """ffx_rerank_host over this store (one device here; the multi-device stores split the
        queries over their replicas, or the candidates over their doc-id-range shards)."""
        return self.device_index(quantizer).rerank_host(mode, qv, q_off, cand, lex, alpha, k, want_ff=want_ff,
                                                        want_int=False, out=out)

    def early_stop(self, quantizer, mode: int, qv, q_off, cand, lex, alpha, cutoff, depths) -> dict:
        """ffx_rerank_early_stop_host (FFXError -5 where the device walk does not apply)."""
        return self.device_index(quantizer).rerank_early_stop_host(mode, qv, q_off, cand, lex, alpha, cutoff, depths)

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


def make_store(device: int = 0, devices=None, shard: str = "query") -> RowStore:
    """The row store behind an index: one device, replicas on several devices (queries split
    across them) or doc-id-range shards on several devices (corpora larger than one GPU)."""
    if shard not in ("query", "doc"):
        raise ValueError(f'shard must be "query" or "doc", not {shard!r}')
    if devices is None or len(list(devices)) == 0:
        return RowStore(device)
    devices = [int(d) for d in devices]
    if len(set(devices)) != len(devices) and not os.environ.get("FFX_ALLOW_DUPLICATE_DEVICES"):
        raise ValueError("devices must be distinct")  # (the test-suite lifts this to run on one GPU)
    if len(devices) == 1:
        return RowStore(devices[0])
    return ReplicatedStore(devices) if shard == "query" else DocShardedStore(devices)


def _split_queries(q_off: np.ndarray, parts: int) -> list[int]:
    """Contiguous query ranges with (nearly) equal pair counts: boundaries [parts + 1]."""
    nq = len(q_off) - 1
    targets = q_off[-1] * np.arange(1, parts) / parts
    inner = np.searchsorted(q_off, targets, side="left")
    return np.maximum.accumulate(np.concatenate([[0], np.minimum(inner, nq), [nq]])).astype(int).tolist()


class ReplicatedStore(RowStore):
    """The same rows on several devices of this process (the index fits one GPU, the
    QUERIES shard).  Every `score` call cuts the query blocks into one contiguous range per
    device — balanced by pair count — and drives the devices from one host thread each (the
    C-ABI calls release the GIL); no collective is involved, results are those of one device."""

    def __init__(self, devices: list[int]) -> None:
        super().__init__(devices[0])
        from concurrent.futures import ThreadPoolExecutor

        self.devices = list(devices)
        self.replicas: list[_ffx.DeviceIndex] = []  # devices[1:]
        self._pool = ThreadPoolExecutor(len(devices))
        self._replica_pq = None

    def all_devices(self) -> list[_ffx.DeviceIndex]:
        return [self.dev, *self.replicas]

    def reserve_for(self, n_new: int, width: int, dtype, first_capacity: int, grow_by: int) -> None:
        super().reserve_for(n_new, width, dtype, first_capacity, grow_by)
        if not self.replicas:
            self.replicas = [_ffx.DeviceIndex(self.dev.dim, capacity=self.dev.capacity, row_kind=self.dev.row_kind, device=d)
                             for d in self.devices[1:]]
        for replica in self.replicas:
            if replica.capacity < self.dev.capacity:
                replica.reserve(self.dev.capacity)

    def append(self, rows: np.ndarray, doc_ids, psg_ids, first_capacity: int, grow_by: int) -> None:
        start = self.count
        super().append(rows, doc_ids, psg_ids, first_capacity, grow_by)
        list(self._pool.map(lambda replica: replica.stage(start, rows), self.replicas))

    def _push_maps(self) -> None:
        super()._push_maps()
        for replica in self.replicas:
            replica.set_docs(self._doc_off, self._doc_rows)

    def device_index(self, quantizer=None) -> _ffx.DeviceIndex:
        dev = super().device_index(quantizer)
        if quantizer is not None and self._replica_pq is not quantizer:
            tables = quantizer.adc_tables()
            for replica in self.replicas:
                replica.set_pq(tables[0], tables[1])
            self._replica_pq = quantizer
        return dev

    def _fan_out(self, q_off, run):
        """run(device index, lo, hi, r0, r1) on every device's share of the queries, in parallel."""
        devs = self.all_devices()
        bounds = _split_queries(q_off, len(devs))
        jobs = [(dev, bounds[i], bounds[i + 1]) for i, dev in enumerate(devs) if bounds[i + 1] > bounds[i]]
        return list(self._pool.map(lambda j: run(j[0], j[1], j[2], int(q_off[j[1]]), int(q_off[j[2]])), jobs))

    def score(self, quantizer, mode, qv, q_off, cand, lex=None, alpha=0.0, k=0, want_ff=True, out=None) -> dict:
        self.device_index(quantizer)
        q_off = np.ascontiguousarray(q_off, np.int64)
        nq, n = len(q_off) - 1, int(q_off[-1])
        out = {} if out is None else out
        if want_ff and out.get("ff") is None:
            out["ff"] = _ffx.pinned_empty(n, np.float32)
        if k > 0 and out.get("topk_score") is None:
            out["topk_score"] = _ffx.pinned_empty((nq, k), np.float32)
            out["topk_pos"] = _ffx.pinned_empty((nq, k), np.int32)

        def run(dev, lo, hi, r0, r1):
            part = {}
            if want_ff:
                part["ff"] = out["ff"][r0:r1]
            if k > 0:
                part["topk_score"], part["topk_pos"] = out["topk_score"][lo:hi], out["topk_pos"][lo:hi]
            dev.rerank_host(mode, qv[lo:hi], q_off[lo:hi + 1] - r0, cand[r0:r1], None if lex is None else lex[r0:r1],
                            alpha, k, want_ff=want_ff, want_int=False, out=part)

        self._fan_out(q_off, run)
        return out

    def early_stop(self, quantizer, mode, qv, q_off, cand, lex, alpha, cutoff, depths) -> dict:
        self.device_index(quantizer)
        q_off = np.ascontiguousarray(q_off, np.int64)
        out = {"ff": np.zeros(int(q_off[-1]), np.float32), "scored": np.zeros(len(q_off) - 1, np.int32)}

        def run(dev, lo, hi, r0, r1):
            part = dev.rerank_early_stop_host(mode, qv[lo:hi], q_off[lo:hi + 1] - r0, cand[r0:r1], lex[r0:r1], alpha,
                                              cutoff, depths)
            out["ff"][r0:r1], out["scored"][lo:hi] = part["ff"], part["scored"]

        self._fan_out(q_off, run)
        return out


class DocShardedStore(RowStore):
    """Doc-id-range shards on several devices of this process (a corpus larger than one
    GPU).  A document lives on ONE device: its first row picks the least-loaded device, later rows
    (documents may be extended by later `add` calls, reference tests/test_index.py:58-69) follow
    it; rows without a document id are spread the same way.  The candidate the kernels take is
    `device * stride + local ordinal` — every device is a shard `[d * stride, d * stride + #docs
    there)` of one global ordinal space (ffx_index_set_shard), skips the pairs it does not own and
    ranks its own at their positions in the full block; the per-device lists are merged with
    ffx_merge_topk.  Results are those of one device holding everything."""

    def __init__(self, devices: list[int]) -> None:
        super().__init__(devices[0])
        from concurrent.futures import ThreadPoolExecutor

        self.devices = list(devices)
        n = len(devices)
        self.stride = ((1 << 31) - 1) // n  # device * stride + local stays a non-negative int32
        self.shards: list[_ffx.DeviceIndex | None] = [None] * n
        self.loads = np.zeros(n, np.int64)       # rows per shard
        self.local_docs = np.zeros(n, np.int64)  # documents per shard
        self._doc_enc = np.zeros(0, np.int32)    # global document ordinal -> candidate
        self._row_enc_parts: list[np.ndarray] = []  # global row -> device * stride + local row
        self._local_row_doc: list[list[np.ndarray]] = [[] for _ in range(n)]  # per shard: local doc of each local row
        self._pool = ThreadPoolExecutor(n)
        self._shard_pq = None

    # ---- growth -------------------------------------------------------------------------
    def _place(self, weights: np.ndarray, loads: np.ndarray) -> np.ndarray:
        """Device of every new unit (documents in order of first appearance, then loose rows):
        contiguous runs, sized so that the loads even out."""
        n = len(self.devices)
        if len(weights) == 0:
            return np.zeros(0, np.int64)
        target = (loads.sum() + weights.sum()) / n
        room = np.maximum(target - loads, 0.0)
        if room.sum() <= 0:
            room[:] = 1.0
        edges = np.cumsum(room) / room.sum() * weights.sum()
        mid = np.cumsum(weights) - weights / 2.0
        return np.minimum(np.searchsorted(edges, mid, side="right"), n - 1)

    def append(self, rows: np.ndarray, doc_ids, psg_ids, first_capacity: int, grow_by: int) -> None:
        n_new = rows.shape[0]
        if doc_ids is None and psg_ids is None:
            raise RuntimeError("a doc-sharded store places rows by their ids: pass them with the rows")
        if psg_ids is not None:
            self.check_new_passages(psg_ids)
        had_docs = len(self.docs)
        self.record_ids(doc_ids, psg_ids, self.count, n_new)
        ordn = self._row_doc_parts[-1]
        fresh = len(self.docs) - had_docs
        # rows of known documents follow them; new documents and loose rows are placed
        known = (ordn >= 0) & (ordn < had_docs)
        loads = self.loads.astype(np.float64)
        if known.any():
            loads += np.bincount(self._doc_enc[ordn[known]] // self.stride, minlength=len(self.devices))
        per_new_doc = np.bincount(ordn[ordn >= had_docs] - had_docs, minlength=fresh)
        loose = np.flatnonzero(ordn < 0)
        unit_dev = self._place(np.concatenate([per_new_doc, np.ones(len(loose))]).astype(np.float64), loads)
        new_dev = unit_dev[:fresh]
        # local ordinals of the new documents, per device in order of appearance
        new_local = np.empty(fresh, np.int64)
        for d in range(len(self.devices)):
            sel = np.flatnonzero(new_dev == d)
            new_local[sel] = self.local_docs[d] + np.arange(len(sel))
            self.local_docs[d] += len(sel)
        self._doc_enc = np.concatenate([self._doc_enc, (new_dev * self.stride + new_local).astype(np.int32)])
        row_dev = np.empty(n_new, np.int64)
        has_doc = ordn >= 0
        row_dev[has_doc] = self._doc_enc[ordn[has_doc]] // self.stride
        row_dev[loose] = unit_dev[fresh:]
        row_enc = np.empty(n_new, np.int32)
        kind = _ffx.ROWS_PQ_U8 if rows.dtype == np.uint8 else _ffx.ROWS_F32
        share = max(1, -(-max(first_capacity, n_new) // len(self.devices)))
        for d, device in enumerate(self.devices):
            sel = np.flatnonzero(row_dev == d)
            if len(sel) == 0:
                continue
            if self.shards[d] is None:
                self.shards[d] = _ffx.DeviceIndex(rows.shape[1], capacity=max(share + share // 4, len(sel)), row_kind=kind,
                                                  device=device)
                if self.dev is None:
                    self.dev = self.shards[d]
            shard, at = self.shards[d], int(self.loads[d])
            if at + len(sel) > shard.capacity:
                shard.reserve(max(at + len(sel), int(shard.capacity * 1.5), shard.capacity + max(grow_by, 1)))
            shard.stage(at, rows[sel])
            row_enc[sel] = d * self.stride + at + np.arange(len(sel))
            local_doc = np.full(len(sel), -1, np.int64)
            with_doc = has_doc[sel]
            local_doc[with_doc] = self._doc_enc[ordn[sel][with_doc]] % self.stride
            self._local_row_doc[d].append(local_doc)
            self.loads[d] += len(sel)
        self._row_enc_parts.append(row_enc)
        self.count += n_new
        self._maps_stale = True

    def reserve_for(self, n_new: int, width: int, dtype, first_capacity: int, grow_by: int) -> None:
        """Nothing to reserve ahead: which shard grows depends on the ids of the rows."""

    def adopt_id_columns(self, doc_col, psg_col) -> None:
        raise RuntimeError("a doc-sharded store places rows by their ids: pass them with the rows")

    def _row_enc(self) -> np.ndarray:
        if len(self._row_enc_parts) != 1:
            merged = np.concatenate(self._row_enc_parts) if self._row_enc_parts else np.zeros(0, np.int32)
            self._row_enc_parts = [merged]
        return self._row_enc_parts[0]

    # ---- id mapping ---------------------------------------------------------------------
    def _push_maps(self) -> None:
        n = len(self.devices)
        for d, shard in enumerate(self.shards):
            if shard is None:
                continue
            local = np.concatenate(self._local_row_doc[d])
            self._local_row_doc[d] = [local]
            off, rows = _ids.csr_from_ordinals(local, int(self.local_docs[d]))
            shard.set_docs(off, rows)
            shard.set_shard(d * self.stride, n * self.stride, d * self.stride, n * self.stride)

    def _encode(self, codes: np.ndarray, passage_mode: bool) -> np.ndarray:
        table = self._row_enc() if passage_mode else self._doc_enc
        out = np.full(len(codes), -1, np.int32)
        found = codes >= 0
        out[found] = table[codes[found]]
        return out

    # ---- device ---------------------------------------------------------------------------
    def live(self) -> list[tuple[int, _ffx.DeviceIndex]]:
        return [(d, s) for d, s in enumerate(self.shards) if s is not None]

    def device_index(self, quantizer=None) -> _ffx.DeviceIndex:
        if self.dev is None:
            raise IndexError("The index is empty.")
        self._refresh()
        if quantizer is not None and self._shard_pq is not quantizer:
            tables = quantizer.adc_tables()
            if tables is None:
                raise RuntimeError(f"{type(quantizer).__name__} cannot be scored from its codes.")
            for _, shard in self.live():
                shard.set_pq(tables[0], tables[1])
            self._shard_pq = quantizer
        return self.dev

    def read(self, rows) -> np.ndarray:
        rows = np.asarray(rows, np.int64)
        if self.dev is None or len(rows) == 0:
            return np.array([])
        enc = self._row_enc()[rows].astype(np.int64)
        out = np.empty((len(rows), self.dev.dim), np.float32 if self.dev.row_kind == _ffx.ROWS_F32 else np.uint8)
        for d, shard in self.live():
            sel = np.flatnonzero(enc // self.stride == d)
            if len(sel):
                out[sel] = shard.read_rows(enc[sel] % self.stride)
        return out

    # ---- scoring --------------------------------------------------------------------------
    def score(self, quantizer, mode, qv, q_off, cand, lex=None, alpha=0.0, k=0, want_ff=True, out=None) -> dict:
        self.device_index(quantizer)
        live = self.live()
        parts = list(self._pool.map(
            lambda ds: ds[1].rerank_host(mode, qv, q_off, cand, lex, alpha, k, want_ff=want_ff, want_int=False), live))
        out = {} if out is None else out
        if want_ff:
            ff = out.get("ff")
            if ff is None:
                ff = out["ff"] = np.empty(len(cand), np.float32)
            owner = np.asarray(cand).astype(np.int64) // self.stride
            for (d, _), part in zip(live, parts):
                mine = owner == d
                ff[mine] = part["ff"][mine]
        if k > 0:
            if len(live) == 1:
                top_s, top_p = parts[0]["topk_score"], parts[0]["topk_pos"]
            else:
                top_s, top_p = self.dev.merge_topk_host(np.stack([p["topk_score"] for p in parts]),
                                                        np.stack([p["topk_pos"] for p in parts]))
            if out.get("topk_score") is None:
                out["topk_score"], out["topk_pos"] = top_s, top_p
            else:
                out["topk_score"][:], out["topk_pos"][:] = top_s, top_p
        return out

    def early_stop(self, quantizer, mode, qv, q_off, cand, lex, alpha, cutoff, depths) -> dict:
        raise _ffx.FFXError(-5, "early stopping on doc-id-range shards is walked by the host")
