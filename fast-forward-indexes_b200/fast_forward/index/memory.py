"""`InMemoryIndex` — drop-in for src/fast_forward/index/memory.py:20-180, with "memory"
meaning the GPU's HBM: rows are staged once (pinned double buffer -> device) into a libffx
row store and never copied back for scoring."""

from __future__ import annotations

import logging
from collections.abc import Iterable, Iterator, Sequence

import numpy as np

from fast_forward import _ffx
from fast_forward.encoder.base import Encoder
from fast_forward.index._store import RowStore, make_store
from fast_forward.index.base import IDSequence, Index, Mode
from fast_forward.quantizer import Quantizer

LOGGER = logging.getLogger(__name__)


class InMemoryIndex(Index):
    """Fast-Forward index held entirely in GPU memory."""

    def __init__(self, query_encoder: Encoder | None = None, quantizer: Quantizer | None = None,
                 mode: Mode = Mode.MAXP, encoder_batch_size: int = 32, init_size: int = 2**16,
                 alloc_size: int = 2**16, device: int = 0, devices: Sequence[int] | None = None,
                 shard: str = "query") -> None:
        """:param init_size: rows allocated up front. :param alloc_size: granularity of later
        growth (rows). :param device: CUDA device ordinal.
        :param devices: several CUDA devices driven by this process (not in the reference, which is
            single-process CPU code): with `shard="query"` every device holds a replica of the rows
            and each call splits its queries over them; with `shard="doc"` the documents are spread
            over the devices (a corpus larger than one GPU) and every device scores its part of
            every query.  Results are identical to one device's.  Other parameters as `Index`."""
        self._store = make_store(device, devices, shard)
        self._init_size = init_size
        self._alloc_size = alloc_size
        super().__init__(query_encoder=query_encoder, quantizer=quantizer, mode=mode,
                         encoder_batch_size=encoder_batch_size)

    @classmethod
    def _adopt(cls, device_index: _ffx.DeviceIndex, doc_ids=None, psg_ids=None, **kwargs) -> "InMemoryIndex":
        """An index over rows that already ARE in HBM (a `_ffx.DeviceIndex` staged from a device
        source: bench.py generates its 61 GB corpus on the GPU).  `doc_ids` / `psg_ids`: one id per
        row (sequence or pyarrow string array; None entries = no id), registered exactly as `add`
        registers them.  Not part of the reference API."""
        index = cls(device=device_index.device, **kwargs)
        store = index._store
        store.dev = device_index
        store.count = len(device_index)
        store.record_ids(doc_ids, psg_ids, 0, store.count)
        return index

    @classmethod
    def _adopt_replicas(cls, device_indexes, doc_ids=None, psg_ids=None, **kwargs) -> "InMemoryIndex":
        """`_adopt` for identical row stores on several devices (`devices=[...]`, `shard="query"`)."""
        index = cls(devices=[d.device for d in device_indexes], shard="query", **kwargs)
        store = index._store
        store.dev, store.replicas = device_indexes[0], list(device_indexes[1:])
        store.count = len(device_indexes[0])
        store.record_ids(doc_ids, psg_ids, 0, store.count)
        return index

    # ---- Index contract -------------------------------------------------------------------
    def _get_num_vectors(self) -> int:
        return self._store.count

    def _get_internal_dim(self) -> int | None:
        return self._store.width

    def _get_doc_ids(self) -> set[str]:
        return self._store.doc_id_set()

    def _get_psg_ids(self) -> set[str]:
        return self._store.psg_id_set()

    def _add(self, vectors: np.ndarray, doc_ids: IDSequence, psg_ids: IDSequence) -> None:
        """Stage rows into HBM.  Vectors are stored as float32 (codes as uint8): integer or
        float64 input is converted, unlike the reference which keeps the first chunk's dtype
        (index/memory.py:79-82)."""
        if self.quantizer is not None:
            if vectors.dtype != np.uint8:
                raise RuntimeError("Only uint8 codes (Ks <= 256) can be stored on the device.")
            rows = np.ascontiguousarray(vectors)
        else:
            rows = np.ascontiguousarray(vectors, dtype=np.float32)
        self._store.append(rows, doc_ids, psg_ids, self._init_size, self._alloc_size)

    def consolidate(self) -> None:
        """Kept for API compatibility: the device store is always one contiguous array."""

    def _get_vectors(self, ids: Iterable[str]) -> tuple[np.ndarray, list[str]]:
        rows, owners = self._store.rows_for(ids, self.mode.name)
        return self._store.read(rows), owners

    def _batch_iter(self, batch_size: int) -> Iterator[tuple[np.ndarray, IDSequence, IDSequence]]:
        total = len(self)
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
