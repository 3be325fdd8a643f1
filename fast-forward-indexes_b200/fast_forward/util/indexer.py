"""`Indexer` — feeds collections into an index (drop-in for src/fast_forward/util/indexer.py:28-187).

A caller of the hot path's storage side: everything goes through the public `Index.add`,
`Index.batch_iter` and `Index.quantizer`, so rows end up in the HBM row store (pinned staging)
and ids in the C++ dictionaries no matter where they come from.
"""

from __future__ import annotations

import logging
from collections.abc import Iterable, Sequence
from typing import TYPE_CHECKING, TypedDict

import numpy as np

if TYPE_CHECKING:
    from fast_forward.encoder.base import Encoder
    from fast_forward.index.base import IDSequence, Index
    from fast_forward.quantizer import Quantizer

LOGGER = logging.getLogger(__name__)


class IndexingDict(TypedDict, total=False):
    """One document/passage for `Indexer.from_dicts`; `doc_id` and `psg_id` are optional."""

    text: str
    doc_id: str | None
    psg_id: str | None


class _FitBuffer:
    """Batches held back until a quantizer has seen enough of them (indexer.py:96-134)."""

    def __init__(self, quantizer: "Quantizer", batches_needed: int) -> None:
        self.quantizer = quantizer
        self.batches_needed = batches_needed
        self.batches: list[tuple[np.ndarray, "IDSequence | None", "IDSequence | None"]] = []

    def ready(self) -> bool:
        return len(self.batches) >= self.batches_needed


class Indexer:
    """Utility for indexing collections, optionally fitting a quantizer on the first batches."""

    def __init__(self, index: "Index", encoder: "Encoder | None" = None, encoder_batch_size: int = 128,
                 batch_size: int = 2**16, quantizer: "Quantizer | None" = None,
                 quantizer_fit_batches: int = 1) -> None:
        """`quantizer` (untrained; the index must be empty) is fit on the first
        `quantizer_fit_batches` batches, attached to the index, and the buffered batches are
        added afterwards.  ValueError for a trained quantizer or a non-empty index."""
        self._index = index
        self._encoder = encoder
        self._encoder_batch_size = encoder_batch_size
        self._batch_size = batch_size
        self._pending: _FitBuffer | None = None
        if quantizer is None:
            return
        if quantizer._trained:
            raise ValueError("The quantizer is already fit. It should be attached to the index directly.")
        if len(index) > 0:
            raise ValueError("The index must be empty for a quantizer to be attached.")
        if quantizer_fit_batches > 1:
            LOGGER.warning("inputs will be buffered and index will remain empty until the quantizer has been fit")
        self._pending = _FitBuffer(quantizer, quantizer_fit_batches)

    def _index_batch(self, vectors: np.ndarray, doc_ids: "IDSequence | None" = None,
                     psg_ids: "IDSequence | None" = None) -> None:
        pending = self._pending
        if pending is None:
            self._index.add(vectors, doc_ids, psg_ids)
            return
        pending.batches.append((vectors, doc_ids, psg_ids))
        if not pending.ready():
            return
        LOGGER.info("fitting quantizer (%s batch(es), batch size %s)", len(pending.batches), self._batch_size)
        if pending.batches[-1][0].shape[0] < self._batch_size:
            LOGGER.warning("the size of the last batch (%s) is smaller than %s",
                           pending.batches[-1][0].shape[0], self._batch_size)
        pending.quantizer.fit(np.concatenate([b[0] for b in pending.batches]))
        self._index.quantizer = pending.quantizer
        self._pending = None
        LOGGER.info("adding buffered vectors to index")
        for held in pending.batches:
            self._index.add(*held)

    def _encode(self, texts: Sequence[str]) -> np.ndarray:
        if self._encoder is None:
            raise RuntimeError("An encoder is required.")
        step = self._encoder_batch_size
        return np.concatenate([self._encoder(texts[lo:lo + step]) for lo in range(0, len(texts), step)])

    def from_dicts(self, data: Iterable[IndexingDict]) -> None:
        """Encode and index `{"text": ..., "doc_id": ..., "psg_id": ...}` items batch by batch."""
        held: list[IndexingDict] = []

        def flush() -> None:
            self._index_batch(self._encode([d["text"] for d in held]),
                              doc_ids=[d.get("doc_id") for d in held], psg_ids=[d.get("psg_id") for d in held])
            held.clear()

        for item in data:
            held.append(item)
            if len(held) == self._batch_size:
                flush()
        if held:
            flush()

    def from_index(self, index: "Index") -> None:
        """Copy vectors (reconstructed if the source is quantized) and ids from another index."""
        for vectors, doc_ids, psg_ids in index.batch_iter(self._batch_size):
            self._index_batch(vectors, doc_ids, psg_ids)
