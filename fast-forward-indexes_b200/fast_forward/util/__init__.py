"""Helpers around the index: `Indexer`, evaluation frames and sequential coalescing
(drop-in for src/fast_forward/util/__init__.py:1-103)."""

from __future__ import annotations

from collections.abc import Callable, Iterator
from typing import TYPE_CHECKING

import numpy as np

from fast_forward.util.indexer import Indexer, IndexingDict

if TYPE_CHECKING:
    import pandas as pd

    from fast_forward.index.base import Index
    from fast_forward.ranking import Ranking

__all__ = ["Indexer", "IndexingDict", "to_ir_measures", "cos_dist", "create_coalesced_index"]


def to_ir_measures(ranking: "Ranking") -> "pd.DataFrame":
    """The ranking as the (query_id, doc_id, score) frame the ir-measures library takes."""
    return ranking._df[["q_id", "id", "score"]].rename(columns={"q_id": "query_id", "id": "doc_id"})


def cos_dist(a: np.ndarray, b: np.ndarray) -> float:
    """Cosine distance of two vectors."""
    assert a.ndim == b.ndim == 1
    return float(1 - np.dot(a, b) / (np.linalg.norm(a) * np.linalg.norm(b)))


def _coalesced(passages: np.ndarray, delta: float, distance: Callable[[np.ndarray, np.ndarray], float]) -> Iterator[np.ndarray]:
    """Sequential coalescing of ONE document: passages join the running group while they stay
    closer than `delta` to the group's mean; each finished group contributes its mean."""
    group: list[np.ndarray] = []
    mean = None
    for vector in passages:
        if group and distance(vector, mean) >= delta:
            yield mean
            group = []
        group.append(vector)
        mean = np.mean(group, axis=0)
    if group:
        yield mean


def _coalesce_on_device(source_index: "Index", target_index: "Index", delta: float, batch_size: int | None) -> bool:
    """The default distance on a single-device fp32 index: every document is coalesced by one warp
    (libffx `ffx_index_coalesce`), a block of documents per launch; the host only keeps the id
    bookkeeping and feeds `target_index.add`.  False when the source is of another kind (codes,
    several devices, another back-end): the host loop below takes over."""
    store = getattr(source_index, "_store", None)
    if store is None or getattr(source_index, "quantizer", None) is not None or store.dev is None:
        return False
    if hasattr(store, "shards") or store.dev.row_kind != 0:
        return False
    dev = source_index._device()
    doc_keys = store.docs.keys()
    off = store._doc_off
    n_docs = len(doc_keys)
    if n_docs == 0:
        return True
    per_block = max(1, (256 << 20) // (4 * store.dev.dim))  # rows per launch: ~256 MB of output
    held_vectors, held_ids, held = [], [], 0
    batch_size = batch_size or None

    def flush(everything: bool) -> None:
        nonlocal held_vectors, held_ids, held
        if not held:
            return
        vectors, ids = np.concatenate(held_vectors), [i for part in held_ids for i in part]
        step = batch_size or len(vectors)
        done = 0
        while len(vectors) - done >= step or (everything and done < len(vectors)):
            hi = min(len(vectors), done + step)
            target_index.add(vectors[done:hi], doc_ids=ids[done:hi])
            done = hi
        held_vectors, held_ids, held = ([vectors[done:]], [ids[done:]], len(vectors) - done) if done < len(vectors) else ([], [], 0)

    d0 = 0
    while d0 < n_docs:
        d1 = int(np.searchsorted(off, off[d0] + per_block, side="right")) - 1
        d1 = min(n_docs, max(d1, d0 + 1))
        rel = (off[d0:d1 + 1] - off[d0]).astype(np.int64)
        vectors, groups = dev.coalesce(d0, rel, delta)
        counts = np.diff(rel)
        within = np.arange(int(rel[-1])) - np.repeat(rel[:-1], counts)
        keep = within < np.repeat(groups, counts)
        held_vectors.append(vectors[keep])
        held_ids.append(np.repeat(np.asarray(doc_keys[d0:d1], dtype=object), groups).tolist())
        held += int(keep.sum())
        if batch_size and held >= batch_size:
            flush(False)
        d0 = d1
    flush(True)
    return True


def create_coalesced_index(source_index: "Index", target_index: "Index", delta: float,
                           distance_function: Callable[[np.ndarray, np.ndarray], float] = cos_dist,
                           batch_size: int | None = None) -> None:
    """Fill the (empty) `target_index` with a compressed copy of `source_index`: per document,
    consecutive passage vectors are merged by sequential coalescing with threshold `delta`.
    `batch_size` = how many coalesced vectors to collect before each `add`.
    ValueError when the target is not empty.  (util/__init__.py:51-103)

    Documents are fetched from the source in blocks (one gather from the row store per block
    instead of one per document); the coalescing itself is a short host loop per document,
    because `distance_function` is an arbitrary Python callable."""
    if len(target_index) > 0:
        raise ValueError("Target index is not empty.")
    if distance_function is cos_dist and _coalesce_on_device(source_index, target_index, delta, batch_size):
        assert source_index.doc_ids == target_index.doc_ids
        return
    doc_ids = list(source_index.doc_ids)
    batch_size = batch_size or len(doc_ids)
    held_vectors: list[np.ndarray] = []
    held_ids: list[str] = []

    def flush(count: int) -> None:
        target_index.add(np.array(held_vectors[:count]), doc_ids=held_ids[:count])
        del held_vectors[:count], held_ids[:count]

    block = 4096
    for lo in range(0, len(doc_ids), block):
        wanted = doc_ids[lo:lo + block]
        vectors, owners = source_index._get_vectors(wanted)
        rows_of: dict[str, list[int]] = {}
        for row, owner in enumerate(owners):
            rows_of.setdefault(owner, []).append(row)
        for doc_id in wanted:
            for merged in _coalesced(vectors[rows_of[doc_id]], delta, distance_function):
                held_vectors.append(merged)
                held_ids.append(doc_id)
            while len(held_vectors) >= batch_size:
                flush(batch_size)
    if held_vectors:
        flush(len(held_vectors))
    assert source_index.doc_ids == target_index.doc_ids
