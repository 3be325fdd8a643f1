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
