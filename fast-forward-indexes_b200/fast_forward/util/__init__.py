"""Small helpers around rankings (reference: src/fast_forward/util/__init__.py:29-48).

The offline tooling of the reference (`Indexer`, `create_coalesced_index`, the PyTerrier
transformers) is outside the re-ranking path this package accelerates; it only uses the
public `Index` API (`add`, `batch_iter`, `_get_vectors`, `__call__`), which is kept."""

from __future__ import annotations

import numpy as np
import pandas as pd

__all__ = ["to_ir_measures", "cos_dist"]


def to_ir_measures(ranking) -> pd.DataFrame:
    """A ranking as the (query_id, doc_id, score) frame the ir-measures library expects."""
    return ranking._df[["q_id", "id", "score"]].rename(columns={"q_id": "query_id", "id": "doc_id"})


def cos_dist(a: np.ndarray, b: np.ndarray) -> float:
    """Cosine distance of two 1-d vectors."""
    assert a.ndim == b.ndim == 1
    return float(1 - np.dot(a, b) / (np.linalg.norm(a) * np.linalg.norm(b)))
