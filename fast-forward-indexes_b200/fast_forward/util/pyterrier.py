"""PyTerrier transformers over an index (drop-in for src/fast_forward/util/pyterrier.py:14-87).
Needs the `pyterrier` package; importing this module without it raises ImportError."""

from __future__ import annotations

from typing import TYPE_CHECKING

import pyterrier as pt

from fast_forward.ranking import Ranking

if TYPE_CHECKING:
    import pandas as pd

    from fast_forward.index.base import Index


class FFScore(pt.Transformer):
    """Scores every (qid, docno) pair of a PyTerrier frame with a Fast-Forward index."""

    def __init__(self, index: "Index") -> None:
        self._index = index
        super().__init__()

    def transform(self, inp: "pd.DataFrame") -> "pd.DataFrame":
        """The semantic scores become `score`; the incoming scores move to `score_0`."""
        pairs = Ranking(inp.rename(columns={"qid": "q_id", "docno": "id"}), copy=False, is_sorted=True)
        scored = self._index(pairs)._df.rename(columns={"q_id": "qid", "id": "docno"})
        merged = scored[["qid", "docno", "score", "query"]].merge(
            inp[["qid", "docno", "score"]], on=["qid", "docno"], suffixes=(None, "_0"))
        return pt.model.add_ranks(merged, single_query=False)

    def __repr__(self) -> str:
        """Unique per index and query encoder (PyTerrier caches on it)."""
        return f"{type(self).__name__}({id(self._index)}, {id(self._index._query_encoder)})"


class FFInterpolate(pt.Transformer):
    """`alpha * score_0 + (1 - alpha) * score` over the output of `FFScore`."""

    def __init__(self, alpha: float) -> None:
        self.alpha = alpha  # this exact attribute name is what pyterrier.GridScan tunes
        super().__init__()

    def transform(self, inp: "pd.DataFrame") -> "pd.DataFrame":
        out = inp[["qid", "docno", "query"]].copy()
        out["score"] = self.alpha * inp["score_0"] + (1 - self.alpha) * inp["score"]
        return pt.model.add_ranks(out, single_query=False)
