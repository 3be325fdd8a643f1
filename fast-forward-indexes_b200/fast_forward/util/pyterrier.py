"""PyTerrier glue: score a PyTerrier result frame with an index, then interpolate.

Counterparts of `FFScore` / `FFInterpolate` in src/fast_forward/util/pyterrier.py:14-87 (same
names, constructor arguments and output columns, so pipelines written for the reference run
unchanged).  The `pyterrier` package is required; it is not part of this image, so this module
is import-checked only.
"""

from __future__ import annotations

from typing import TYPE_CHECKING

import pyterrier as pt

from fast_forward.ranking import Ranking

if TYPE_CHECKING:
    import pandas as pd

    from fast_forward.index.base import Index

_PT_TO_FF = {"qid": "q_id", "docno": "id"}
_FF_TO_PT = {ours: theirs for theirs, ours in _PT_TO_FF.items()}
_PAIR = ["qid", "docno"]


class FFScore(pt.Transformer):
    """Re-scores every (qid, docno) row with the index; the incoming score moves to `score_0`."""

    def __init__(self, index: "Index") -> None:
        super().__init__()
        self._index = index

    def transform(self, inp: "pd.DataFrame") -> "pd.DataFrame":
        # the frame is only scored, never cut: its order does not matter to the index
        pairs = Ranking(inp.rename(columns=_PT_TO_FF), copy=False, is_sorted=True)
        semantic = self._index(pairs)._df.rename(columns=_FF_TO_PT)[_PAIR + ["score", "query"]]
        lexical = inp[_PAIR + ["score"]].rename(columns={"score": "score_0"})
        return pt.model.add_ranks(semantic.merge(lexical, on=_PAIR), single_query=False)

    def __repr__(self) -> str:
        # PyTerrier caches on the representation: one per (index, query encoder)
        return f"{type(self).__name__}({id(self._index)}, {id(self._index._query_encoder)})"


class FFInterpolate(pt.Transformer):
    """`score = alpha * score_0 + (1 - alpha) * score` on the output of `FFScore`."""

    def __init__(self, alpha: float) -> None:
        super().__init__()
        self.alpha = alpha  # `pyterrier.GridScan` tunes exactly this attribute

    def transform(self, inp: "pd.DataFrame") -> "pd.DataFrame":
        mixed = self.alpha * inp["score_0"] + (1 - self.alpha) * inp["score"]
        return pt.model.add_ranks(inp[_PAIR + ["query"]].assign(score=mixed), single_query=False)
