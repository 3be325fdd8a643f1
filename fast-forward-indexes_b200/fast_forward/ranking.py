"""`Ranking` — TREC-style rankings on a pandas frame, drop-in for the reference class.

Mirrors `fast_forward.ranking.Ranking` of mrjleo/fast-forward-indexes v0.8.0
(src/fast_forward/ranking.py:64-409): same constructor contract, same public methods, same
frame invariants (`_df` has columns q_id, id, score[, query]; rows ordered by q_id DESC then
score DESC with a stable sort; scores in `dtype`).

What differs underneath: a ranking produced by `Index.__call__` remembers the integer-coded
candidate lists it was computed from (`_origin`).  `first_stage.interpolate(ff_out, alpha)`
followed by `.cut(k)` then runs interpolation and the per-query ordering on the GPU
(`ffx_rerank` epilogue kernels through libffx) instead of a pandas outer merge on two string
keys plus a lexsort — the result is the frame the reference would produce.
"""

from __future__ import annotations

import logging
from collections.abc import Iterator, Mapping
from pathlib import Path

import numpy as np
import pandas as pd

LOGGER = logging.getLogger(__name__)

Run = Mapping[str, Mapping[str, float]]

_KEYS = ["q_id", "id"]


def _with_queries(df: pd.DataFrame, queries: Mapping[str, str]) -> pd.DataFrame:
    """Left-join query texts onto a ranking frame (ranking.py:16-28)."""
    missing = set(pd.unique(df["q_id"])) - set(queries)
    if missing:
        raise ValueError("Queries are incomplete.")
    lookup = pd.DataFrame({"q_id": list(queries.keys()), "query": list(queries.values())})
    return df.merge(lookup, how="left", on="q_id")


def _rank_column(df: pd.DataFrame) -> pd.Series:
    """1-based rank of each row inside its query group (ranking.py:31-42)."""
    return (df.groupby("q_id").cumcount() + 1).rename("rank")


def _minmax(df: pd.DataFrame) -> pd.DataFrame:
    """Min-max normalise the score column over the WHOLE frame (ranking.py:45-61)."""
    out = df.copy()
    lo, hi = out["score"].min(), out["score"].max()
    if lo == hi:
        LOGGER.warning("all scores are equal, setting scores to 0")
        out["score"] = 0
    else:
        out["score"] = (out["score"] - lo) / (hi - lo)
    return out


_CODED_FROM = 50_000  # rows from which the integer-coded constructor path pays off
_NATIVE_RUN_FROM = 1 << 20  # bytes from which a run file is tokenised by libffx instead of pandas


def _cols_module():
    from fast_forward import _cols

    return _cols


class Ranking:
    """Rankings of documents/passages w.r.t. queries."""

    def __init__(
        self,
        df: pd.DataFrame,
        name: str | None = None,
        queries: Mapping[str, str] | None = None,
        dtype: np.dtype = np.dtype(np.float32),
        copy: bool = True,
        is_sorted: bool = False,
    ) -> None:
        """Create a ranking from a frame with columns q_id, id, score (optionally query).

        Rows with NaN scores are dropped; a (q_id, id) pair may appear only once
        (ValueError); `queries` must cover every q_id (ValueError).  ranking.py:67-121.
        """
        self.name = name
        self._origin = None  # set by Index.__call__ (device-side provenance)
        self._frame = None   # the pandas frame of the reference (`_df`), built on first access ...
        self._cols = None    # ... from the integer-coded columns, where those came first
        self._q_id_set = None

        if len(df) >= _CODED_FROM and np.dtype(dtype) == np.float32:
            cols = _cols_module().from_frame(df, is_sorted, queries)
            if cols is not None:
                self._cols = cols
                return

        if df.duplicated(subset=_KEYS).any():
            raise ValueError("Only one score per query-document/passage pair is allowed.")

        keep = ["q_id", "id", "score"] + (["query"] if "query" in df.columns else [])
        frame = df.loc[:, keep].dropna()
        if copy:
            frame = frame.copy()
        for col, want in (("score", dtype), ("q_id", str), ("id", str)):
            if frame[col].dtype != want:
                frame[col] = frame[col].astype(want)
        if not is_sorted:
            # q_id DESC (as strings), score DESC, stable: ties keep the incoming order
            frame.sort_values(by=["q_id", "score"], ascending=False, inplace=True)
        frame.reset_index(drop=True, inplace=True)

        self._q_id_set = set(pd.unique(frame["q_id"]))
        self._frame = frame if queries is None else _with_queries(frame, queries)

    @classmethod
    def _from_cols(cls, cols, name: str | None) -> "Ranking":
        """Adopt integer-coded columns that already satisfy the invariants of `__init__` (unique
        pairs, no NaN, float32, blocks by q_id, scores descending inside a block)."""
        out = cls.__new__(cls)
        out.name = name
        out._origin = None
        out._frame = None
        out._cols = cols
        out._q_id_set = None
        return out

    # ------------------------------------------------------------------ the two representations
    @property
    def _df(self) -> pd.DataFrame:
        """The frame the reference keeps (q_id, id, score[, query]; RangeIndex)."""
        if self._frame is None:
            self._frame = self._cols.to_frame()
        return self._frame

    @_df.setter
    def _df(self, frame: pd.DataFrame) -> None:
        self._frame = frame
        self._cols = None
        self._q_id_set = None

    @property
    def _q_ids(self) -> set[str]:
        if self._q_id_set is None:
            self._q_id_set = set(self._cols.q_keys.to_pylist()) if self._frame is None else \
                set(pd.unique(self._frame["q_id"]))
        return self._q_id_set

    def _columns(self):
        """Integer-coded columns of this ranking (`fast_forward._cols.Cols`), or None when the frame
        is not of the plain kind (float32 scores, string ids, blocks by query, no NaN queries)."""
        if self._cols is None and self._frame is not None and len(self._frame):
            frame = self._frame
            if frame["score"].dtype != np.float32:
                return None
            cm = _cols_module()
            if "query" in frame.columns:
                if frame["query"].isna().any() or not pd.api.types.is_string_dtype(frame["query"].dtype):
                    return None
                cols = cm.from_frame(frame[["q_id", "id", "score"]], True, None)
                if cols is None:
                    return None
                texts = frame["query"].iloc[cols.q_off[:-1]]
                cols.queries = cm.pa.array(texts.to_numpy(dtype=object), type=cm.pa.large_string())
            else:
                cols = cm.from_frame(frame, True, None)
            self._cols = cols
        return self._cols

    @property
    def num_rows(self) -> int:
        """Number of (query, id) pairs."""
        return len(self._cols) if self._frame is None else len(self._frame)

    # ------------------------------------------------------------------ container protocol
    @property
    def has_queries(self) -> bool:
        """Whether query texts are attached."""
        if self._frame is None:  # integer-coded columns: no pandas frame is built for the answer
            return self._cols.queries is not None
        return "query" in self._frame.columns

    @property
    def q_ids(self) -> set[str]:
        """Query IDs with at least one scored document."""
        return self._q_ids

    def __getitem__(self, q_id: str) -> dict[str, float]:
        if self._frame is None:
            cols = self._cols
            b = cols.block_of(q_id)
            if b < 0:
                return {}
            lo, hi = int(cols.q_off[b]), int(cols.q_off[b + 1])
            names = cols.ids.keys.take(_cols_module().pa.array(cols.id_code[lo:hi])).to_pylist()
            return dict(zip(names, cols.score[lo:hi]))
        rows = self._df.loc[self._df["q_id"] == q_id, ["id", "score"]]
        return dict(rows.values)

    def __len__(self) -> int:
        return len(self._q_ids)

    def __iter__(self) -> Iterator[str]:
        yield from self._q_ids

    def __contains__(self, key: object) -> bool:
        return key in self._q_ids

    def __eq__(self, o: object) -> bool:
        """Same (q_id, id, score) triples, exact float equality (ranking.py:171-186)."""
        if not isinstance(o, Ranking):
            return False
        if self.num_rows >= _CODED_FROM and self.num_rows == o.num_rows:
            a, b = self._columns(), o._columns()
            if a is not None and b is not None:  # both float32: same pairs with the same scores
                pos = _cols_module().match_pairs(a, b)
                return bool(pos is not None and (a.score == b.score[pos]).all())
        cols = ["q_id", "id", "score"]
        mine = self._df.sort_values(_KEYS).reset_index(drop=True)[cols]
        theirs = o._df.sort_values(_KEYS).reset_index(drop=True)[cols]
        return mine.equals(theirs)

    __hash__ = None  # mutable container

    def __repr__(self) -> str:
        return repr(self._df)

    # ------------------------------------------------------------------ arithmetic
    def _derive(self, frame: pd.DataFrame, is_sorted: bool, copy: bool = False) -> "Ranking":
        return Ranking(frame, name=self.name, dtype=self._df.dtypes["score"], copy=copy,
                       is_sorted=is_sorted)

    def _coded(self):
        """Columns to compute on: present already, or worth building (large plain frames)."""
        if self._cols is not None:
            return self._cols
        return self._columns() if self.num_rows >= _CODED_FROM else None

    def _rescored(self, score: np.ndarray) -> "Ranking | None":
        """Same rows, same order (`is_sorted=True`), new float32 scores; None if a NaN appeared
        (`__init__` would drop that row: the pandas route handles it)."""
        if score.dtype != np.float32 or np.isnan(score).any():
            return None
        return Ranking._from_cols(self._cols.with_scores(score), self.name)

    def _outer(self, other: pd.DataFrame, mine: pd.DataFrame | None = None) -> pd.DataFrame:
        """Outer join on (q_id, id); a score missing on either side counts as 0."""
        left = self._df if mine is None else mine
        return left.merge(other, on=_KEYS, suffixes=(None, "_other"), how="outer").fillna(0)

    def _combined(self, other: "Ranking", fn) -> "Ranking | None":
        """`fn(self.score, other.score)` over the outer merge of two rankings that hold the SAME
        pairs, on integer codes (fast_forward._cols.combine); None = take the pandas route."""
        if max(self.num_rows, other.num_rows) < _CODED_FROM and (self._cols is None or other._cols is None):
            return None
        a, b = self._columns(), other._columns()
        if a is None or b is None:
            return None
        cols = _cols_module().combine(a, b, fn)
        return None if cols is None else Ranking._from_cols(cols, self.name)

    def __add__(self, o: "Ranking | float") -> "Ranking":
        """Add a constant or another ranking's scores (ranking.py:188-217)."""
        if isinstance(o, Ranking):
            out = self._combined(o, lambda mine, theirs: mine + theirs)
            if out is not None:
                return out
            joined = self._outer(o._df)
            joined["score"] = joined["score"] + joined["score_other"]
            return self._derive(joined, is_sorted=False)
        if isinstance(o, (int, float)):
            if self._coded() is not None:
                out = self._rescored(self._cols.score + o)
                if out is not None:
                    return out
            frame = self._df.copy()
            frame["score"] += o
            return self._derive(frame, is_sorted=True)
        return NotImplemented

    __radd__ = __add__

    def __mul__(self, o: float) -> "Ranking":
        """Multiply the scores by a constant (ranking.py:221-239)."""
        if not isinstance(o, (int, float)):
            return NotImplemented
        if self._coded() is not None:
            out = self._rescored(self._cols.score * o)
            if out is not None:
                return out
        frame = self._df.copy()
        frame["score"] *= o
        return self._derive(frame, is_sorted=True)

    __rmul__ = __mul__

    # ------------------------------------------------------------------ transformations
    def attach_queries(self, queries: Mapping[str, str]) -> "Ranking":
        """Return a copy with query texts attached (ValueError if incomplete)."""
        cols = self._coded()
        if cols is not None and cols.queries is None:
            cm = _cols_module()
            try:
                texts = cm.pa.array([queries[k] for k in cols.q_keys.to_pylist()], type=cm.pa.large_string())
            except KeyError:
                raise ValueError("Queries are incomplete.") from None
            return Ranking._from_cols(cm.Cols(cols.q_keys, cols.q_off, cols.ids, cols.id_code.copy(),
                                              cols.score.copy(), texts), self.name)
        return Ranking(self._df, self.name, queries=queries, dtype=self._df.dtypes["score"],
                       copy=True, is_sorted=True)

    def normalize(self) -> "Ranking":
        """Min-max normalise scores into [0, 1] (all-equal scores become 0)."""
        if self._coded() is not None and len(self._cols):
            s = self._cols.score
            lo, hi = s.min(), s.max()
            if lo == hi:
                LOGGER.warning("all scores are equal, setting scores to 0")
                out = self._rescored(np.zeros(len(s), np.float32))
            else:
                out = self._rescored((s - lo) / (hi - lo))
            if out is not None:
                return out
        return self._derive(_minmax(self._df), is_sorted=True)

    def cut(self, cutoff: int) -> "Ranking":
        """Keep the `cutoff` best rows of every query (ranking.py:279-291)."""
        if self._coded() is not None:
            return Ranking._from_cols(self._cols.head(cutoff), self.name)
        top = self._df.groupby("q_id").head(cutoff).reset_index(drop=True)
        return self._derive(top, is_sorted=True, copy=True)

    def interpolate(self, other: "Ranking", alpha: float, normalize: bool = False) -> "Ranking":
        """`score = alpha * self.score + (1 - alpha) * other.score` (ranking.py:293-326).

        Scores missing on either side count as 0.  When `other` was computed by an index
        from this very ranking, the arithmetic and the ordering run on the GPU.
        """
        origin = getattr(other, "_origin", None)
        if origin is not None and not normalize and origin.matches(self):
            return origin.interpolate(self, other, float(alpha))

        a_r, b_r = (self.normalize(), other.normalize()) if normalize else (self, other)
        out = a_r._combined(b_r, lambda mine, theirs: alpha * mine + (1 - alpha) * theirs)
        if out is not None:
            return Ranking._from_cols(out._cols, self.name)
        a = _minmax(self._df) if normalize else self._df
        b = _minmax(other._df) if normalize else other._df
        joined = self._outer(b, mine=a)
        joined["score"] = alpha * joined["score"] + (1 - alpha) * joined["score_other"]
        return self._derive(joined, is_sorted=False)

    def rr_scores(self, k: int = 60) -> "Ranking":
        """Reciprocal-rank scores `1 / (rank + k)` (ranking.py:328-346)."""
        if self._coded() is not None:
            cols = self._cols
            rank = np.arange(1, len(cols) + 1, dtype=np.int64) - np.repeat(cols.q_off[:-1], cols.counts())
            out = self._rescored((1 / (rank + k)).astype(np.float32))
            if out is not None:
                return out
        frame = self._df.copy()
        frame["score"] = 1 / (_rank_column(self._df) + k)
        return self._derive(frame, is_sorted=True)

    # ------------------------------------------------------------------ I/O
    def save(self, target: Path) -> None:
        """Write a TREC run file: q_id Q0 id rank score name (ranking.py:348-366)."""
        cols = self._coded()
        if cols is not None and len(cols):
            target.parent.mkdir(parents=True, exist_ok=True)
            if _cols_module().write_run(cols, target, str(self.name)):
                return
        out = self._df.join(_rank_column(self._df))
        out["name"] = str(self.name)
        out["q0"] = "Q0"
        target.parent.mkdir(parents=True, exist_ok=True)
        out.to_csv(target, sep="\t", columns=["q_id", "q0", "id", "rank", "score", "name"],
                   index=False, header=False)

    @classmethod
    def from_run(cls, run: Run, name: str | None = None, queries: Mapping[str, str] | None = None,
                 dtype: np.dtype = np.dtype(np.float32)) -> "Ranking":
        """Build a ranking from `{q_id: {id: score}}` (ranking.py:368-386)."""
        # column-major stack of the (id x q_id) table: the row order the reference produces
        table = pd.DataFrame.from_dict(dict(run)).stack().reset_index()
        table.columns = ("id", "q_id", "score")
        return cls(table, name=name, queries=queries, dtype=dtype, copy=False)

    @classmethod
    def from_file(cls, f: Path, queries: Mapping[str, str] | None = None,
                  dtype: np.dtype = np.dtype(np.float32)) -> "Ranking":
        """Read a whitespace-separated TREC run file (ranking.py:388-409)."""
        import os

        if os.path.getsize(f) >= _NATIVE_RUN_FROM:  # large files: tokenised on all host cores by libffx
            parsed = _cols_module().read_run(f)
            if parsed is not None:
                return cls(parsed[0], name=parsed[1], queries=queries, dtype=dtype, copy=False)
        table = pd.read_csv(f, sep=r"\s+", skipinitialspace=True, header=None,
                            names=["q_id", "q0", "id", "rank", "score", "name"])
        return cls(table, name=table["name"][0], queries=queries, dtype=dtype, copy=False)
