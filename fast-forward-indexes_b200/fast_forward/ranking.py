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


def _coded_frame(df: pd.DataFrame, is_sorted: bool) -> pd.DataFrame | None:
    """The frame `Ranking.__init__` builds (ranking.py:95-117: duplicate check, NaN rows dropped,
    q_id DESC / score DESC stable order), computed on integer codes by libffx's host routines
    instead of pandas' `duplicated` + `sort_values` over two string columns: ids are coded by
    the C++ dictionary, pairs are checked as int64 keys, the order is one radix sort.  Returns
    None (the caller then takes the pandas route, with identical results) unless both id
    columns are strings (or integers / Python strings, converted first) without nulls and the
    scores are floats."""
    from fast_forward import _ffx, _ids

    s_col = df["score"]
    if len(df) == 0 or s_col.dtype.kind != "f":
        return None
    id_columns = {}
    for col in ("q_id", "id"):
        values = df[col]
        if values.dtype.kind in "iu" or (values.dtype == object and pd.api.types.infer_dtype(values, skipna=False) == "string"):
            # run files with numeric ids (MS MARCO), frames built from Python strings: the ids
            # become strings anyway (ranking.py:107-113); doing it first changes nothing for such
            # columns (no nulls, no mixed types) and opens the coded route
            values = values.astype(str)
        if not pd.api.types.is_string_dtype(values.dtype) or values.dtype == object or values.isna().any():
            return None
        id_columns[col] = values
    if id_columns["q_id"] is not df["q_id"] or id_columns["id"] is not df["id"]:
        df = df.assign(**id_columns)
    q_col, id_col = df["q_id"], df["id"]
    import ctypes as C

    n = len(df)
    q_code, q_keys = pd.factorize(q_col)
    id_code = _ids.IdDict().insert_ordinal(id_col)
    pair = q_code.astype(np.int64) * (int(id_code.max()) + 1) + id_code
    first = C.c_int64(-1)
    _ffx.check(_ffx.lib().ffx_first_repeat(C.c_void_p(pair.ctypes.data), n, C.byref(first)))
    if first.value >= 0:
        raise ValueError("Only one score per query-document/passage pair is allowed.")

    keep = ["q_id", "id", "score"] + (["query"] if "query" in df.columns else [])
    score = np.ascontiguousarray(s_col.to_numpy(), dtype=np.float32)  # cast first, then order (ranking.py:107-117)
    alive = ~np.isnan(score)
    if "query" in df.columns:
        alive &= ~df["query"].isna().to_numpy()
    rows = None if alive.all() else np.flatnonzero(alive)  # rows that survive dropna, in frame order
    if not is_sorted:
        # rank of every distinct q_id in DEscending string order, then one stable radix sort
        q_rank_of = np.empty(len(q_keys), np.int32)
        q_rank_of[np.argsort(np.asarray(q_keys, dtype=object), kind="stable")[::-1]] = np.arange(len(q_keys), dtype=np.int32)
        q_rank = q_rank_of[q_code]
        kept_score = score
        if rows is not None:
            q_rank, kept_score = q_rank[rows], score[rows]
        order = np.empty(len(q_rank), np.int64)
        _ffx.check(_ffx.lib().ffx_ranking_order(C.c_void_p(q_rank.ctypes.data), C.c_void_p(kept_score.ctypes.data),
                                                len(q_rank), C.c_void_p(order.ctypes.data), 0))
        rows = order if rows is None else rows[order]
    if rows is None:
        frame = df.loc[:, keep].copy()
    else:
        frame = df.loc[:, keep].take(rows)
        score = score[rows]
    frame["score"] = score
    frame.reset_index(drop=True, inplace=True)
    return frame


def _match_pairs(left: pd.DataFrame, right: pd.DataFrame):
    """Row of `right` holding the (q_id, id) pair of every row of `left`, on integer codes (one
    C++ dictionary per key column, a hash match of int64 pair keys).  Returns (q codes of left,
    id codes of left, q dictionary, id dictionary, positions), or None when the frames are
    small, differ in length, hold different pairs, or have key columns that are not plain
    strings — the callers then use pandas."""
    if len(left) != len(right) or len(left) < _CODED_FROM:
        return None
    for frame in (left, right):
        for col in _KEYS:
            dtype = frame[col].dtype
            if not pd.api.types.is_string_dtype(dtype) or dtype == object or frame[col].isna().any():
                return None
    import ctypes as C

    from fast_forward import _ffx, _ids

    n = len(left)
    q_dict, id_dict = _ids.IdDict(), _ids.IdDict()
    lq, rq = q_dict.insert_ordinal(left["q_id"]), q_dict.insert_ordinal(right["q_id"])
    li, ri = id_dict.insert_ordinal(left["id"]), id_dict.insert_ordinal(right["id"])
    n_id = len(id_dict)
    l_key, r_key = lq * n_id + li, rq * n_id + ri
    pos = np.empty(n, np.int64)
    _ffx.check(_ffx.lib().ffx_match_keys(C.c_void_p(r_key.ctypes.data), n, C.c_void_p(l_key.ctypes.data), n,
                                         C.c_void_p(pos.ctypes.data)))
    if (pos < 0).any():
        return None  # a pair of `left` is missing on the right
    return lq, li, q_dict, id_dict, pos


def _outer_coded(left: pd.DataFrame, right: pd.DataFrame) -> pd.DataFrame | None:
    """`left.merge(right, on=[q_id, id], how="outer", suffixes=(None, "_other")).fillna(0)` for the
    usual case of two rankings over the SAME set of pairs (a first-stage ranking and its
    re-scored copy), on integer codes: ids of both frames go through one C++ dictionary, the
    join is a hash match of int64 pair keys (`ffx_match_keys`), and the merge's key order —
    q_id then id, ascending as strings — comes from ranking the distinct strings once and one
    radix sort (`ffx_order_u64`).  None (pandas takes over, same result) when the pair sets
    differ, the frames are small, or the key columns are not plain strings."""
    matched = _match_pairs(left, right)
    if matched is None:
        return None
    lq, li, q_dict, id_dict, pos = matched
    try:
        import pyarrow.compute as pc
    except ImportError:  # pragma: no cover - depends on the environment
        return None
    import ctypes as C

    from fast_forward import _ffx

    def ptr(a):
        return C.c_void_p(a.ctypes.data)

    n = len(left)

    def string_rank(dictionary) -> np.ndarray:
        keys, _ = dictionary.export()
        rank = np.empty(len(keys), np.uint64)
        rank[pc.sort_indices(keys).to_numpy()] = np.arange(len(keys), dtype=np.uint64)
        return rank

    sort_key = (string_rank(q_dict)[lq] << np.uint64(32)) | string_rank(id_dict)[li]
    order = np.empty(n, np.int64)
    _ffx.check(_ffx.lib().ffx_order_u64(ptr(sort_key), n, ptr(order), 0))
    out = left.take(order).reset_index(drop=True)
    theirs = pos[order]
    for col in right.columns:
        if col not in _KEYS:
            name = col + "_other" if col in left.columns else col
            out[name] = right[col].take(theirs).reset_index(drop=True)
    return out.fillna(0) if out.isna().any().any() else out


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

        if len(df) >= _CODED_FROM and np.dtype(dtype) == np.float32:
            frame = _coded_frame(df, is_sorted)
            if frame is not None:
                self._q_ids = set(pd.unique(frame["q_id"]))
                self._df = frame if queries is None else _with_queries(frame, queries)
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

        self._q_ids = set(pd.unique(frame["q_id"]))
        self._df = frame if queries is None else _with_queries(frame, queries)

    @classmethod
    def _reordered(cls, source: "Ranking", rows: np.ndarray, scores: np.ndarray, name: str | None) -> "Ranking":
        """Rows `rows` of `source` (a selection without repeats, already in ranking order: query
        blocks in frame order, scores descending inside a block) with new scores.  The checks of
        `__init__` would only re-derive what holds by construction — keys of a valid ranking
        stay unique under row selection, dtypes are kept, the order is the kernel's — so the
        frame is adopted as is (at 26 M rows `duplicated()` alone costs more than the GPU pass).
        NaN scores never get here: the kernel does not rank them."""
        out = cls.__new__(cls)
        out.name = name
        out._origin = None
        frame = source._df.iloc[rows].reset_index(drop=True)
        frame["score"] = scores
        out._df = frame
        out._q_ids = set(pd.unique(frame["q_id"])) if len(rows) != len(source._df) else set(source._q_ids)
        return out

    # ------------------------------------------------------------------ container protocol
    @property
    def has_queries(self) -> bool:
        """Whether query texts are attached."""
        return "query" in self._df.columns

    @property
    def q_ids(self) -> set[str]:
        """Query IDs with at least one scored document."""
        return self._q_ids

    def __getitem__(self, q_id: str) -> dict[str, float]:
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
        matched = _match_pairs(self._df, o._df)
        if matched is not None:  # same pairs (both frames hold each pair once): compare aligned scores
            mine, theirs = self._df["score"].to_numpy(), o._df["score"].to_numpy()
            return bool(mine.dtype == theirs.dtype and (mine == theirs[matched[4]]).all())
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

    def _outer(self, other: pd.DataFrame, mine: pd.DataFrame | None = None) -> pd.DataFrame:
        """Outer join on (q_id, id); a score missing on either side counts as 0."""
        left = self._df if mine is None else mine
        joined = _outer_coded(left, other)
        if joined is None:
            joined = left.merge(other, on=_KEYS, suffixes=(None, "_other"), how="outer").fillna(0)
        return joined

    def __add__(self, o: "Ranking | float") -> "Ranking":
        """Add a constant or another ranking's scores (ranking.py:188-217)."""
        if isinstance(o, Ranking):
            joined = self._outer(o._df)
            joined["score"] = joined["score"] + joined["score_other"]
            return self._derive(joined, is_sorted=False)
        if isinstance(o, (int, float)):
            frame = self._df.copy()
            frame["score"] += o
            return self._derive(frame, is_sorted=True)
        return NotImplemented

    __radd__ = __add__

    def __mul__(self, o: float) -> "Ranking":
        """Multiply the scores by a constant (ranking.py:221-239)."""
        if not isinstance(o, (int, float)):
            return NotImplemented
        frame = self._df.copy()
        frame["score"] *= o
        return self._derive(frame, is_sorted=True)

    __rmul__ = __mul__

    # ------------------------------------------------------------------ transformations
    def attach_queries(self, queries: Mapping[str, str]) -> "Ranking":
        """Return a copy with query texts attached (ValueError if incomplete)."""
        return Ranking(self._df, self.name, queries=queries, dtype=self._df.dtypes["score"],
                       copy=True, is_sorted=True)

    def normalize(self) -> "Ranking":
        """Min-max normalise scores into [0, 1] (all-equal scores become 0)."""
        return self._derive(_minmax(self._df), is_sorted=True)

    def cut(self, cutoff: int) -> "Ranking":
        """Keep the `cutoff` best rows of every query (ranking.py:279-291)."""
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

        a = _minmax(self._df) if normalize else self._df
        b = _minmax(other._df) if normalize else other._df
        joined = self._outer(b, mine=a)
        joined["score"] = alpha * joined["score"] + (1 - alpha) * joined["score_other"]
        return self._derive(joined, is_sorted=False)

    def rr_scores(self, k: int = 60) -> "Ranking":
        """Reciprocal-rank scores `1 / (rank + k)` (ranking.py:328-346)."""
        frame = self._df.copy()
        frame["score"] = 1 / (_rank_column(self._df) + k)
        return self._derive(frame, is_sorted=True)

    # ------------------------------------------------------------------ I/O
    def save(self, target: Path) -> None:
        """Write a TREC run file: q_id Q0 id rank score name (ranking.py:348-366)."""
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
        table = pd.read_csv(f, sep=r"\s+", skipinitialspace=True, header=None,
                            names=["q_id", "q0", "id", "rank", "score", "name"])
        return cls(table, name=table["name"][0], queries=queries, dtype=dtype, copy=False)
