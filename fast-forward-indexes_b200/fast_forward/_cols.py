"""Columnar, integer-coded form of a ranking frame — what `Ranking` computes on.

The reference keeps every ranking as a pandas frame with two string key columns and runs
`duplicated`, `sort_values`, `merge`, `groupby` over them (ranking.py:95-117,188-217,279-326);
at the sizes of BASELINE.json (2.6 x 10^7 pairs) that string work dwarfs the GPU pass.  Here a
ranking is

    q_keys  [nq]      the q_ids in frame order (q_id DESC as strings: ranking.py:115-117)
    q_off   [nq + 1]  rows [q_off[b], q_off[b+1]) of the frame belong to q_keys[b]
    ids               an `IdTable`: the DISTINCT id strings of the ranking (any order)
    id_code [n]       int32 code of every row's id into `ids`
    score   [n]       float32, descending inside a block (stable)
    queries [nq]      query text per block, or None

and the pandas frame the reference would hold (`Ranking._df`) is built from it on first access.
An index resolves the distinct ids once (`IdTable.lut_for`), not every pair of every call, and
the per-pair candidate array is cached on the columns, so a second `index(ranking)` /
`index.rerank(ranking, ...)` hashes no string at all.  Host-only: no CUDA device is needed.
"""

from __future__ import annotations

import ctypes as C
import os

import numpy as np
import pandas as pd

from fast_forward import _ffx, _ids

pa = _ids.pa


def _str_dtype():
    return pd.StringDtype("pyarrow", na_value=np.nan)


def _series(arr) -> pd.Series:
    """A pandas `str` column (what `astype(str)` produces, ranking.py:107-113) over an Arrow array."""
    chunked = arr if isinstance(arr, pa.ChunkedArray) else pa.chunked_array([arr])
    return pd.Series(pd.arrays.ArrowStringArray(chunked, dtype=_str_dtype()), copy=False)


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


_TAKE_FROM = 1 << 20


def _take(values, indices: np.ndarray):
    """`values.take(indices)` for a (long) integer index array: pyarrow gathers strings on one
    thread, so long gathers are cut into slices taken concurrently (the GIL is released) and come
    back as one chunked array."""
    n = len(indices)
    threads = min(os.cpu_count() or 1, 32, n // _TAKE_FROM)
    if threads < 2:
        return values.take(pa.array(indices))
    from concurrent.futures import ThreadPoolExecutor

    step = -(-n // threads)
    with ThreadPoolExecutor(threads) as pool:
        parts = list(pool.map(lambda lo: values.take(pa.array(indices[lo:lo + step])), range(0, n, step)))
    return pa.chunked_array(parts)


class IdTable:
    """The distinct id strings of one (family of) ranking(s): a pyarrow large_string array plus,
    per index store, the candidate of every id (`lut_for`).  Rankings derived from one another
    (cut, interpolate, the output of `Index.__call__`) share the table."""

    __slots__ = ("keys", "_luts", "_dict", "_rank")

    def __init__(self, keys) -> None:
        if isinstance(keys, pa.ChunkedArray):
            keys = keys.combine_chunks()
        if keys.type != pa.large_string():
            keys = keys.cast(pa.large_string())
        self.keys = keys
        self._luts: dict = {}
        self._dict = None  # IdDict over `keys` (value = code), built on demand
        self._rank = None  # ascending string rank of every key, built on demand

    def __len__(self) -> int:
        return len(self.keys)

    def lut_for(self, store, passage_mode: bool) -> np.ndarray:
        """int32 candidate (document ordinal / row number) of every id of the table in `store`,
        -1 where the index does not hold the id.  Cached per (store, contents version, mode)."""
        key = (store.token, store.version, bool(passage_mode))
        lut = self._luts.get(key)
        if lut is None:
            lut = store.lookup_keys(self.keys, passage_mode)
            if len(self._luts) > 8:
                self._luts.clear()
            self._luts[key] = lut
        return lut

    def as_dict(self) -> _ids.IdDict:
        if self._dict is None:
            d = _ids.IdDict()
            codes = d.insert_ordinal(self.keys)
            assert len(d) == len(self.keys) and (len(codes) == 0 or codes[-1] == len(codes) - 1)
            self._dict = d
        return self._dict

    def codes_of(self, other: "IdTable") -> np.ndarray:
        """Code in THIS table of every key of `other` (-1 = absent)."""
        if other is self:
            return np.arange(len(self), dtype=np.int32)
        codes, _ = self.as_dict().lookup(other.keys)
        return codes

    def string_rank(self) -> np.ndarray:
        """Position of every key in ascending string order (the key order of a pandas merge)."""
        if self._rank is None:
            import pyarrow.compute as pc

            rank = np.empty(len(self.keys), np.int64)
            rank[pc.sort_indices(self.keys).to_numpy()] = np.arange(len(self.keys))
            self._rank = rank
        return self._rank


class Cols:
    """One ranking frame in columnar form (see the module docstring)."""

    __slots__ = ("q_keys", "q_off", "ids", "id_code", "score", "queries", "_cand", "_q_index", "__weakref__")

    def __init__(self, q_keys, q_off, ids: IdTable, id_code, score, queries=None) -> None:
        self.q_keys = q_keys          # pa.Array large_string [nq]
        self.q_off = q_off            # int64 [nq + 1]
        self.ids = ids
        self.id_code = id_code        # int32 [n]
        self.score = score            # float32 [n]
        self.queries = queries        # pa.Array large_string [nq] | None
        self._cand: dict = {}
        self._q_index = None

    # ---- shape --------------------------------------------------------------------------
    @property
    def nq(self) -> int:
        return len(self.q_off) - 1

    def __len__(self) -> int:
        return int(self.q_off[-1])

    def counts(self) -> np.ndarray:
        return np.diff(self.q_off)

    def block_of(self, q_id: str) -> int:
        if self._q_index is None:
            self._q_index = {k: i for i, k in enumerate(self.q_keys.to_pylist())}
        return self._q_index.get(q_id, -1)

    def with_scores(self, score: np.ndarray) -> "Cols":
        """Same rows, new scores (order kept as is: `is_sorted=True` of the reference)."""
        return Cols(self.q_keys, self.q_off, self.ids, self.id_code, score, self.queries)

    # ---- index side ---------------------------------------------------------------------
    def candidates(self, store, passage_mode: bool) -> np.ndarray:
        """int32 candidate per row for ffx_rerank (page-locked).  IndexError names the first row's
        id that the index does not hold (index/util.py:38-39)."""
        key = (store.token, store.version, bool(passage_mode))
        cand = self._cand.get(key)
        if cand is None:
            lut = self.ids.lut_for(store, passage_mode)
            n = len(self.id_code)
            cand = _ffx.pinned_empty(n, np.int32)
            miss = C.c_int64(-1)
            _ffx.check(_ffx.lib().ffx_lut_gather(_ptr(lut), len(lut), _ptr(self.id_code), n, _ptr(cand),
                                                 C.byref(miss), 0))
            if miss.value >= 0:
                name = self.ids.keys[int(self.id_code[miss.value])].as_py()
                raise IndexError(f"ID {name} not found in the index.")
            self._cand.clear()
            self._cand[key] = cand
        return cand

    # ---- frame --------------------------------------------------------------------------
    def to_frame(self) -> pd.DataFrame:
        """The frame the reference keeps in `Ranking._df`: q_id, id (pandas `str`), score[, query],
        RangeIndex."""
        block = np.repeat(np.arange(self.nq, dtype=np.int32), self.counts())
        data = {"q_id": _series(_take(self.q_keys, block)),
                "id": _series(_take(self.ids.keys, self.id_code)),
                "score": np.array(self.score, copy=True)}
        if self.queries is not None:
            data["query"] = _series(_take(self.queries, block))
        return pd.DataFrame(data, copy=False)

    # ---- derived rankings ---------------------------------------------------------------
    def head(self, k: int) -> "Cols":
        """ranking.py:279-291: the first k rows of every block."""
        counts = self.counts()
        keep = np.minimum(counts, max(int(k), 0))
        if (keep == counts).all():
            return Cols(self.q_keys, self.q_off, self.ids, self.id_code.copy(), self.score.copy(), self.queries)
        return self.take_heads(keep)

    def take_heads(self, keep: np.ndarray) -> "Cols":
        off = np.zeros(self.nq + 1, np.int64)
        np.cumsum(keep, out=off[1:])
        src = np.repeat(self.q_off[:-1] - off[:-1], keep) + np.arange(off[-1])
        alive = np.flatnonzero(keep > 0)
        q_keys, queries = self.q_keys, self.queries
        if len(alive) != self.nq:  # blocks cut to nothing disappear (q_ids = queries with a scored row)
            picks = pa.array(alive)
            q_keys = q_keys.take(picks)
            queries = None if queries is None else queries.take(picks)
            off = np.concatenate([[0], np.cumsum(keep[alive])]).astype(np.int64)
        return Cols(q_keys, off, self.ids, self.id_code[src], self.score[src], queries)

    def from_lists(self, pos: np.ndarray, score: np.ndarray, keep: int, want_straddle: bool = False):
        """Rows of a result ranking from the [nq, k] (position, score) lists of ffx_rerank over THIS
        ranking's blocks: block b keeps its first min(keep, valid) entries.  Returns
        (cols, n_ties, straddle | None)."""
        nq, k = pos.shape
        assert nq == self.nq and keep <= k
        off = np.empty(nq + 1, np.int64)
        dense = keep == k or nq == 0
        m_cap = nq * keep
        code = _ffx.pinned_empty(m_cap, np.int32)  # recycled block: no page faults on 100 MB per call
        # when every list is full (no -1 padding) the [nq, keep] score matrix IS the column
        out_score = None if dense else np.empty(m_cap, np.float32)
        ties = C.c_int64(0)
        straddle = np.zeros(nq, np.uint8) if want_straddle and keep < k else None
        pos = np.ascontiguousarray(pos, np.int32)
        score = np.ascontiguousarray(score, np.float32)
        _ffx.check(_ffx.lib().ffx_topk_gather(_ptr(pos), _ptr(score), nq, k, keep, _ptr(self.q_off), _ptr(self.id_code),
                                              _ptr(off), _ptr(code), _ptr(out_score), C.byref(ties),
                                              _ptr(straddle), 0))
        m = int(off[-1])
        if dense and m == m_cap:
            out_score = score.reshape(-1)
        elif dense:  # padded lists after all: compact the scores too
            out_score = np.empty(m_cap, np.float32)
            _ffx.check(_ffx.lib().ffx_topk_gather(_ptr(pos), _ptr(score), nq, k, keep, _ptr(self.q_off), None,
                                                  _ptr(off), None, _ptr(out_score), None, None, 0))
        # blocks without ranked rows (all-NaN scores) stay as empty blocks: `drop_empty` removes them
        return Cols(self.q_keys, off, self.ids, code[:m], out_score[:m], self.queries), int(ties.value), straddle

    def drop_empty(self) -> "Cols":
        """Blocks without rows disappear (q_ids = the queries with at least one scored row)."""
        counts = self.counts()
        return self if (counts > 0).all() else self.take_heads(counts)

    def order_ties_by_id(self, ties_hint: int | None = None) -> None:
        """The reference leaves equal scores of a query in ascending id order after an outer merge
        (ranking.py:312-326: the merge sorts its keys, the sort that follows is stable); the
        kernels order ties by position.  Re-orders, in place, the runs of equal scores (found by
        ffx_tie_runs on all host cores; `ties_hint` = an upper bound of their number, if known)."""
        n = len(self.score)
        if n < 2 or ties_hint == 0:
            return
        score = np.ascontiguousarray(self.score, np.float32)
        cap = int(ties_hint) if ties_hint else 1024
        while True:
            starts, lens = np.empty(cap, np.int64), np.empty(cap, np.int64)
            found = C.c_int64(0)
            _ffx.check(_ffx.lib().ffx_tie_runs(_ptr(score), _ptr(self.q_off), self.nq, cap, _ptr(starts), _ptr(lens),
                                               C.byref(found), 0))
            if found.value <= cap:
                break
            cap = int(found.value)
        runs = int(found.value)
        if runs == 0:
            return
        starts, lens = starts[:runs], lens[:runs]
        # only the ids inside runs need an order: rank their DISTINCT strings among themselves
        # (a few thousand), never the whole id table (millions)
        members = np.repeat(starts - np.concatenate([[0], np.cumsum(lens)[:-1]]), lens) + np.arange(int(lens.sum()))
        codes = self.id_code[members]
        distinct, inverse = np.unique(codes, return_inverse=True)
        names = np.asarray(self.ids.keys.take(pa.array(distinct)).to_pylist(), dtype=object)
        rank_of_distinct = np.empty(len(distinct), np.int64)
        rank_of_distinct[np.argsort(names, kind="stable")] = np.arange(len(distinct))
        rank = rank_of_distinct[inverse]
        # one stable sort by (run, id rank) re-orders every run at once
        run_of = np.repeat(np.arange(runs), lens)
        order = np.lexsort((rank, run_of))
        self.id_code[members] = codes[order]
        self._cand.clear()


# ------------------------------------------------------------------------------------------
# frame -> columns
# ------------------------------------------------------------------------------------------
def _plain_strings(values: pd.Series):
    """The column as a null-free pandas `str` column, or None if it is something else.  Integer
    columns (run files with numeric ids) and Python-string object columns become strings first,
    exactly what ranking.py:107-113 does to them."""
    if values.dtype.kind in "iu" or (values.dtype == object and pd.api.types.infer_dtype(values, skipna=False) == "string"):
        values = values.astype(str)
    if not pd.api.types.is_string_dtype(values.dtype) or values.dtype == object or values.isna().any():
        return None
    return values


def from_frame(df: pd.DataFrame, is_sorted: bool, queries=None):
    """`Ranking.__init__` (ranking.py:95-121) on integer codes: duplicate-pair check, NaN rows
    dropped, scores cast to float32, q_id DESC / score DESC stable order, queries attached.
    Returns Cols, or None when the frame is not of the plain kind (the caller then takes the
    pandas route with identical results): key columns must be strings (or integers / Python
    strings) without nulls, scores floats, no `query` column."""
    if len(df) == 0 or df["score"].dtype.kind != "f" or "query" in df.columns:
        return None
    q_col, id_col = _plain_strings(df["q_id"]), _plain_strings(df["id"])
    if q_col is None or id_col is None:
        return None
    n = len(df)
    q_code, q_uniques = _ids.factorize(q_col)  # all host cores; the numbering is arbitrary
    id_code, id_keys = _ids.factorize(id_col)
    n_ids = len(id_keys)
    pair = q_code.astype(np.int64) * n_ids + id_code
    first = C.c_int64(-1)
    _ffx.check(_ffx.lib().ffx_first_repeat(_ptr(pair), n, C.byref(first)))
    del pair
    if first.value >= 0:
        raise ValueError("Only one score per query-document/passage pair is allowed.")

    score = np.ascontiguousarray(df["score"].to_numpy(), dtype=np.float32)  # cast, then order (ranking.py:107-117)
    alive = ~np.isnan(score)
    rows = None if alive.all() else np.flatnonzero(alive)
    q_names = np.asarray(q_uniques.to_pylist(), dtype=object)
    if not is_sorted:
        q_rank_of = np.empty(len(q_names), np.int32)
        q_rank_of[np.argsort(q_names, kind="stable")[::-1]] = np.arange(len(q_names), dtype=np.int32)
        q_rank = q_rank_of[q_code]
        kept = score
        if rows is not None:
            q_rank, kept = q_rank[rows], score[rows]
        order = np.empty(len(q_rank), np.int64)
        _ffx.check(_ffx.lib().ffx_ranking_order(_ptr(q_rank), _ptr(kept), len(q_rank), _ptr(order), 0))
        rows = order if rows is None else rows[order]
    if rows is not None:
        q_code, id_code, score = q_code[rows], id_code[rows], score[rows]
    if len(q_code) == 0:
        return None
    # blocks: every q_id must own one contiguous run of rows
    change = np.flatnonzero(q_code[1:] != q_code[:-1]) + 1
    starts = np.concatenate([[0], change])
    block_q = q_code[starts]
    if len(np.unique(block_q)) != len(block_q):
        return None  # is_sorted=True on a frame that is not grouped by query: pandas keeps it as is
    q_off = np.concatenate([starts, [len(q_code)]]).astype(np.int64)
    q_keys = pa.array(q_names[block_q], type=pa.large_string())
    texts = None
    if queries is not None:
        try:
            texts = pa.array([queries[k] for k in q_names[block_q]], type=pa.large_string())
        except KeyError:
            raise ValueError("Queries are incomplete.") from None
    pinned = _ffx.pinned_empty(len(score), np.float32)
    pinned[:] = score
    return Cols(q_keys, q_off, IdTable(id_keys), np.ascontiguousarray(id_code, np.int32), pinned, texts)


def match_pairs(a: Cols, b: Cols):
    """Row of `b` holding the (q_id, id) pair of every row of `a`, or None when the two rankings
    do not hold exactly the same pairs."""
    if len(a) != len(b):
        return None
    qa, qb = a.q_keys.to_pylist(), b.q_keys.to_pylist()
    if a.q_keys is b.q_keys or qa == qb:
        q_map = np.arange(len(qb), dtype=np.int64)
    else:
        where = {k: i for i, k in enumerate(qa)}
        q_map = np.array([where.get(k, -1) for k in qb], np.int64)
        if (q_map < 0).any() or len(set(qb)) != len(qa):
            return None
    id_map = a.ids.codes_of(b.ids).astype(np.int64)  # b's id code -> a's
    n_ids = len(a.ids) + 1
    key_a = np.repeat(np.arange(a.nq, dtype=np.int64), a.counts()) * n_ids + a.id_code
    b_ids = id_map[b.id_code]
    if (b_ids < 0).any():
        return None
    key_b = np.repeat(q_map, b.counts()) * n_ids + b_ids
    pos = np.empty(len(a), np.int64)
    _ffx.check(_ffx.lib().ffx_match_keys(_ptr(key_b), len(key_b), _ptr(key_a), len(key_a), _ptr(pos)))
    if (pos < 0).any():
        return None
    return pos


def combine(a: Cols, b: Cols, fn):
    """An outer merge of two rankings over the SAME pairs followed by `Ranking.__init__`
    (ranking.py:188-217,293-326): score = fn(a.score, b.score aligned), rows re-ordered by score
    DESC inside every block, equal scores in ascending id order.  None when the pair sets differ."""
    pos = match_pairs(a, b)
    if pos is None:
        return None
    score = np.asarray(fn(a.score, b.score[pos]), dtype=np.float32)
    if np.isnan(score).any():
        return None
    # the merge leaves rows in (q_id, id) ascending string order; the stable sort by score that
    # follows keeps that order among ties: sort key = (block, score desc, id rank)
    block = np.repeat(np.arange(a.nq, dtype=np.int32), a.counts())
    order = np.empty(len(score), np.int64)
    _ffx.check(_ffx.lib().ffx_ranking_order(_ptr(block), _ptr(score), len(score), _ptr(order), 0))
    # ffx_ranking_order breaks ties by incoming position; make the incoming order id-ascending
    # inside blocks only if there are ties at all (checked after the sort, on the sorted scores)
    queries = a.queries
    if queries is None and b.queries is not None:  # the merge carries the right frame's `query` column over
        where = {k: i for i, k in enumerate(b.q_keys.to_pylist())}
        queries = b.queries.take(pa.array([where[k] for k in a.q_keys.to_pylist()], type=pa.int64()))
    out = Cols(a.q_keys, a.q_off, a.ids, a.id_code[order], score[order], queries)
    out.order_ties_by_id()
    return out


# ------------------------------------------------------------------------------------------
# run files
# ------------------------------------------------------------------------------------------
def read_run(path) -> tuple[pd.DataFrame, str] | None:
    """`pd.read_csv(f, sep=r"\\s+", header=None, names=[q_id, q0, id, rank, score, name])` of
    ranking.py:401-402 for the columns a ranking keeps, tokenised on all host cores by libffx
    (`ffx_run_open`): a frame with q_id / id (pandas `str`) and score (float64, the very doubles
    pandas' default float parser yields) plus the first row's name.  None when pandas would have
    produced something else from this file — all-numeric id columns whose tokens are not
    canonical integers (pandas renumbers them), NA or boolean tokens, quotes, lines without six
    fields, scores that are not plain decimals — the caller then lets pandas read it."""
    info = np.zeros(16, np.int64)
    handle = C.c_void_p()
    _ffx.check(_ffx.lib().ffx_run_open(str(path).encode(), 0, C.byref(handle), _ptr(info)))
    try:
        rows = int(info[0])
        if rows == 0 or info[3] or info[4] or info[5] or info[14]:
            return None
        for col in (0, 1):
            numeric, canonical, na, boolean = (int(x) for x in info[6 + 4 * col:10 + 4 * col])
            if na or boolean == rows or (numeric == rows and canonical != rows):
                return None
        q_off, id_off = np.empty(rows + 1, np.int64), np.empty(rows + 1, np.int64)
        q_data, id_data = np.empty(max(int(info[1]), 1), np.uint8), np.empty(max(int(info[2]), 1), np.uint8)
        score = np.empty(rows, np.float64)
        name = C.create_string_buffer(int(info[15]) + 1)
        _ffx.check(_ffx.lib().ffx_run_read(handle, _ptr(q_off), _ptr(q_data), _ptr(id_off), _ptr(id_data), _ptr(score), name))
    finally:
        _ffx.lib().ffx_run_close(handle)
    q_arr = pa.Array.from_buffers(pa.large_string(), rows, [None, pa.py_buffer(q_off), pa.py_buffer(q_data)])
    id_arr = pa.Array.from_buffers(pa.large_string(), rows, [None, pa.py_buffer(id_off), pa.py_buffer(id_data)])
    frame = pd.DataFrame({"q_id": _series(q_arr), "id": _series(id_arr), "score": score}, copy=False)
    return frame, name.raw[:int(info[15])].decode("utf-8")


def write_run(cols: "Cols", path, name: str) -> bool:
    """`Ranking.save` (ranking.py:348-366) from the columns, formatted on all host cores
    (`ffx_run_write`); False when an id needs the csv writer's quoting (tab, quote, newline in an
    id or in the name): pandas writes those files."""
    import pyarrow.compute as pc

    for strings in (cols.q_keys, cols.ids.keys):
        if len(strings) and pc.any(pc.match_substring_regex(strings, '[\\t\\n\\r"]')).as_py():
            return False
    if any(ch in name for ch in "\t\n\r\""):
        return False

    def buffers(arr):
        if arr.type != pa.large_string():
            arr = arr.cast(pa.large_string())
        _, offsets, data = arr.buffers()
        return arr, offsets.address + 8 * arr.offset, (data.address if data is not None else 0)

    qk, qk_off, qk_data = buffers(cols.q_keys)
    ik, ik_off, ik_data = buffers(cols.ids.keys)
    score = np.ascontiguousarray(cols.score, np.float32)
    code = np.ascontiguousarray(cols.id_code, np.int32)
    dummy = np.zeros(1, np.uint8)
    _ffx.check(_ffx.lib().ffx_run_write(str(path).encode(), cols.nq, _ptr(cols.q_off), C.c_void_p(qk_off),
                                        C.c_void_p(qk_data or dummy.ctypes.data), C.c_void_p(ik_off),
                                        C.c_void_p(ik_data or dummy.ctypes.data), _ptr(code), _ptr(score),
                                        name.encode("utf-8"), 0))
    del qk, ik
    return True
