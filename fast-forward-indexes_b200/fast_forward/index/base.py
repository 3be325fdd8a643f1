"""`Mode` and the abstract `Index` — the scoring front-end, drop-in for
src/fast_forward/index/base.py of the reference (v0.8.0).

Public surface and error behaviour follow the reference (`__call__`, `add`, `encode_queries`,
`mode` / `quantizer` / `query_encoder` properties, `dim`, `doc_ids`, `psg_ids`, `batch_iter`,
iteration).  What differs is everything underneath `_compute_scores`: vectors (or PQ codes)
live in HBM inside a libffx index, ids are integer-coded once on the host with exactly the
semantics of index/util.py:12-42, and one CUDA launch does look-up, dot products, the
per-document reduce, and — through `rerank` or `Ranking.interpolate` on the result — the
interpolation and per-query ordering.  There is no numpy scoring path.
"""

from __future__ import annotations

import abc
import logging
from collections.abc import Iterable, Iterator, Sequence
from enum import Enum
from time import perf_counter

import numpy as np
import pandas as pd

from fast_forward import _ffx
from fast_forward.encoder.base import Encoder
from fast_forward.quantizer import Quantizer
from fast_forward.ranking import Ranking

LOGGER = logging.getLogger(__name__)

IDSequence = Sequence["str | None"]


class Mode(Enum):
    """Ranking mode of an index (same values as the reference and as ffx_mode)."""

    PASSAGE = 1
    MAXP = 2
    FIRSTP = 3
    AVEP = 4


def _fl32_interpolate(alpha: float, lex, ff):
    """ranking.py:319 arithmetic: fl32(fl32(alpha)*lex) + fl32(fl32(1-alpha)*ff)."""
    a, b = np.float32(alpha), np.float32(1 - alpha)
    return a * np.asarray(lex, np.float32) + b * np.asarray(ff, np.float32)


class _Origin:
    """Provenance of a ranking computed by `Index.__call__`: the integer-coded columns of the
    source ranking (identity) and the semantic scores aligned with its rows.  Lets
    `source.interpolate(result, alpha)` skip the string outer-merge and run on the GPU."""

    __slots__ = ("index", "source", "ff")

    def __init__(self, index: "Index", source, ff: np.ndarray):
        self.index, self.source, self.ff = index, source, ff

    def matches(self, ranking: Ranking) -> bool:
        return ranking._cols is self.source and len(self.source) == len(self.ff)

    def interpolate(self, first: Ranking, _other: Ranking, alpha: float) -> Ranking:
        src = self.source
        max_c = int(src.counts().max()) if src.nq else 0
        out = self.index._device().interpolate_topk_host(src.score, self.ff, src.q_off, alpha, max_c,
                                                         want_int=False)
        cols, ties, _ = src.from_lists(out["topk_pos"], out["topk_score"], max_c)
        cols.order_ties_by_id(ties)
        return Ranking._from_cols(cols, first.name)


def _number_queries(q_ids: pd.Series):
    """`q_no` per row in order of first appearance + the first row of every query
    (index/base.py:418-422)."""
    q_codes, _ = pd.factorize(q_ids)
    first_row = np.unique(q_codes, return_index=True)[1]
    return q_codes.astype(np.int64), first_row


class Index(abc.ABC):
    """Abstract base class for Fast-Forward indexes."""

    _query_encoder: Encoder | None = None
    _quantizer: Quantizer | None = None

    def __init__(self, query_encoder: Encoder | None = None, quantizer: Quantizer | None = None,
                 mode: Mode = Mode.MAXP, encoder_batch_size: int = 32) -> None:
        """:param query_encoder: the query encoder. :param quantizer: the quantizer (only
        attachable while the index is empty). :param mode: the ranking mode.
        :param encoder_batch_size: queries per encoder call."""
        super().__init__()
        if query_encoder is not None:
            self.query_encoder = query_encoder
        self.mode = mode
        if quantizer is not None:
            self.quantizer = quantizer
        self._encoder_batch_size = encoder_batch_size

    # ------------------------------------------------------------------ properties
    def encode_queries(self, queries: Sequence[str]) -> np.ndarray:
        """Encode queries in batches of `encoder_batch_size` (RuntimeError without encoder)."""
        if self.query_encoder is None:
            raise RuntimeError("Index does not have a query encoder.")
        step = self._encoder_batch_size
        parts = [self.query_encoder(queries[i:i + step]) for i in range(0, len(queries), step)]
        return np.concatenate(parts)

    @property
    def query_encoder(self) -> Encoder | None:
        return self._query_encoder

    @query_encoder.setter
    def query_encoder(self, encoder: Encoder) -> None:
        assert isinstance(encoder, Encoder)
        self._query_encoder = encoder

    @property
    def quantizer(self) -> Quantizer | None:
        return self._quantizer

    def _on_quantizer_set(self) -> None:
        """Hook for back-ends (the on-disk index persists the quantizer)."""

    @quantizer.setter
    def quantizer(self, quantizer: Quantizer) -> None:
        """Attach a (trained) quantizer; RuntimeError unless the index is empty."""
        assert isinstance(quantizer, Quantizer)
        if len(self) > 0:
            raise RuntimeError("Quantizers can only be attached to empty indexes.")
        self._quantizer = quantizer
        self._on_quantizer_set()
        quantizer.set_attached()

    @property
    def mode(self) -> Mode:
        return self._mode

    @mode.setter
    def mode(self, mode: Mode) -> None:
        assert isinstance(mode, Mode)
        self._mode = mode

    @property
    def dim(self) -> int | None:
        """Vector dimensionality (of the ORIGINAL vectors when a quantizer is attached);
        None while the index is empty and has no quantizer."""
        if self._quantizer is not None:
            return self._quantizer.dims[0]
        return self._get_internal_dim()

    @property
    def doc_ids(self) -> set[str]:
        return self._get_doc_ids()

    @property
    def psg_ids(self) -> set[str]:
        return self._get_psg_ids()

    def __len__(self) -> int:
        return self._get_num_vectors()

    # ------------------------------------------------------------------ back-end contract
    @abc.abstractmethod
    def _get_internal_dim(self) -> int | None: ...

    @abc.abstractmethod
    def _get_doc_ids(self) -> set[str]: ...

    @abc.abstractmethod
    def _get_psg_ids(self) -> set[str]: ...

    @abc.abstractmethod
    def _get_num_vectors(self) -> int: ...

    @abc.abstractmethod
    def _add(self, vectors: np.ndarray, doc_ids: IDSequence, psg_ids: IDSequence) -> None:
        """Append (possibly quantized) vectors with their ids."""

    @abc.abstractmethod
    def _get_vectors(self, ids: Iterable[str]) -> tuple[np.ndarray, list[str]]:
        """Vectors (codes when quantized) needed to score `ids` in the current mode, with one
        id per returned row.  IndexError for an unknown id."""

    @abc.abstractmethod
    def _batch_iter(self, batch_size: int) -> Iterator[tuple[np.ndarray, IDSequence, IDSequence]]: ...

    @abc.abstractmethod
    def _device(self) -> _ffx.DeviceIndex:
        """The libffx index holding this index's rows in HBM, maps synchronised."""

    @abc.abstractmethod
    def _score(self, mode: Mode, qv, q_off, cand, lex=None, alpha: float = 0.0, k: int = 0, want_ff: bool = True,
               out: dict | None = None) -> dict:
        """ffx_rerank_host over the index's device(s): ff [n] and / or topk_score, topk_pos [nq, k]."""

    @abc.abstractmethod
    def _early_stop(self, mode: Mode, qv, q_off, cand, lex, alpha, cutoff, depths) -> dict:
        """ffx_rerank_early_stop_host over the index's device(s)."""

    @abc.abstractmethod
    def _candidates(self, cols, mode: Mode) -> np.ndarray:
        """int32 candidate per row of an integer-coded ranking (`fast_forward._cols.Cols`), cached
        on the columns; IndexError names the first unknown id (index/util.py:38-39)."""

    @abc.abstractmethod
    def _resolve(self, ids, mode: Mode, missing_ok: bool = False) -> np.ndarray:
        """An id column (Series / array, one entry per pair) -> int32 candidates for ffx_rerank
        (document ordinals, or row numbers in PASSAGE mode) with the semantics of
        index/util.py:29-41; IndexError names the first unknown id — or, with `missing_ok`, unknown
        ids come back as -1 (early stopping only fails for the ids it actually scores)."""

    # ------------------------------------------------------------------ adding
    def add(self, vectors: np.ndarray, doc_ids: IDSequence | None = None,
            psg_ids: IDSequence | None = None) -> None:
        """Add vectors with document and/or passage ids (index/base.py:211-256).

        :raises ValueError: id count mismatch, dimension mismatch, or a vector without any id.
        :raises RuntimeError: when the back-end cannot add the items (duplicate passage id).
        """
        count, dim = vectors.shape
        doc_ids = [None] * count if doc_ids is None else doc_ids
        psg_ids = [None] * count if psg_ids is None else psg_ids
        if not len(doc_ids) == len(psg_ids) == count:
            raise ValueError("Number of IDs does not match number of vectors.")
        if self.dim is not None and dim != self.dim:
            raise ValueError(f"Input vector dimensionality ({dim}) does not match "
                             f"index dimensionality ({self.dim}).")
        if any(d is None and p is None for d, p in zip(doc_ids, psg_ids)):
            raise ValueError("Vector has neither document nor passage ID.")
        payload = vectors if self.quantizer is None else self.quantizer.encode(vectors)
        self._add(payload, doc_ids, psg_ids)

    # ------------------------------------------------------------------ scoring
    def _launch(self, mode: Mode, q_no: np.ndarray, id_values: np.ndarray, query_vectors: np.ndarray,
                lex: np.ndarray | None = None, alpha: float = 0.0, k: int = 0, want_ff: bool = True):
        """Integer-code the pairs and run ffx_rerank_host.  Pairs must be grouped by `q_no`
        (non-decreasing).  Returns (out dict, q_off)."""
        qv = np.ascontiguousarray(query_vectors, dtype=np.float32)
        if qv.ndim != 2 or (self.dim is not None and qv.shape[1] != self.dim):
            raise ValueError(f"Query vectors of shape {qv.shape} do not match index dimensionality {self.dim}.")
        cand = self._resolve(id_values, mode)  # every pair's id, coded in C++ on all host cores
        q_off = np.zeros(qv.shape[0] + 1, np.int64)
        np.cumsum(np.bincount(q_no, minlength=qv.shape[0]), out=q_off[1:])
        out = self._score(mode, qv, q_off, cand, lex, alpha, k, want_ff=want_ff)
        return out, q_off

    def _launch_cols(self, cols, mode: Mode, query_vectors: np.ndarray, lo: int, hi: int, alpha: float = 0.0,
                     k: int = 0, interpolate: bool = False, want_ff: bool = True):
        """ffx_rerank_host over the query blocks [lo, hi) of an integer-coded ranking: candidates
        come from the cache on the columns (every distinct id is resolved once per index, not once
        per pair and call), inputs and outputs live in recycled page-locked memory."""
        qv = np.ascontiguousarray(query_vectors[lo:hi], dtype=np.float32)
        if qv.ndim != 2 or (self.dim is not None and qv.shape[1] != self.dim):
            raise ValueError(f"Query vectors of shape {qv.shape} do not match index dimensionality {self.dim}.")
        cand = self._candidates(cols, mode)
        r0, r1 = int(cols.q_off[lo]), int(cols.q_off[hi])
        q_off = cols.q_off[lo:hi + 1] - r0 if lo else cols.q_off[:hi + 1]
        out = {}
        if want_ff:
            out["ff"] = _ffx.pinned_empty(r1 - r0, np.float32)
        if k > 0:
            out["topk_score"] = _ffx.pinned_empty((hi - lo, k), np.float32)
            out["topk_pos"] = _ffx.pinned_empty((hi - lo, k), np.int32)
        return self._score(mode, qv, q_off, cand[r0:r1], cols.score[r0:r1] if interpolate else None, alpha, k,
                           want_ff=want_ff, out=out)

    def _compute_scores(self, data: pd.DataFrame, query_vectors: np.ndarray) -> pd.DataFrame:
        """Semantic scores for the (id, q_no) rows of `data` (index/base.py:279-314).

        Returns `data` with an added `ff_score` column (float32).  One CUDA launch gathers
        the rows of every candidate, takes the dot products with the query vector and reduces
        per document by the current mode; PQ/OPQ indexes are scored from the codes."""
        n = len(data)
        if n == 0:
            return data.assign(ff_score=np.zeros(0, np.float32))
        q_no = data["q_no"].to_numpy(dtype=np.int64)
        ids = data["id"]
        order = None
        if (np.diff(q_no) < 0).any():
            order = np.argsort(q_no, kind="stable")
            q_no, ids = q_no[order], ids.iloc[order]
        out, _ = self._launch(self.mode, q_no, ids, query_vectors)
        ff = out["ff"]
        if order is not None:
            unsorted = np.empty_like(ff)
            unsorted[order] = ff
            ff = unsorted
        return data.assign(ff_score=ff)

    def _early_stopping(self, df: pd.DataFrame, query_vectors: np.ndarray, cutoff: int,
                        alpha: float, depths: Iterable[int]) -> pd.DataFrame:
        """Score in depth intervals and stop a query once its `cutoff`-th best interpolated
        score can no longer be beaten (index/base.py:316-387).  Only scored rows are returned.

        `df` must hold each query's rows in rank order (what `__call__` passes).  The whole
        walk — every interval's scoring and the stopping criterion between intervals — runs on
        the device (ffx_rerank_early_stop: one kernel launch with one CTA per query on fp32
        indexes of the common dimensions, a stream-ordered sequence of launches per depth on PQ
        indexes and the other dimensions); only beyond its limits (more than 16384 candidates per
        query or 32 depths) are the depths walked here."""
        n = len(df)
        q_no = df["q_no"].to_numpy(dtype=np.int64)
        if n and (np.diff(q_no) < 0).any():
            df = df.iloc[np.argsort(q_no, kind="stable")]
            q_no = df["q_no"].to_numpy(dtype=np.int64)
        depths = [int(d) for d in depths]
        if n == 0:
            return df.assign(ff_score=np.zeros(0, np.float32))
        lex = df["score"].to_numpy(dtype=np.float32)
        present, start, count = np.unique(q_no, return_index=True, return_counts=True)
        depth_of_row = np.arange(n) - np.repeat(start, count)
        slot_of_row = np.repeat(np.arange(len(present)), count)

        on_device = int(count.max()) <= 16384 and len({d for d in depths if d >= cutoff}) <= 32
        qv = np.ascontiguousarray(query_vectors, dtype=np.float32)[present]
        # every id coded once, whatever the number of depths.  The reference looks ids up depth by depth
        # (index/base.py:373 -> index/util.py:38-39), so an id the index does not hold only raises if
        # its row is actually scored: unknown ids are scored as a stand-in candidate and checked
        # against the scored prefixes afterwards.
        cand = self._resolve(df["id"], self.mode, missing_ok=True)
        unknown = cand < 0
        if unknown.any():
            if unknown.all():
                raise IndexError(f"ID {df['id'].iloc[0]} not found in the index.")
            cand = np.where(unknown, cand[~unknown][0], cand).astype(np.int32)
        if on_device:
            q_off = np.concatenate([[0], np.cumsum(count)]).astype(np.int64)
            try:
                out = self._early_stop(self.mode, qv, q_off, cand, lex, alpha, cutoff, depths)
            except _ffx.FFXError as e:
                if e.code != -5:  # FFX_ERR_UNSUPPORTED: beyond the device walk's limits
                    raise
                on_device = False
            else:
                ff, done_depth = out["ff"], out["scored"].astype(np.int64)
                LOGGER.info("early stopping: %s of %s rows scored", int(done_depth.sum()), n)
        if not on_device:
            ff, done_depth = self._early_stopping_walk(qv, cand, cutoff, alpha, depths, lex, start, count,
                                                       depth_of_row, slot_of_row)
        scored = depth_of_row < done_depth[slot_of_row]
        if unknown.any() and (unknown & scored).any():
            # the first one the reference would have met: earliest depth interval, then frame order
            bad = np.flatnonzero(unknown & scored)
            walked = sorted(d for d in set(depths) if d >= cutoff)
            interval = np.searchsorted(walked, depth_of_row[bad], side="right")
            first = bad[np.lexsort((bad, interval))[0]]
            raise IndexError(f"ID {df['id'].iloc[first]} not found in the index.")
        result = df.loc[scored].copy()
        result["ff_score"] = ff[scored]
        return result

    def _early_stopping_walk(self, qv, cand, cutoff, alpha, depths, lex, start, count, depth_of_row, slot_of_row):
        """Host-walked depths (index/base.py:339-385) over integer-coded pairs: per interval one
        ffx_rerank launch for the scores and one ffx_interpolate_topk launch for the criterion."""
        n = len(cand)
        ff = np.zeros(n, np.float32)
        done_depth = np.zeros(len(start), np.int64)  # rows scored so far per query
        active = np.ones(len(start), bool)
        a32, b32 = np.float32(alpha), np.float32(1 - alpha)
        lo = 0
        for hi in sorted(depths):
            if hi < cutoff:
                continue
            if lo > 0 and active.any():
                # the stopping criterion of every query still going, in one ffx_interpolate_topk call
                # over the rows scored so far: cutoff-th best interpolated score (the worst one when
                # fewer rows were scored) against alpha*lex[last] + (1-alpha)*max ff  (base.py:351-356)
                going = np.flatnonzero(active)
                sizes = done_depth[going]
                sub_off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
                rows = np.repeat(start[going] - sub_off[:-1], sizes) + np.arange(sub_off[-1])
                top = self._device().interpolate_topk_host(lex[rows], ff[rows], sub_off, alpha, cutoff,
                                                           want_int=False)["topk_score"]
                kth = top[np.arange(len(going)), np.minimum(cutoff, sizes) - 1]
                bound = a32 * lex[start[going] + sizes - 1] + b32 * np.maximum.reduceat(ff[rows], sub_off[:-1])
                active[going] = kth < bound
            LOGGER.info("depth %s: %s queries left", hi, int(active.sum()))
            take = active[slot_of_row] & (depth_of_row >= lo) & (depth_of_row < hi)
            if not take.any():
                break
            taken = np.flatnonzero(take)
            part_off = np.concatenate([[0], np.cumsum(np.bincount(slot_of_row[taken], minlength=len(start)))])
            ff[taken] = self._score(self.mode, qv, part_off.astype(np.int64), cand[taken])["ff"]
            done_depth[active] = np.minimum(count[active], hi)
            lo = hi
        return ff, done_depth

    def _query_vectors_for(self, src: pd.DataFrame):
        """Number the queries in order of appearance and encode each once (base.py:418-429)."""
        q_codes, first_row = _number_queries(src["q_id"])
        vectors = self.encode_queries(src["query"].iloc[first_row].tolist())
        return q_codes, len(first_row), vectors

    def __call__(self, ranking: Ranking, early_stopping: int | None = None,
                 early_stopping_alpha: float | None = None,
                 early_stopping_depths: Iterable[int] | None = None,
                 batch_size: int | None = None) -> Ranking:
        """Compute semantic scores for a ranking (index/base.py:389-469).

        :raises ValueError: the ranking has no queries attached, or early stopping is requested
            without alpha and depths.
        :raises IndexError: an id of the ranking is not in the index.
        :return: a ranking named "fast-forward" with `score` = semantic score.
        """
        if not ranking.has_queries:
            raise ValueError("Input ranking has no queries attached.")
        if early_stopping is not None and (early_stopping_alpha is None or early_stopping_depths is None):
            raise ValueError("Early stopping requires alpha and depths.")
        started = perf_counter()

        cols = ranking._columns() if early_stopping is None else None
        if cols is not None and cols.queries is not None:
            # integer-coded route: one launch per query batch gives the semantic scores AND the
            # per-query order (ties keep the incoming order, like the reference's stable sort)
            query_vectors = self.encode_queries(cols.queries.to_pylist())
            nq, counts = cols.nq, cols.counts()
            step = nq if batch_size is None or batch_size >= nq else max(int(batch_size), 1)
            widest = int(counts.max())
            if self._lists_are_skewed(nq, widest, len(cols)):
                result = self._ordered_on_host(cols, query_vectors, None)
                if result is not None:
                    LOGGER.info("computed scores in %s seconds", perf_counter() - started)
                    out_ranking = Ranking._from_cols(result[0], "fast-forward")
                    out_ranking._origin = _Origin(self, cols, result[1])
                    return out_ranking
            if step >= nq:
                out = self._launch_cols(cols, self.mode, query_vectors, 0, nq, k=widest)
                ff, pos, top = out["ff"], out["topk_pos"], out["topk_score"]
            else:
                ff = np.empty(len(cols), np.float32)
                pos = np.full((nq, widest), -1, np.int32)
                top = np.full((nq, widest), -np.inf, np.float32)
                for lo in range(0, nq, step):
                    hi = min(nq, lo + step)
                    w = int(counts[lo:hi].max())
                    out = self._launch_cols(cols, self.mode, query_vectors, lo, hi, k=w)
                    ff[cols.q_off[lo]:cols.q_off[hi]] = out["ff"]
                    pos[lo:hi, :w], top[lo:hi, :w] = out["topk_pos"], out["topk_score"]
            LOGGER.info("computed scores in %s seconds", perf_counter() - started)
            if not np.isnan(ff).any():  # NaN rows are dropped by Ranking: those take the generic route
                result, _, _ = cols.from_lists(pos, top, widest)
                out_ranking = Ranking._from_cols(result.drop_empty(), "fast-forward")
                out_ranking._origin = _Origin(self, cols, ff)
                return out_ranking
            frame = ranking._df[["q_id", "id", "query"]].assign(score=ff)
            return Ranking(frame, name="fast-forward", dtype=np.float32, copy=False, is_sorted=False)

        src = ranking._df
        q_codes, nq, query_vectors = self._query_vectors_for(src)
        step = nq if batch_size is None or batch_size >= nq else int(batch_size)
        dtype = src.dtypes["score"]
        work = src.assign(q_no=q_codes, orig_index=np.arange(len(src)))

        def run(part: pd.DataFrame) -> pd.DataFrame:
            if early_stopping is None:
                return self._compute_scores(part, query_vectors)
            return self._early_stopping(part, query_vectors, early_stopping,
                                        float(early_stopping_alpha), early_stopping_depths)

        # sequential query batches as in the reference; an empty trailing batch is skipped
        parts = [run(work[(q_codes >= lo) & (q_codes < lo + step)]) for lo in range(0, nq, max(step, 1))]
        result = pd.concat(parts) if parts else work.assign(ff_score=np.zeros(0, np.float32))
        frame = result[["q_id", "id", "query"]].copy()
        frame["score"] = result["ff_score"].to_numpy()
        LOGGER.info("computed scores in %s seconds", perf_counter() - started)
        return Ranking(frame, name="fast-forward", dtype=dtype, copy=False, is_sorted=False)

    def rerank(self, ranking: Ranking, alpha: float, cutoff: int | None = None) -> Ranking:
        """Fused re-ranking: the result of
        `ranking.interpolate(self(ranking), alpha).cut(cutoff)` from ONE kernel launch
        (look-up, dots, per-document reduce, interpolation and per-query top-k all on the GPU).
        The ranking's ids are resolved against the index once; the integer candidates stay
        cached on the ranking, so further calls hash no string.
        """
        if not ranking.has_queries:
            raise ValueError("Input ranking has no queries attached.")
        cols = ranking._columns()
        if cols is None or cols.queries is None:  # float64 scores, NaN queries, ...: the three-call route
            out = ranking.interpolate(self(ranking), alpha)
            return out if cutoff is None else out.cut(cutoff)
        query_vectors = self.encode_queries(cols.queries.to_pylist())
        nq = cols.nq
        widest = int(cols.counts().max())
        k = widest if cutoff is None else int(min(max(cutoff, 0), widest))
        if k == 0:
            return ranking.cut(0)
        if self._lists_are_skewed(nq, widest, len(cols)):
            result = self._ordered_on_host(cols, query_vectors, alpha)
            if result is not None:
                return Ranking._from_cols(result[0].head(k) if k < widest else result[0], ranking.name)
        # one slot more than the cut: it tells whether equal scores straddle the cut boundary
        kk = min(k + 1, widest)
        out = self._launch_cols(cols, self.mode, query_vectors, 0, nq, alpha=alpha, k=kk, interpolate=True,
                                want_ff=False)
        result, ties, straddle = cols.from_lists(out["topk_pos"], out["topk_score"], k, want_straddle=kk > k)
        if straddle is not None and straddle.any():
            # equal scores on both sides of the cut (float32 collisions; a handful of queries in
            # millions of pairs): the reference keeps the smaller ids.  Rank those queries in
            # full, order their ties by id and cut again.
            self._recut_straddling(np.flatnonzero(straddle), cols, result, query_vectors, alpha, k)
        result.order_ties_by_id(ties)  # ties inside the kept lists: ascending id, like the reference
        return Ranking._from_cols(result.drop_empty(), ranking.name)

    @staticmethod
    def _lists_are_skewed(nq: int, widest: int, n: int) -> bool:
        """The per-query lists of ffx_rerank are dense [nq, widest] matrices: fine when the lists
        are about equally long, quadratic when one query has a million candidates and ten thousand
        others have ten (the reference is O(n) there)."""
        return nq * widest > 4 * n + (1 << 20)

    def _ordered_on_host(self, cols, query_vectors, alpha):
        """Skewed list lengths: semantic scores only from the device (O(n) memory), the
        interpolation (ranking.py:319, same two roundings in float32) and the per-query order
        (stable radix sort, ffx_ranking_order) on the host.  Returns (ordered columns, ff in source
        order), or None when a NaN score needs the generic route."""
        import ctypes as C

        from fast_forward._cols import Cols

        ff = self._launch_cols(cols, self.mode, query_vectors, 0, cols.nq, k=0, want_ff=True)["ff"]
        score = ff if alpha is None else _fl32_interpolate(alpha, cols.score, ff)
        if np.isnan(score).any():
            return None
        score = np.ascontiguousarray(score, np.float32)
        block = np.repeat(np.arange(cols.nq, dtype=np.int32), cols.counts())
        order = np.empty(len(score), np.int64)
        _ffx.check(_ffx.lib().ffx_ranking_order(C.c_void_p(block.ctypes.data), C.c_void_p(score.ctypes.data), len(score),
                                                C.c_void_p(order.ctypes.data), 0))
        ordered = Cols(cols.q_keys, cols.q_off, cols.ids, cols.id_code[order], score[order], cols.queries)
        if alpha is not None:
            ordered.order_ties_by_id()
        return ordered, ff

    def _recut_straddling(self, blocks, cols, result, query_vectors, alpha, k) -> None:
        """Exact cut for the queries whose k-th and (k+1)-th interpolated scores are equal;
        overwrites their (full, k-row) blocks of `result`."""
        from fast_forward._cols import Cols

        for b in blocks.tolist():
            n_b = int(cols.q_off[b + 1] - cols.q_off[b])
            full = self._launch_cols(cols, self.mode, query_vectors, b, b + 1, alpha=alpha, k=n_b, interpolate=True,
                                     want_ff=False)
            pos, score = full["topk_pos"][0], full["topk_score"][0]
            valid = pos >= 0
            codes = cols.id_code[cols.q_off[b] + pos[valid]]
            one = Cols(cols.q_keys.slice(b, 1), np.array([0, len(codes)], np.int64), cols.ids, codes,
                       np.array(score[valid], np.float32), None)
            one.order_ties_by_id()
            at = int(result.q_off[b])
            assert int(result.q_off[b + 1]) - at == k
            result.id_code[at:at + k] = one.id_code[:k]
            result.score[at:at + k] = one.score[:k]

    # ------------------------------------------------------------------ iteration
    def batch_iter(self, batch_size: int) -> Iterator[tuple[np.ndarray, IDSequence, IDSequence]]:
        """Yield (vectors, doc ids, passage ids) batches; codes are decoded when quantized."""
        if self._quantizer is None:
            yield from self._batch_iter(batch_size)
        else:
            for codes, doc_ids, psg_ids in self._batch_iter(batch_size):
                yield self._quantizer.decode(codes), doc_ids, psg_ids

    def __iter__(self) -> Iterator[tuple[np.ndarray, "str | None", "str | None"]]:
        for vectors, doc_ids, psg_ids in self.batch_iter(2**9):
            yield from zip(vectors, doc_ids, psg_ids)
