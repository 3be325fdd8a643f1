"""TEST INFRASTRUCTURE — CPU oracle for the Fast-Forward re-ranking hot path (numpy).

This is a restatement, in plain numpy, of what the reference (mrjleo/fast-forward-indexes
v0.8.0, pure Python) computes on the path

    look-up by id -> q.p dot products -> reduce per doc by Mode -> interpolate -> top-k

Every function cites the reference `file:line` it follows (paths relative to
`/root/reference/src/fast_forward/`).  It is the CHECKER for the CUDA path and the
`cpu_baseline` of `bench.py`; it is never the product.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / `--impl reference` legs may
import it.

Parity pin: `oracle/gen_golden.py` runs the UNMODIFIED reference (via `oracle/ref_shim.py`)
on seeded inputs and on the reference's own known-answer fixtures
(`tests/test_index.py:19-46,135-200,273-349`, `tests/test_ranking.py:116-121,157-188`) and
commits the outputs under `tests/golden/`; `tests/test_oracle.py` checks this file against
them bit-for-bit.  The PQ/OPQ decode values are "parity unpinned" (see `nanopq_port.py`).

Arithmetic facts the oracle relies on (verified against numpy 2.3 / pandas 3.0 here):
  N1  `np.sum(a * b, axis=1)` in fp32 = products rounded to fp32, then numpy's pairwise
      summation tree (8 strided accumulators per <=128-element leaf); restated explicitly
      in `oracle/ff_oracle_c.c` and followed lane-for-lane by the CUDA kernel.
  N2  pandas fp32 `groupby.mean` = fp32 Kahan-compensated sum in row order / fp32(count).
  N3  `alpha * s + (1 - alpha) * f` on fp32 Series = fl32(fl32(alpha)*s) + fl32(fl32(1-alpha)*f),
      with `1 - alpha` formed in double first; no FMA.
"""

from __future__ import annotations

import numpy as np

# index/base.py:18-24
MODE_PASSAGE, MODE_MAXP, MODE_FIRSTP, MODE_AVEP = 1, 2, 3, 4


# --------------------------------------------------------------------------------------
# storage + id mapping
# --------------------------------------------------------------------------------------
class OracleIndex:
    """Row store + the two id maps of `InMemoryIndex` (index/memory.py:20-140).

    Row i is the i-th vector ever added; `doc_rows[id]` lists a document's rows in
    insertion order (memory.py:86-88), `psg_row[id]` is a passage's single row
    (memory.py:90-95).  Unlike the reference we keep one growing array in the dtype of the
    first `add` (the reference's float64 growth chunks, memory.py:106, are a quirk we do
    not restate: `InMemoryIndex(init_size=N)` never triggers it).
    """

    def __init__(self, mode: int = MODE_MAXP):
        self.mode = mode
        self.vectors: np.ndarray | None = None
        self.doc_rows: dict[str, list[int]] = {}
        self.psg_row: dict[str, int] = {}

    def __len__(self):
        return 0 if self.vectors is None else self.vectors.shape[0]

    def add(self, vectors, doc_ids=None, psg_ids=None):
        # validation: index/base.py:233-250
        n, dim = vectors.shape
        doc_ids = [None] * n if doc_ids is None else list(doc_ids)
        psg_ids = [None] * n if psg_ids is None else list(psg_ids)
        if not len(doc_ids) == len(psg_ids) == n:
            raise ValueError("Number of IDs does not match number of vectors.")
        if self.vectors is not None and dim != self.vectors.shape[1]:
            raise ValueError("Input vector dimensionality does not match index dimensionality.")
        for d, p in zip(doc_ids, psg_ids):
            if d is None and p is None:
                raise ValueError("Vector has neither document nor passage ID.")
        base = len(self)
        for i, p in enumerate(psg_ids):  # memory.py:90-95 (checked before any mutation here)
            if p is not None and p in self.psg_row:
                raise RuntimeError(f"Passage ID {p} already exists.")
        for i, d in enumerate(doc_ids):
            if d is not None:
                self.doc_rows.setdefault(d, []).append(base + i)
        for i, p in enumerate(psg_ids):
            if p is not None:
                self.psg_row[p] = base + i
        self.vectors = vectors.copy() if self.vectors is None else np.concatenate(
            [self.vectors, vectors.astype(self.vectors.dtype)])

    def rows_for(self, id_: str, mode: int | None = None) -> list[int]:
        """index/util.py:29-41 — the id -> rows rule per Mode; IndexError when unknown."""
        mode = self.mode if mode is None else mode
        if mode in (MODE_MAXP, MODE_AVEP):
            rows = self.doc_rows.get(id_, [])
        elif mode == MODE_FIRSTP:
            rows = self.doc_rows.get(id_, [])[:1]
        else:
            r = self.psg_row.get(id_)
            rows = [] if r is None else [r]
        if len(rows) == 0:
            raise IndexError(f"ID {id_} not found in the index.")
        return rows

    def get_vectors(self, ids, mode: int | None = None):
        """index/util.py:84-99 — gathered copies + one id per returned row."""
        rows, out_ids = [], []
        for id_ in ids:
            r = self.rows_for(id_, mode)
            rows.extend(r)
            out_ids.extend([id_] * len(r))
        if not rows:
            return np.array([]), []
        return self.vectors[rows], out_ids

    def csr(self, ids, mode: int | None = None):
        """Integer-coded view for the C-ABI tests: (offsets[len(ids)+1], rows[])."""
        off = [0]
        rows: list[int] = []
        for id_ in ids:
            rows.extend(self.rows_for(id_, mode))
            off.append(len(rows))
        return np.asarray(off, dtype=np.int64), np.asarray(rows, dtype=np.int64)


# --------------------------------------------------------------------------------------
# scoring
# --------------------------------------------------------------------------------------
def kahan_mean_f32(values: np.ndarray, seg_off: np.ndarray) -> np.ndarray:
    """N2: pandas fp32 groupby-mean, vectorised over segments (row order inside each)."""
    values = np.asarray(values, dtype=np.float32)
    nseg = len(seg_off) - 1
    cnt = (seg_off[1:] - seg_off[:-1]).astype(np.int64)
    s = np.zeros(nseg, np.float32)
    c = np.zeros(nseg, np.float32)
    for j in range(int(cnt.max()) if nseg else 0):
        live = np.nonzero(cnt > j)[0]
        x = values[seg_off[live] + j]
        y = x - c[live]
        t = s[live] + y
        comp = (t - s[live]) - y
        comp[comp != comp] = 0  # pandas resets a NaN compensation (inf inputs)
        c[live] = comp
        s[live] = t
    return s / cnt.astype(np.float32)


def reduce_segments(row_scores: np.ndarray, seg_off: np.ndarray, mode: int) -> np.ndarray:
    """index/base.py:305-312 — groupby(id, q_no).aggregate(max | mean | first)."""
    seg_off = np.asarray(seg_off, dtype=np.int64)
    if len(seg_off) <= 1:
        return row_scores[:0].copy()
    if mode == MODE_MAXP:
        return np.maximum.reduceat(row_scores, seg_off[:-1])
    if mode == MODE_AVEP:
        if row_scores.dtype == np.float32:
            return kahan_mean_f32(row_scores, seg_off)
        # integer / float64 columns: pandas computes the mean in float64 (Kahan, float64)
        out = np.empty(len(seg_off) - 1, np.float64)
        for i in range(len(out)):
            s = c = 0.0
            for x in row_scores[seg_off[i]:seg_off[i + 1]].astype(np.float64):
                y = x - c
                t = s + y
                c = (t - s) - y
                s = t
            out[i] = s / (seg_off[i + 1] - seg_off[i])
        return out
    return row_scores[seg_off[:-1]]


def score_pairs(vectors: np.ndarray, unit_off: np.ndarray, unit_rows: np.ndarray,
                pair_q: np.ndarray, pair_unit: np.ndarray, qvecs: np.ndarray, mode: int,
                chunk_pairs: int = 1 << 15) -> np.ndarray:
    """`Index._compute_scores` (index/base.py:279-314) on integer-coded pairs.

    `unit_off/unit_rows` is the CSR unit -> rows (already mode-resolved: one row for
    PASSAGE/FIRSTP); pair i scores query `pair_q[i]` against unit `pair_unit[i]`.
    Steps, as in the reference: expand to one line per (pair, passage) (base.py:296-298),
    gather `q_reps`/`d_reps` copies (:301-302), `np.sum(q_reps * d_reps, axis=1)` (:303),
    segmented reduce (:306-312).  Chunked over pairs like `batch_size` (:445-459) so the
    temporaries stay bounded; the result does not depend on the chunking.
    """
    unit_off = np.asarray(unit_off, dtype=np.int64)
    unit_rows = np.asarray(unit_rows, dtype=np.int64)
    pair_q = np.asarray(pair_q, dtype=np.int64)
    pair_unit = np.asarray(pair_unit, dtype=np.int64)
    n = len(pair_q)
    outs = []
    for lo in range(0, n, chunk_pairs):
        hi = min(n, lo + chunk_pairs)
        u = pair_unit[lo:hi]
        cnt = unit_off[u + 1] - unit_off[u]
        seg = np.concatenate([[0], np.cumsum(cnt)])
        # line -> (pair, k-th row of the pair's unit)
        line_pair = np.repeat(np.arange(hi - lo), cnt)
        line_k = np.arange(seg[-1]) - seg[line_pair]
        line_row = unit_rows[unit_off[u][line_pair] + line_k]
        q_reps = qvecs[pair_q[lo:hi][line_pair]]
        d_reps = vectors[line_row]
        line_score = np.sum(q_reps * d_reps, axis=1)
        outs.append(reduce_segments(line_score, seg, mode))
    if not outs:
        return np.zeros(0, np.result_type(vectors.dtype, qvecs.dtype))
    return np.concatenate(outs)


def interpolate_f32(score_self: np.ndarray, score_other: np.ndarray, alpha: float) -> np.ndarray:
    """ranking.py:319 (N3).  `self` is the ranking `.interpolate` is called on."""
    a = np.float32(alpha)
    b = np.float32(1 - alpha)
    return (a * score_self.astype(np.float32)).astype(np.float32) + \
        (b * score_other.astype(np.float32)).astype(np.float32)


def early_stopping_depth(q_off: np.ndarray, lex: np.ndarray, ff: np.ndarray, alpha: float, cutoff: int,
                         depths) -> np.ndarray:
    """`Index._early_stopping` (index/base.py:316-387) on integer-coded pairs: how many rows of
    every query's block (rows in rank order) the reference scores.

    A row's `ff_score` does not depend on early stopping, so the walk is restated over the
    full score vector `ff`: depths ascending, those below `cutoff` skipped (:341-343); from
    the second interval on a query goes on only while the `cutoff`-th best interpolated score
    so far (`nlargest(cutoff).iat[-1]`, the worst one when fewer rows were scored) is below
    `alpha * score(last scored row) + (1 - alpha) * max(ff_score)` (:351-356; fp32 with N3
    arithmetic under numpy 2 scalar promotion); the loop ends at the first interval that adds
    no row for any query still going (:365-366).
    """
    q_off = np.asarray(q_off, dtype=np.int64)
    lex = np.asarray(lex, dtype=np.float32)
    ff = np.asarray(ff, dtype=np.float32)
    a32, b32 = np.float32(alpha), np.float32(1 - alpha)
    nq = len(q_off) - 1
    size = np.diff(q_off)
    done = np.zeros(nq, np.int64)
    going = np.ones(nq, bool)
    a = 0
    for b in sorted(depths):
        if b < cutoff:
            continue
        if a > 0:
            for q in range(nq):
                lo, e = q_off[q], q_off[q] + done[q]
                inter = a32 * lex[lo:e] + b32 * ff[lo:e]
                kth = np.sort(inter)[::-1][:cutoff][-1]
                bound = a32 * lex[e - 1] + b32 * ff[lo:e].max()
                going[q] = kth < bound
        new_done = np.where(going, np.maximum(done, np.minimum(size, b)), done)
        if a >= b or not (new_done > done).any():
            break
        done = new_done
        a = b
    return done


def topk_per_query(q_off: np.ndarray, scores: np.ndarray, k: int):
    """ranking.py:115-117 + :285-291 on one query's block of pairs: stable sort by score
    DESC (ties keep position order), keep the first k.  Returns per query (scores, pos)
    padded with (-inf, -1) when the query has fewer than k candidates.
    """
    nq = len(q_off) - 1
    out_s = np.full((nq, k), -np.inf, np.float32)
    out_p = np.full((nq, k), -1, np.int32)
    for q in range(nq):
        s = scores[q_off[q]:q_off[q + 1]]
        order = np.argsort(-s, kind="stable")[:k]
        out_s[q, :len(order)] = s[order]
        out_p[q, :len(order)] = order
    return out_s, out_p


# --------------------------------------------------------------------------------------
# string-level front (what Index.__call__ / Ranking do with the DataFrame)
# --------------------------------------------------------------------------------------
def sort_ranking(q_ids, ids, scores):
    """ranking.py:115-117 — order rows by q_id DESC (as strings) then score DESC, stable.
    Returns the permutation."""
    q_ids = np.asarray(q_ids, dtype=object)
    _, q_code = np.unique(q_ids.astype(str), return_inverse=True)
    scores = np.asarray(scores)
    return np.lexsort((-scores.astype(np.float64), -q_code))


def call_index(index: OracleIndex, q_ids, ids, query_vectors_by_qid: dict, mode: int | None = None):
    """`Index.__call__` (index/base.py:389-469) for an already sorted ranking frame given
    as parallel arrays; `query_vectors_by_qid` plays the encoder.  Returns ff_score per row
    in input row order (the caller applies `sort_ranking`)."""
    mode = index.mode if mode is None else mode
    q_ids = [str(q) for q in q_ids]
    ids = [str(i) for i in ids]
    # q_no in order of appearance (base.py:418-422)
    q_no_of: dict[str, int] = {}
    for q in q_ids:
        q_no_of.setdefault(q, len(q_no_of))
    qvecs = np.stack([np.asarray(query_vectors_by_qid[q]) for q in q_no_of])
    uniq: dict[str, int] = {}
    for i in ids:
        uniq.setdefault(i, len(uniq))
    unit_off, unit_rows = index.csr(list(uniq), mode)  # raises IndexError first (util.py:38-39)
    pair_q = np.array([q_no_of[q] for q in q_ids], dtype=np.int64)
    pair_unit = np.array([uniq[i] for i in ids], dtype=np.int64)
    return score_pairs(index.vectors, unit_off, unit_rows, pair_q, pair_unit, qvecs, mode)


def interpolate_rankings(a: dict, b: dict, alpha: float):
    """`Ranking.interpolate` (ranking.py:293-326) on {(q_id, id): score} maps: outer join,
    missing -> 0, N3 arithmetic; rows come out in ascending (q_id, id) key order (the
    outer merge sorts keys) and are then re-sorted by `sort_ranking`."""
    keys = sorted(set(a) | set(b))
    sa = np.array([a.get(k, 0.0) for k in keys], dtype=np.float32)
    sb = np.array([b.get(k, 0.0) for k in keys], dtype=np.float32)
    s = interpolate_f32(sa, sb, alpha)
    q_ids = [k[0] for k in keys]
    ids = [k[1] for k in keys]
    order = sort_ranking(q_ids, ids, s)
    return [q_ids[i] for i in order], [ids[i] for i in order], s[order]


def cut(q_ids, ids, scores, k: int):
    """`Ranking.cut` (ranking.py:279-291): first k rows of every q_id group, frame order kept."""
    seen: dict[str, int] = {}
    keep = []
    for i, q in enumerate(q_ids):
        c = seen.get(q, 0)
        if c < k:
            keep.append(i)
        seen[q] = c + 1
    return [q_ids[i] for i in keep], [ids[i] for i in keep], np.asarray(scores)[keep]


# --------------------------------------------------------------------------------------
# quantizers (nanopq 0.2.1 semantics; see nanopq_port.py)
# --------------------------------------------------------------------------------------
def pq_decode(codes: np.ndarray, codewords: np.ndarray) -> np.ndarray:
    """quantizer/nanopq.py:43-44 -> nanopq.PQ.decode."""
    M, _, Ds = codewords.shape
    out = np.empty((codes.shape[0], M * Ds), np.float32)
    for m in range(M):
        out[:, m * Ds:(m + 1) * Ds] = codewords[m][codes[:, m], :]
    return out


def opq_decode(codes: np.ndarray, codewords: np.ndarray, R: np.ndarray) -> np.ndarray:
    """quantizer/nanopq.py:111-112 -> nanopq.OPQ.decode = PQ.decode(codes) @ R.T."""
    return pq_decode(codes, codewords) @ R.T


def score_pairs_quantized(codes, codewords, R, unit_off, unit_rows, pair_q, pair_unit, qvecs,
                          mode, chunk_pairs: int = 1 << 13):
    """index/base.py:291-293 then :296-314: decode the gathered codes, then score."""
    unit_off = np.asarray(unit_off, dtype=np.int64)
    unit_rows = np.asarray(unit_rows, dtype=np.int64)
    used = np.unique(unit_rows)
    remap = np.full(codes.shape[0], -1, np.int64)
    remap[used] = np.arange(len(used))
    dec = pq_decode(codes[used], codewords) if R is None else opq_decode(codes[used], codewords, R)
    return score_pairs(dec, unit_off, remap[unit_rows], pair_q, pair_unit, qvecs, mode,
                       chunk_pairs=chunk_pairs)
