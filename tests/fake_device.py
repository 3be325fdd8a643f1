"""A stand-in for `fast_forward._ffx.DeviceIndex` that keeps its rows in host memory and scores
with the ORACLE (oracle/ff_oracle.py) — test infrastructure for the host logic above the C ABI
(the multi-device stores: placement of documents, candidate encoding, query split, list merge)
on machines without a GPU.  Never used by the product."""

import numpy as np

import ff_oracle as fo


class FakeDeviceIndex:
    def __init__(self, dim, capacity=0, row_kind=0, device=0):
        self.dim, self.row_kind, self.device = int(dim), row_kind, device
        self.rows = np.zeros((int(capacity), self.dim), np.float32 if row_kind == 0 else np.uint8)
        self.n = 0
        self.off = self.doc_rows = None
        self.shard = (0, 0, 0, 0)
        self.calls = 0

    # ---- storage
    def __len__(self):
        return self.n

    @property
    def capacity(self):
        return len(self.rows)

    def reserve(self, capacity):
        if capacity > len(self.rows):
            grown = np.zeros((capacity, self.dim), self.rows.dtype)
            grown[:self.n] = self.rows[:self.n]
            self.rows = grown

    def stage(self, row0, rows):
        assert row0 + len(rows) <= len(self.rows)
        self.rows[row0:row0 + len(rows)] = rows
        self.n = max(self.n, row0 + len(rows))

    def read_rows(self, rows):
        return self.rows[np.asarray(rows, np.int64)].copy()

    def copy_rows_from(self, other):
        assert self.n == 0 and self.dim == other.dim and self.row_kind == other.row_kind
        self.reserve(other.n)
        self.rows[:other.n] = other.rows[:other.n]
        self.n = other.n

    def set_docs(self, doc_off, doc_rows=None):
        self.off = np.asarray(doc_off, np.int64)
        self.doc_rows = np.arange(self.off[-1]) if doc_rows is None else np.asarray(doc_rows, np.int64)

    def set_shard(self, doc_base=0, global_docs=0, row_base=0, global_rows=0):
        # the limits ffx_index_set_shard enforces
        assert global_docs <= 0x7fffffff and global_rows <= 0xffffffff
        assert doc_base + (len(self.off) - 1 if self.off is not None else 0) <= global_docs and row_base + self.n <= global_rows
        self.shard = (doc_base, global_docs, row_base, global_rows)

    def set_pq(self, codewords, R=None):
        raise NotImplementedError

    def close(self):
        pass

    # ---- scoring
    def _units(self, mode):
        if mode == fo.MODE_PASSAGE:
            return np.arange(self.n + 1), np.arange(self.n)
        if mode == fo.MODE_FIRSTP:
            return np.arange(len(self.off)), self.doc_rows[self.off[:-1]]
        return self.off, self.doc_rows

    def rerank_host(self, mode, qvecs, q_off, cand, lex=None, alpha=0.0, k=0, want_ff=True, want_int=False, out=None):
        self.calls += 1
        q_off = np.asarray(q_off, np.int64)
        cand = np.asarray(cand).astype(np.int64) & 0xffffffff
        nq = len(q_off) - 1
        base, n_glob = (self.shard[2], self.shard[3]) if mode == fo.MODE_PASSAGE else (self.shard[0], self.shard[1])
        count = self.n if mode == fo.MODE_PASSAGE else len(self.off) - 1
        if n_glob == 0:
            base = 0
            assert ((cand >= 0) & (cand < count)).all(), "candidate out of range"
        mine = (cand >= base) & (cand < base + count)
        u_off, u_rows = self._units(mode)
        pair_q = np.repeat(np.arange(nq), np.diff(q_off))
        ff_mine = fo.score_pairs(self.rows[:self.n], u_off, u_rows, pair_q[mine], cand[mine] - base,
                                 np.asarray(qvecs, np.float32), fo.MODE_PASSAGE if mode == fo.MODE_FIRSTP else mode)
        out = {} if out is None else out
        ff = np.zeros(len(cand), np.float32)
        ff[mine] = ff_mine
        if want_ff:
            if out.get("ff") is None:
                out["ff"] = np.full(len(cand), 12345.0, np.float32)  # foreign pairs: garbage on the device
            out["ff"][mine] = ff_mine
        if k > 0:
            inter = ff if lex is None else fo.interpolate_f32(np.asarray(lex, np.float32), ff, alpha)
            ts = np.full((nq, k), -np.inf, np.float32)
            tp = np.full((nq, k), -1, np.int32)
            for q in range(nq):
                lo, hi = q_off[q], q_off[q + 1]
                pos = np.flatnonzero(mine[lo:hi] & ~np.isnan(inter[lo:hi]))
                order = pos[np.argsort(-inter[lo:hi][pos], kind="stable")][:k]
                ts[q, :len(order)], tp[q, :len(order)] = inter[lo:hi][order], order
            if out.get("topk_score") is None:
                out["topk_score"], out["topk_pos"] = ts, tp
            else:
                out["topk_score"][:], out["topk_pos"][:] = ts, tp
        return out

    def rerank_early_stop_host(self, *a, **kw):
        from fast_forward import _ffx

        raise _ffx.FFXError(-5, "fake device: the host walks the depths")

    def interpolate_topk_host(self, lex, ff, q_off, alpha, k, want_int=True):
        inter = np.asarray(ff, np.float32) if lex is None else fo.interpolate_f32(np.asarray(lex, np.float32),
                                                                                  np.asarray(ff, np.float32), alpha)
        ts, tp = fo.topk_per_query(np.asarray(q_off, np.int64), inter, k) if k > 0 else (None, None)
        return {"int": inter if want_int else None, "topk_score": ts, "topk_pos": tp}

    def merge_topk_host(self, shard_scores, shard_pos):
        S, nq, k = shard_scores.shape
        out_s = np.full((nq, k), -np.inf, np.float32)
        out_p = np.full((nq, k), -1, np.int32)
        for q in range(nq):
            s, p = shard_scores[:, q].reshape(-1), shard_pos[:, q].reshape(-1)
            keep = p >= 0
            s, p = s[keep], p[keep]
            order = np.lexsort((p, -s.astype(np.float64)))[:k]
            out_s[q, :len(order)], out_p[q, :len(order)] = s[order], p[order]
        return out_s, out_p
