"""Product / optimised-product codebooks with nanopq-compatible state.

The reference delegates to the PyPI package nanopq 0.2.1 (quantizer/nanopq.py:26,94), which
is not a dependency here.  This module keeps nanopq's state layout — `codewords[M, Ks, Ds]`
float32, rotation `R[D, D]`, code dtype by `Ks` — and its algorithm (per-subspace k-means,
nearest-codeword encoding, OPQ = alternating PQ fit / orthogonal Procrustes), so index files
and serialised quantizers are interchangeable with the reference.  Training and encoding are
offline index-build steps; scoring never decodes (see ffx_adc_xor.cuh).

By default they run on the host (scipy), like the reference.  With `device=<cuda ordinal>` the
expensive pieces — the Lloyd iterations of `fit`, the nearest-codeword search of `encode` and
the two matrix products of every OPQ rotation round (`vecs @ R`, `vecs.T @ X_hat`; the D x D SVD
stays in numpy) — run on the GPU instead (libffx `ffx_pq_kmeans` / `ffx_pq_encode` / `ffx_sgemm`,
csrc/ffx_pq_build.cuh): same algorithm and same initial centroids (the `minit="points"` draws
of kmeans2 are repeated on the host), direct fp32 distances.  That is an explicit choice, not a
fallback: with `device` set and no CUDA device the calls raise.
"""

from __future__ import annotations

import numpy as np
from scipy.cluster.vq import kmeans2, vq


def code_dtype_for(Ks: int) -> type:
    return np.uint8 if Ks <= 1 << 8 else (np.uint16 if Ks <= 1 << 16 else np.uint32)


class Codebook:
    """M subspaces x Ks codewords; optional learned rotation (OPQ)."""

    def __init__(self, M: int, Ks: int, metric: str, verbose: bool, rotated: bool,
                 device: int | None = None) -> None:
        if not 0 < Ks <= 1 << 32:
            raise ValueError("Ks out of range")
        if metric not in ("l2", "dot"):
            raise ValueError("metric must be 'l2' or 'dot'")
        self.M, self.Ks, self.metric, self.verbose, self.rotated = M, Ks, metric, verbose, rotated
        self.device = device
        self.Ds: int | None = None
        self.codewords: np.ndarray | None = None
        self.R: np.ndarray | None = None

    @property
    def code_dtype(self):
        return code_dtype_for(self.Ks)

    # ---- plain PQ pieces ----------------------------------------------------------
    def __eq__(self, other: object) -> bool:
        """Same configuration and the same trained tables (what nanopq's `PQ.__eq__` compares)."""
        if not isinstance(other, Codebook):
            return NotImplemented
        same = (self.M, self.Ks, self.metric, self.verbose, self.rotated, self.Ds) == \
               (other.M, other.Ks, other.metric, other.verbose, other.rotated, other.Ds)
        for mine, theirs in ((self.codewords, other.codewords), (self.R, other.R)):
            same = same and (mine is None) == (theirs is None) and (mine is None or np.array_equal(mine, theirs))
        return bool(same)

    __hash__ = None

    def _split(self, vecs: np.ndarray):
        for m in range(self.M):
            yield m, vecs[:, m * self.Ds:(m + 1) * self.Ds]

    def _fit_pq(self, vecs: np.ndarray, iters: int, seed: int, minit: str) -> np.ndarray:
        np.random.seed(seed)  # kmeans2(minit="points") draws from the global generator
        if self.device is not None:
            if minit != "points":
                raise ValueError("device k-means supports minit='points' (the nanopq default) only")
            if self.Ks > 256:
                raise ValueError("device k-means needs Ks <= 256")
            from fast_forward import _ffx

            # the draws kmeans2 would make: Ks distinct training points per subspace, in order
            init = np.stack([sub[np.random.mtrand._rand.choice(vecs.shape[0], size=self.Ks, replace=False)]
                             for _, sub in self._split(vecs)]).astype(np.float32)
            return _ffx.pq_kmeans(vecs, init, iters, self.device)
        words = np.zeros((self.M, self.Ks, self.Ds), np.float32)
        for m, sub in self._split(vecs):
            words[m], _ = kmeans2(sub, self.Ks, iter=iters, minit=minit)
        return words

    def _assign(self, vecs: np.ndarray, words: np.ndarray) -> np.ndarray:
        if self.device is not None and self.Ks <= 256:
            from fast_forward import _ffx

            return _ffx.pq_encode(vecs, words, self.device)
        codes = np.empty((vecs.shape[0], self.M), self.code_dtype)
        for m, sub in self._split(vecs):
            codes[:, m], _ = vq(sub, words[m])
        return codes

    @staticmethod
    def _lookup(codes: np.ndarray, words: np.ndarray) -> np.ndarray:
        M, _, Ds = words.shape
        out = np.empty((codes.shape[0], M * Ds), np.float32)
        for m in range(M):
            out[:, m * Ds:(m + 1) * Ds] = words[m][codes[:, m]]
        return out

    # ---- public -----------------------------------------------------------------------
    def fit(self, vecs: np.ndarray, iter: int = 20, seed: int = 123, minit: str = "points",
            pq_iter: int = 20, rotation_iter: int = 10, parametric_init: bool = False) -> None:
        vecs = self._check(vecs, train=True)
        if not self.rotated:
            self.codewords = self._fit_pq(vecs, iter, seed, minit)
            return
        D = vecs.shape[1]
        R = np.eye(D, dtype=np.float32)

        def product(a, b, trans_a=False):
            """The two O(N D^2) products of a round: on the GPU with `device` (ffx_sgemm), else numpy."""
            if self.device is None:
                return (a.T if trans_a else a) @ b
            from fast_forward import _ffx

            return _ffx.sgemm(a, b, trans_a, self.device)

        for it in range(rotation_iter):
            last = it == rotation_iter - 1
            X = product(vecs, R)
            words = self._fit_pq(X, pq_iter if last else 1, seed, minit)
            if last:
                self.codewords, self.R = words, R
                return
            X_hat = self._lookup(self._assign(X, words), words)
            U, _, Vt = np.linalg.svd(product(vecs, X_hat, trans_a=True))  # orthogonal Procrustes (D x D, host)
            R = (U @ Vt).astype(np.float32)
        self.codewords, self.R = self._fit_pq(product(vecs, R), pq_iter, seed, minit), R

    def _check(self, vecs: np.ndarray, train: bool = False) -> np.ndarray:
        if vecs.ndim != 2 or vecs.dtype != np.float32:
            raise AssertionError("expected a 2-d float32 array")
        if train:
            if not self.Ks < vecs.shape[0]:
                raise AssertionError("the number of training vectors should be more than Ks")
            if vecs.shape[1] % self.M:
                raise AssertionError("input dimension must be dividable by M")
            self.Ds = vecs.shape[1] // self.M
        elif vecs.shape[1] != self.Ds * self.M:
            raise AssertionError("input dimension does not match the codebook")
        return vecs

    def encode(self, vecs: np.ndarray) -> np.ndarray:
        vecs = self._check(vecs)
        return self._assign(vecs @ self.R if self.R is not None else vecs, self.codewords)

    def decode(self, codes: np.ndarray) -> np.ndarray:
        if codes.ndim != 2 or codes.shape[1] != self.M or codes.dtype != self.code_dtype:
            raise AssertionError("codes have the wrong shape or dtype")
        flat = self._lookup(codes, self.codewords)
        return flat @ self.R.T if self.R is not None else flat
