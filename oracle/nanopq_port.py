"""TEST INFRASTRUCTURE — CPU restatement of the third-party `nanopq` quantizers.

The reference's quantizer adapters (`src/fast_forward/quantizer/nanopq.py:26,30,41,44,94,98,
109,112`) delegate all arithmetic to the PyPI package **nanopq, pinned 0.2.1**
(`uv.lock:606-607`; spec `>=0.2.1,<0.3`, `pyproject.toml:21`).  That package is neither
vendored under `/root/reference` nor installed in this image, so this file restates its
published algorithm (Jégou et al. product quantization; Ge et al. non-parametric OPQ):

* `PQ.fit`      per-subspace k-means (`scipy.cluster.vq.kmeans2`, `minit="points"`)
* `PQ.encode`   nearest codeword per subspace (`scipy.cluster.vq.vq`)
* `PQ.decode`   `vecs[n, m*Ds:(m+1)*Ds] = codewords[m, codes[n, m], :]`
* `OPQ.fit`     alternate PQ fitting and an orthogonal Procrustes update of `R`
* `OPQ.encode`  `PQ.encode(vecs @ R)`;  `OPQ.decode`  `PQ.decode(codes) @ R.T`

PARITY UNPINNED for decode *values*: the reference's own tests at this boundary
(`tests/test_quantizer.py:17-46`, `tests/test_index.py:389-399`) only check shapes, dtypes
and serialise/re-encode round trips.  This port passes those tests when injected as the
`nanopq` module (see `oracle/ref_shim.py`); numeric parity of the GPU ADC path is anchored
on the algebraic identity `q . (dec(c) R^T) == sum_m LUT[m][c_m]` instead.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline leg may import this.
"""

import numpy as np
from scipy.cluster.vq import kmeans2, vq


def _code_dtype(Ks: int):
    if Ks <= 2**8:
        return np.uint8
    if Ks <= 2**16:
        return np.uint16
    return np.uint32


class PQ:
    """Product quantizer with nanopq's attribute names (M, Ks, Ds, codewords, ...)."""

    def __init__(self, M, Ks=256, metric="l2", verbose=True):
        assert 0 < Ks <= 2**32
        assert metric in ("l2", "dot")
        self.M, self.Ks, self.metric, self.verbose = M, Ks, metric, verbose
        self.code_dtype = _code_dtype(Ks)
        self.codewords = None
        self.Ds = None

    def __eq__(self, other):
        if not isinstance(other, PQ):
            return False
        same_cfg = (self.M, self.Ks, self.metric, self.verbose, self.Ds) == (
            other.M, other.Ks, other.metric, other.verbose, other.Ds)
        return same_cfg and np.array_equal(self.codewords, other.codewords)

    def fit(self, vecs, iter=20, seed=123, minit="points"):
        assert vecs.dtype == np.float32 and vecs.ndim == 2
        N, D = vecs.shape
        assert self.Ks < N, "the number of training vectors should be more than Ks"
        assert D % self.M == 0, "input dimension must be dividable by M"
        self.Ds = D // self.M
        np.random.seed(seed)
        self.codewords = np.zeros((self.M, self.Ks, self.Ds), dtype=np.float32)
        for m in range(self.M):
            sub = vecs[:, m * self.Ds:(m + 1) * self.Ds]
            self.codewords[m], _ = kmeans2(sub, self.Ks, iter=iter, minit=minit)
        return self

    def encode(self, vecs):
        assert vecs.dtype == np.float32 and vecs.ndim == 2
        N, D = vecs.shape
        assert D == self.Ds * self.M
        codes = np.empty((N, self.M), dtype=self.code_dtype)
        for m in range(self.M):
            sub = vecs[:, m * self.Ds:(m + 1) * self.Ds]
            codes[:, m], _ = vq(sub, self.codewords[m])
        return codes

    def decode(self, codes):
        assert codes.ndim == 2 and codes.dtype == self.code_dtype
        N, M = codes.shape
        assert M == self.M
        vecs = np.empty((N, self.Ds * self.M), dtype=np.float32)
        for m in range(self.M):
            vecs[:, m * self.Ds:(m + 1) * self.Ds] = self.codewords[m][codes[:, m], :]
        return vecs


class OPQ:
    """Optimised PQ (non-parametric): a learned rotation R followed by PQ."""

    def __init__(self, M, Ks=256, metric="l2", verbose=True):
        self.pq = PQ(M, Ks, metric=metric, verbose=verbose)
        self.R = None

    def __eq__(self, other):
        return isinstance(other, OPQ) and self.pq == other.pq and np.array_equal(self.R, other.R)

    @property
    def M(self):
        return self.pq.M

    @property
    def Ks(self):
        return self.pq.Ks

    @property
    def verbose(self):
        return self.pq.verbose

    @property
    def code_dtype(self):
        return self.pq.code_dtype

    @property
    def codewords(self):
        return self.pq.codewords

    @property
    def Ds(self):
        return self.pq.Ds

    def fit(self, vecs, parametric_init=False, pq_iter=20, rotation_iter=10, seed=123,
            minit="points"):
        assert vecs.dtype == np.float32 and vecs.ndim == 2
        _, D = vecs.shape
        self.R = np.eye(D, dtype=np.float32)
        for i in range(rotation_iter):
            X = vecs @ self.R
            last = i == rotation_iter - 1
            pq_tmp = PQ(self.M, self.Ks, metric=self.pq.metric, verbose=self.pq.verbose)
            pq_tmp.fit(X, iter=pq_iter if last else 1, seed=seed, minit=minit)
            X_hat = pq_tmp.decode(pq_tmp.encode(X))
            U, _, Vt = np.linalg.svd(vecs.T @ X_hat)
            if last:
                self.pq = pq_tmp
                break
            self.R = (U @ Vt).astype(np.float32)
        return self

    def rotate(self, vecs):
        return vecs @ self.R

    def encode(self, vecs):
        return self.pq.encode(self.rotate(vecs))

    def decode(self, codes):
        return self.pq.decode(codes) @ self.R.T
