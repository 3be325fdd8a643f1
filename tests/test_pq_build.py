"""GPU build side of product quantization (ffx_pq_encode / ffx_pq_kmeans) against float64
numpy restatements of what nanopq delegates to scipy (vq: nearest codeword per subspace;
kmeans2: Lloyd rounds, empty clusters keep their centroid).  Tolerances: an encode may pick
another codeword only when its squared distance is within 1e-5 relative of the minimum (fp32
rounding of near ties); k-means centroids within 1e-4 on data without near ties."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ffx():
    import __graft_entry__ as g

    g.build()
    from fast_forward import _ffx

    assert _ffx.device_count() >= 1
    return _ffx


def sq_dists(x, cw, m):
    Ds = cw.shape[2]
    sub = x[:, m * Ds:(m + 1) * Ds].astype(np.float64)
    return ((sub[:, None, :] - cw[m].astype(np.float64)[None, :, :]) ** 2).sum(-1)


@pytest.mark.parametrize("M,Ks,Ds", [(96, 256, 8), (8, 256, 96), (16, 100, 4), (12, 37, 16), (6, 256, 32), (5, 9, 3)])
def test_encode_picks_the_nearest_codeword(ffx, M, Ks, Ds):
    rng = np.random.default_rng(M * Ds)
    n = 3000
    x = rng.standard_normal((n, M * Ds)).astype(np.float32)
    cw = rng.standard_normal((M, Ks, Ds)).astype(np.float32)
    codes = ffx.pq_encode(x, cw)
    assert codes.shape == (n, M) and codes.dtype == np.uint8 and codes.max() < Ks
    exact = 0
    for m in range(M):
        d = sq_dists(x, cw, m)
        best = d.min(axis=1)
        got = d[np.arange(n), codes[:, m]]
        assert (got <= best * (1 + 1e-5) + 1e-6).all()
        exact += int((codes[:, m] == d.argmin(axis=1)).sum())
    assert exact >= 0.999 * n * M
    # duplicates of a codeword: ties go to the lower index
    cw[:, 5] = cw[:, 2]
    x[:10, :] = np.tile(cw[:, 2].reshape(-1), (10, 1))
    assert (ffx.pq_encode(x[:10], cw) == 2).all()


def lloyd64(x, init, iters):
    cw = init.astype(np.float64).copy()
    M, Ks, Ds = cw.shape
    for _ in range(iters):
        for m in range(M):
            sub = x[:, m * Ds:(m + 1) * Ds].astype(np.float64)
            label = ((sub[:, None, :] - cw[m][None]) ** 2).sum(-1).argmin(1)
            for k in range(Ks):
                members = sub[label == k]
                if len(members):
                    cw[m, k] = members.mean(0)
    return cw


@pytest.mark.parametrize("M,Ks,Ds", [(4, 16, 8), (3, 10, 4), (2, 7, 5)])
def test_kmeans_matches_float64_lloyd(ffx, M, Ks, Ds):
    rng = np.random.default_rng(Ks)
    centers = rng.standard_normal((M, Ks - 1, Ds)) * 20  # well separated blobs, one cluster short
    n = 6000
    which = rng.integers(0, Ks - 1, (n, M))
    x = np.concatenate([centers[m][which[:, m]] + rng.standard_normal((n, Ds)) for m in range(M)], axis=1).astype(np.float32)
    init = np.stack([np.concatenate([centers[m] + rng.standard_normal((Ks - 1, Ds)), np.full((1, Ds), 1e4)])
                     for m in range(M)]).astype(np.float32)  # the last codeword never gets a member
    for iters in (1, 4):
        got = ffx.pq_kmeans(x, init, iters)
        want = lloyd64(x, init, iters)
        assert np.allclose(got, want, rtol=1e-4, atol=1e-3)
        assert (got[:, -1] == init[:, -1]).all()
    assert (ffx.pq_kmeans(x, init, 0) == init).all()


@pytest.mark.parametrize("cls_name", ["NanoPQ", "NanoOPQ"])
def test_quantizer_on_device_is_as_good_as_on_host(ffx, cls_name):
    """Same algorithm, same initial centroids: the GPU-built quantizer reconstructs as well as
    the scipy-built one (and mostly picks the same codes)."""
    from fast_forward import quantizer as Q

    cls = getattr(Q, cls_name)
    rng = np.random.default_rng(3)
    x = rng.standard_normal((4000, 64)).astype(np.float32)
    kw = {"pq_iter": 5, "rotation_iter": 3} if cls_name == "NanoOPQ" else {"iter": 5}
    host, dev = cls(8, 32), cls(8, 32, device=0)
    host.fit(x, **kw)
    dev.fit(x, **kw)
    assert dev.dims == host.dims == (64, 8) and dev.dtype == np.uint8
    y = rng.standard_normal((1000, 64)).astype(np.float32)
    err = [float(((q.decode(q.encode(y)) - y) ** 2).mean()) for q in (host, dev)]
    assert err[1] <= err[0] * 1.03
    # encoding with the SAME codebook: device == host except on fp32 near-ties
    dev2 = cls.deserialize(*host.serialize())
    dev2._book.device = 0
    assert (dev2.encode(y) == host.encode(y)).mean() > 0.995
    meta, attrs, data = dev.serialize()
    assert cls.deserialize(meta, attrs, data) == dev


def test_limits(ffx):
    x = np.zeros((10, 8), np.float32)
    with pytest.raises(ValueError):
        ffx.pq_encode(x, np.zeros((2, 4, 3), np.float32))  # 2*3 != 8
    with pytest.raises(ffx.FFXError):
        ffx.pq_encode(np.zeros((10, 600), np.float32), np.zeros((2, 300, 300), np.float32))  # Ks > 256


@pytest.mark.parametrize("m,n,k", [(1000, 768, 768), (768, 768, 50_000), (70, 33, 17), (1, 5, 3), (130, 64, 4100)])
@pytest.mark.parametrize("trans_a", [False, True])
def test_sgemm_against_float64(ffx, m, n, k, trans_a):
    """ffx_sgemm — the two products of an OPQ rotation round (`vecs @ R`, `vecs.T @ X_hat`,
    quantizer/nanopq.py:94-98 -> nanopq OPQ.fit) — against float64 numpy: fp32 FMA accumulation in
    k-slices, so the error stays within a few fp32 ulps of |a||b| even for k = 50 000; repeated calls
    are bit-identical (fixed slice order)."""
    rng = np.random.default_rng(m + n + k)
    a = rng.standard_normal((k, m) if trans_a else (m, k)).astype(np.float32)
    b = rng.standard_normal((k, n)).astype(np.float32)
    got = ffx.sgemm(a, b, trans_a)
    want = (a.T if trans_a else a).astype(np.float64) @ b.astype(np.float64)
    scale = np.sqrt(k)  # |row of a| |column of b| ~ k for unit-variance entries; rounding errors add up like sqrt(k)
    assert got.shape == want.shape and np.abs(got - want).max() <= 4e-6 * scale * max(1.0, np.sqrt(k) / 16)
    assert np.array_equal(got, ffx.sgemm(a, b, trans_a))


def test_device_opq_rotation_is_orthogonal_and_reduces_the_error(ffx):
    """`NanoOPQ(device=0).fit`: rotation products on the GPU (ffx_sgemm), SVD on the host.  R stays
    orthogonal, and the rotated quantizer reconstructs correlated data better than plain PQ —
    what OPQ is for — as well as the host-trained one."""
    from fast_forward.quantizer import NanoOPQ, NanoPQ

    rng = np.random.default_rng(9)
    mix = rng.standard_normal((64, 64)).astype(np.float32)
    x = (rng.standard_normal((6000, 64)).astype(np.float32) * np.linspace(3, 0.2, 64, dtype=np.float32)) @ mix
    y = (rng.standard_normal((1500, 64)).astype(np.float32) * np.linspace(3, 0.2, 64, dtype=np.float32)) @ mix
    dev, host, plain = NanoOPQ(8, 32, device=0), NanoOPQ(8, 32), NanoPQ(8, 32)
    dev.fit(x, pq_iter=5, rotation_iter=4)
    host.fit(x, pq_iter=5, rotation_iter=4)
    plain.fit(x, iter=5)
    R = dev.adc_tables()[1]
    assert np.allclose(R @ R.T, np.eye(64), atol=1e-4)
    err = {name: float(((q.decode(q.encode(y)) - y) ** 2).mean()) for name, q in (("dev", dev), ("host", host), ("plain", plain))}
    assert err["dev"] < err["plain"] and err["dev"] <= err["host"] * 1.05, err
