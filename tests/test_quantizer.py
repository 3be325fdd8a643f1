"""Quantizer API, modelled on the reference's tests/test_quantizer.py:12-79 (properties,
shapes/dtypes, serialise -> deserialise equality, guards), plus the algebra the ADC kernel
relies on.  Decode VALUES are "parity unpinned" (SURVEY 8c): nanopq is not available, so the
product implementation is checked against the oracle's independent restatement instead."""

import numpy as np
import pytest

import ff_oracle as fo
from fast_forward.quantizer import NanoOPQ, NanoPQ, Quantizer


@pytest.fixture(scope="module", params=[NanoPQ, NanoOPQ])
def pair(request):
    rng = np.random.default_rng(0)
    fresh = request.param(8, 256)
    trained = request.param(8, 256)
    kw = {"rotation_iter": 2, "pq_iter": 3} if request.param is NanoOPQ else {"iter": 3}
    trained.fit(rng.normal(size=(2**9, 768)).astype(np.float32), **kw)
    return fresh, trained


def test_properties(pair):
    fresh, trained = pair
    assert fresh.dims == (None, 8) and fresh.dtype == np.uint8 and not fresh._trained
    assert trained.dims == (768, 8) and trained.dtype == np.uint8 and trained._trained
    assert fresh == fresh and trained == trained and fresh != trained and fresh != object()
    assert fresh.adc_tables() is None
    cw, R = trained.adc_tables()
    assert cw.shape == (8, 256, 96) and (R is None) == isinstance(trained, NanoPQ)


def test_encode_decode_shapes_and_values(pair):
    _, q = pair
    x = np.random.default_rng(1).normal(size=(8, 768)).astype(np.float32)
    codes = q.encode(x)
    assert codes.shape == (8, 8) and codes.dtype == np.uint8
    dec = q.decode(codes)
    assert dec.shape == x.shape and dec.dtype == np.float32
    cw, R = q.adc_tables()
    want = fo.pq_decode(codes, cw) if R is None else fo.opq_decode(codes, cw, R)
    assert np.array_equal(dec, want)
    # nearest codeword per subspace (of the rotated vector for OPQ)
    xr = x if R is None else x @ R
    for m in range(8):
        d = ((xr[:, None, m * 96:(m + 1) * 96] - cw[m][None]) ** 2).sum(-1)
        assert (d.argmin(1) == codes[:, m]).all()
    if R is not None:
        assert np.allclose(R @ R.T, np.eye(768), atol=1e-4)


def test_serialization_roundtrip(pair):
    fresh, trained = pair
    assert Quantizer.deserialize(*fresh.serialize()) == fresh
    loaded = Quantizer.deserialize(*trained.serialize())
    assert loaded == trained and type(loaded) is type(trained)
    meta, attrs, data = trained.serialize()
    assert meta["__module__"] == "fast_forward.quantizer.nanopq" and meta["_trained"] is True
    assert set(attrs) == {"M", "Ks", "Ds", "metric", "verbose"} and "codewords" in data
    x = np.random.default_rng(2).normal(size=(8, 768)).astype(np.float32)
    assert np.array_equal(trained.encode(x), loaded.encode(x))


def test_guards(pair):
    fresh, trained = pair
    x = np.zeros((8, 768), np.float32)
    with pytest.raises(RuntimeError):
        fresh.encode(x)
    with pytest.raises(RuntimeError):
        fresh.decode(np.zeros((8, 8), np.uint8))
    with pytest.raises(RuntimeError):
        fresh.set_attached()
    q = Quantizer.deserialize(*trained.serialize())
    q.set_attached()
    with pytest.raises(RuntimeError):
        q.fit(x)


def test_encode_is_scipy_vq_and_decode_is_the_codeword_gather(pair):
    """The one pin available without nanopq (absent from the image and the wheelhouse, like h5py
    and libhdf5): nanopq 0.2.1's `PQ.encode` IS `scipy.cluster.vq.vq` per subspace and its
    `PQ.decode` IS the gather `codewords[m][codes[:, m], :]` (the calls the reference makes at
    quantizer/nanopq.py:41-44,109-112).  scipy is installed: same data through both must give
    identical codes, and decode must be the pure gather (then `@ R.T` for OPQ)."""
    from scipy.cluster.vq import vq

    _, q = pair
    rng = np.random.default_rng(3)
    x = rng.normal(size=(300, 768)).astype(np.float32)
    cw, R = q.adc_tables()
    M, Ks, Ds = cw.shape
    xr = x if R is None else x @ R
    want = np.empty((len(x), M), np.uint8)
    for m in range(M):
        want[:, m], _ = vq(xr[:, m * Ds:(m + 1) * Ds], cw[m])
    codes = q.encode(x)
    assert codes.dtype == np.uint8 and np.array_equal(codes, want)
    gathered = np.concatenate([cw[m][codes[:, m], :] for m in range(M)], axis=1)
    assert np.array_equal(q.decode(codes), gathered if R is None else gathered @ R.T)
    # random codes, not only the ones encode produces
    rnd = rng.integers(0, Ks, (200, M)).astype(np.uint8)
    gathered = np.concatenate([cw[m][rnd[:, m], :] for m in range(M)], axis=1)
    assert np.array_equal(q.decode(rnd), gathered if R is None else gathered @ R.T)
