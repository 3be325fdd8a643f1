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
