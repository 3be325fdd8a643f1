"""ffx_rerank_early_stop (index/base.py:316-387 as one launch) through the C ABI against the
oracle's restatement, on cases larger and more ragged than the golden frames: hundreds of
queries of uneven length, scattered document rows, depth lists with repeats and depths
below the cutoff, queries shorter than the first depth."""

import ctypes

import numpy as np
import pytest

import ff_oracle as fo
from test_gpu_parity import MODES, bits, c_scores, make_corpus, make_pairs, units_for_mode

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ffx():
    import __graft_entry__ as g

    g.build()
    from fast_forward import _ffx

    assert _ffx.device_count() >= 1
    return _ffx


def falling_lex(rng, q_off, scale):
    """first-stage scores in rank order, with a different steepness per query"""
    parts = []
    for q in range(len(q_off) - 1):
        n = int(q_off[q + 1] - q_off[q])
        steep = rng.uniform(0.6, 0.999)
        parts.append(6.0 * scale * steep ** np.arange(n))
    return np.concatenate(parts + [np.zeros(0)]).astype(np.float32)


@pytest.mark.parametrize("dim,contiguous", [(768, True), (768, False), (384, True), (2048, True)])
@pytest.mark.parametrize("mode", ["MAXP", "AVEP", "FIRSTP", "PASSAGE"])
def test_early_stop_matches_oracle(ffx, oracle_c, dim, contiguous, mode):
    rng = np.random.default_rng(dim + 7 * contiguous)
    m = MODES[mode]
    n_docs = 900
    off, rows, vec = make_corpus(rng, n_docs, 6, dim, contiguous)
    idx = ffx.DeviceIndex(dim, capacity=len(vec))
    idx.stage(0, vec)
    idx.set_docs(off, None if contiguous else rows)
    nq = 150
    pool = len(vec) if m == fo.MODE_PASSAGE else n_docs
    qv = rng.standard_normal((nq, dim)).astype(np.float32)
    q_off, cand, pair_q = make_pairs(rng, nq, pool, 1, 400)
    lex = falling_lex(rng, q_off, float(np.sqrt(dim)))
    u_off, u_rows = units_for_mode(off, rows, len(vec), m)
    ff = c_scores(oracle_c, vec, u_off, u_rows, pair_q, cand, qv, m)
    sizes = np.diff(q_off)
    seen = set()
    for cutoff, alpha, depths in ((10, 0.5, (10, 50, 100, 200, 400)), (5, 0.2, (3, 7, 7, 300)),
                                  (20, 0.8, (400, 25, 100)), (1, 0.5, tuple(range(1, 33))),
                                  (50, 0.3, (20, 30))):
        want = fo.early_stopping_depth(q_off, lex, ff, alpha, cutoff, depths)
        out = idx.rerank_early_stop_host(m, qv, q_off, cand, lex, alpha, cutoff, depths, want_int=True)
        assert (out["scored"] == want).all(), (cutoff, alpha, depths)
        scored = (np.arange(len(cand)) - np.repeat(q_off[:-1], sizes)) < np.repeat(want, sizes)
        assert (bits(out["ff"][scored]) == bits(ff[scored])).all()
        assert (out["ff"][~scored] == 0).all()
        assert (bits(out["int"][scored]) == bits(fo.interpolate_f32(lex, ff, alpha)[scored])).all()
        seen.update(np.unique(want / sizes).round(3).tolist())
    assert len(seen) > 3  # the cases really stop at different depths
    idx.close()


def test_early_stop_limits(ffx):
    rng = np.random.default_rng(0)
    q_off = np.array([0, 10], np.int64)
    idx = ffx.DeviceIndex(768, capacity=64)
    idx.stage(0, rng.standard_normal((64, 768)).astype(np.float32))
    qv = rng.standard_normal((1, 768)).astype(np.float32)
    args = (fo.MODE_PASSAGE, qv, q_off, np.arange(10, dtype=np.int32), np.ones(10, np.float32), 0.5)
    with pytest.raises(ffx.FFXError):
        idx.rerank_early_stop_host(*args, 1, tuple(range(1, 40)))  # > 32 distinct depths
    with pytest.raises(ffx.FFXError):
        idx.rerank_early_stop_host(*args, 0, (2, 4))  # cutoff < 1
    out = idx.rerank_early_stop_host(*args, 3, (1, 2))  # every depth below the cutoff: nothing scored
    assert (out["scored"] == 0).all() and (out["ff"] == 0).all()
    bad = np.arange(10, dtype=np.int32)
    bad[4] = 64
    with pytest.raises(ffx.FFXError):  # out-of-range candidate is reported, never dereferenced
        idx.rerank_early_stop_host(fo.MODE_PASSAGE, qv, q_off, bad, np.ones(10, np.float32), 0.5, 2, (10,))
    idx.close()


@pytest.mark.parametrize("kind", ["dim100", "dim128", "dim3072", "pq96", "pq8"])
@pytest.mark.parametrize("mode", ["MAXP", "AVEP", "PASSAGE"])
def test_device_walk_for_the_other_index_kinds(ffx, oracle_c, kind, mode):
    """Indexes the one-launch kernel does not cover (dimensions without a lane-major plan, short
    and long rows, PQ / OPQ codes) walk the depths as a stream-ordered sequence of launches on
    the device.  Checked against the oracle's restatement applied to the scores the index itself
    produces without early stopping (so that the PQ case does not depend on ADC rounding)."""
    rng = np.random.default_rng(len(kind) * 7 + len(mode))
    m = MODES[mode]
    n_docs = 500
    off, rows, _ = make_corpus(rng, n_docs, 5, 4, True)
    n_rows = int(off[-1])
    if kind.startswith("pq"):
        M, Ks, Ds = int(kind[2:]), 32, 4
        D = M * Ds
        idx = ffx.DeviceIndex(M, capacity=n_rows, row_kind=ffx.ROWS_PQ_U8)
        idx.stage(0, rng.integers(0, Ks, (n_rows, M)).astype(np.uint8))
        idx.set_docs(off)
        idx.set_pq(rng.standard_normal((M, Ks, Ds)).astype(np.float32),
                   np.linalg.qr(rng.standard_normal((D, D)))[0].astype(np.float32))
    else:
        D = int(kind[3:])
        idx = ffx.DeviceIndex(D, capacity=n_rows)
        idx.stage(0, rng.standard_normal((n_rows, D)).astype(np.float32))
        idx.set_docs(off)
    nq = 90
    pool = n_rows if m == fo.MODE_PASSAGE else n_docs
    qv = rng.standard_normal((nq, D)).astype(np.float32)
    q_off, cand, pair_q = make_pairs(rng, nq, pool, 1, 300)
    lex = falling_lex(rng, q_off, float(np.sqrt(D)))
    full = idx.rerank_host(m, qv, q_off, cand, want_ff=True)["ff"]  # no early stopping
    sizes = np.diff(q_off)
    stops = set()
    for cutoff, alpha, depths in ((5, 0.5, (5, 20, 60, 150, 300)), (3, 0.8, (10, 10, 200)), (10, 0.2, (4, 50, 120))):
        out = idx.rerank_early_stop_host(m, qv, q_off, cand, lex, alpha, cutoff, depths, want_int=True)
        scored = (np.arange(len(cand)) - np.repeat(q_off[:-1], sizes)) < np.repeat(out["scored"], sizes)
        if kind == "pq96":
            # the XOR-swizzled ADC kernel sums a row's table entries in an order that depends on the
            # lane that scores it, and the compact lists of the walk move pairs to other lanes: scores
            # agree to reassociation, and the walk must be the oracle's walk over ITS OWN scores
            assert np.allclose(out["ff"][scored], full[scored], rtol=1e-4, atol=1e-3)
            seen_scores = np.where(scored, out["ff"], full)
        else:
            assert (bits(out["ff"][scored]) == bits(full[scored])).all()
            seen_scores = full
        want = fo.early_stopping_depth(q_off, lex, seen_scores, alpha, cutoff, depths)
        assert (out["scored"] == want).all(), (cutoff, alpha, depths)
        assert (out["ff"][~scored] == 0).all()
        assert (bits(out["int"][scored]) == bits(fo.interpolate_f32(lex, seen_scores, alpha)[scored])).all()
        stops.update((want / sizes).round(2).tolist())
    assert len(stops) > 2
    idx.close()
