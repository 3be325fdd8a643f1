"""Randomised differential test of ffx_rerank_host against the plain-C oracle: seeded random
shapes (lane-major and generic dimensions, contiguous and scattered documents, all modes, empty
and ragged candidate lists, k from 0 to beyond the list length, alpha in {0, 1, random}, with
and without lexical scores, with and without per-pair outputs), bit-exact every time.  Meant to
catch the rare-shape bugs that fixed-shape tests miss (e.g. a partial tail chunk)."""

import numpy as np
import pytest

import ff_oracle as fo
from test_gpu_parity import bits, c_scores, make_corpus, units_for_mode

pytestmark = pytest.mark.gpu

DIMS = [768, 384, 1024, 512, 100, 33, 8]


@pytest.fixture(scope="module")
def ffx():
    import __graft_entry__ as g

    g.build()
    from fast_forward import _ffx

    assert _ffx.device_count() >= 1
    return _ffx


@pytest.mark.parametrize("seed", range(60))
def test_random_shapes_bit_exact(ffx, oracle_c, seed):
    rng = np.random.default_rng(1000 + seed)
    dim = int(rng.choice(DIMS))
    n_docs = int(rng.integers(1, 400))
    contiguous = bool(rng.integers(0, 2))
    off, rows, vec = make_corpus(rng, n_docs, int(rng.integers(1, 40)), dim, contiguous)
    idx = ffx.DeviceIndex(dim, capacity=len(vec) + int(rng.integers(0, 50)))
    idx.stage(0, vec)
    idx.set_docs(off, None if contiguous else rows)
    ffx.set_option("kernel", int(rng.integers(0, 3)))
    try:
        for _ in range(4):
            mode = int(rng.choice([fo.MODE_PASSAGE, fo.MODE_MAXP, fo.MODE_FIRSTP, fo.MODE_AVEP]))
            pool = len(vec) if mode == fo.MODE_PASSAGE else n_docs
            nq = int(rng.choice([1, 2, 7, 150, 310, 700]))
            hi = int(rng.choice([0, 1, 5, 40, 260]))
            cnts = rng.integers(0, hi + 1, nq)
            q_off = np.concatenate([[0], np.cumsum(cnts)]).astype(np.int64)
            cand = rng.integers(0, pool, int(q_off[-1])).astype(np.int32)  # repeats inside a list are allowed
            pair_q = np.repeat(np.arange(nq), cnts)
            qv = rng.standard_normal((nq, dim)).astype(np.float32)
            lex = None if rng.random() < 0.25 else (rng.integers(0, 9, len(cand)) * rng.choice([0.5, 1.7])).astype(np.float32)
            alpha = float(rng.choice([0.0, 1.0, rng.random()]))
            k = int(rng.choice([0, 1, 3, 64, max(1, hi), hi + 9]))
            want_ff, want_int = bool(rng.integers(0, 2)), bool(rng.integers(0, 2))
            if k == 0:
                want_ff = True
            out = idx.rerank_host(mode, qv, q_off, cand, lex, alpha, k, want_ff=want_ff, want_int=want_int)
            u_off, u_rows = units_for_mode(off, rows, len(vec), mode)
            ff = c_scores(oracle_c, vec, u_off, u_rows, pair_q, cand, qv, mode) if len(cand) else np.zeros(0, np.float32)
            it = ff if lex is None else fo.interpolate_f32(lex, ff, alpha)
            tag = (seed, dim, mode, nq, hi, k, alpha, lex is None, contiguous)
            if want_ff:
                assert (bits(out["ff"]) == bits(ff)).all(), tag
            if want_int:
                assert (bits(out["int"]) == bits(it)).all(), tag
            if k > 0:
                ts, tp = fo.topk_per_query(q_off, it, k)
                assert (out["topk_pos"] == tp).all(), tag
                assert (bits(out["topk_score"]) == bits(ts)).all(), tag
    finally:
        ffx.set_option("kernel", 0)
        idx.close()


@pytest.mark.parametrize("seed", range(16))
def test_random_early_stopping_against_the_oracle(ffx, oracle_c, seed):
    """ffx_rerank_early_stop on random shapes: rows scored per query and their scores equal the
    restatement of index/base.py:316-387 (itself pinned on the reference's frames)."""
    rng = np.random.default_rng(5000 + seed)
    dim = int(rng.choice([768, 384, 1024, 640]))
    n_docs = int(rng.integers(30, 500))
    contiguous = bool(rng.integers(0, 2))
    off, rows, vec = make_corpus(rng, n_docs, int(rng.integers(1, 12)), dim, contiguous)
    idx = ffx.DeviceIndex(dim, capacity=len(vec))
    idx.stage(0, vec)
    idx.set_docs(off, None if contiguous else rows)
    for _ in range(3):
        mode = int(rng.choice([fo.MODE_PASSAGE, fo.MODE_MAXP, fo.MODE_FIRSTP, fo.MODE_AVEP]))
        pool = len(vec) if mode == fo.MODE_PASSAGE else n_docs
        nq = int(rng.choice([1, 5, 60, 300]))
        cnts = rng.integers(1, int(rng.choice([3, 30, 200])) + 1, nq)
        q_off = np.concatenate([[0], np.cumsum(cnts)]).astype(np.int64)
        cand = rng.integers(0, pool, int(q_off[-1])).astype(np.int32)
        pair_q = np.repeat(np.arange(nq), cnts)
        qv = rng.standard_normal((nq, dim)).astype(np.float32)
        scale = float(np.sqrt(dim)) * float(rng.choice([0.3, 3.0, 10.0]))
        lex = np.concatenate([np.sort(rng.random(c))[::-1] * scale * rng.uniform(0.2, 5) for c in cnts]).astype(np.float32)
        alpha = float(rng.choice([0.05, 0.5, 0.9, 1.0, 0.0]))
        cutoff = int(rng.choice([1, 2, 5, 20]))
        depths = tuple(int(d) for d in rng.choice([1, 2, 3, 5, 8, 13, 40, 100, 250], size=int(rng.integers(1, 6))))
        u_off, u_rows = units_for_mode(off, rows, len(vec), mode)
        ff = c_scores(oracle_c, vec, u_off, u_rows, pair_q, cand, qv, mode)
        want = fo.early_stopping_depth(q_off, lex, ff, alpha, cutoff, depths)
        out = idx.rerank_early_stop_host(mode, qv, q_off, cand, lex, alpha, cutoff, depths)
        tag = (seed, dim, mode, nq, alpha, cutoff, depths)
        assert (out["scored"] == want).all(), tag
        scored = (np.arange(len(cand)) - np.repeat(q_off[:-1], cnts)) < np.repeat(want, cnts)
        assert (bits(out["ff"][scored]) == bits(ff[scored])).all(), tag
    idx.close()


@pytest.mark.parametrize("seed", range(16))
def test_random_adc_shapes_agree_across_kernels(ffx, seed):
    """PQ scoring on random shapes: the three ADC kernels agree with each other (reassociation
    tolerance) and rank their own scores exactly like the oracle's ordering rule."""
    rng = np.random.default_rng(9000 + seed)
    M = int(rng.choice([32, 64, 96, 128, 8, 20, 72]))
    Ks = int(rng.choice([256, 16, 100]))
    Ds = int(rng.choice([2, 4, 8]))
    n_docs = int(rng.integers(5, 300))
    contiguous = bool(rng.integers(0, 2))
    off, rows, _ = make_corpus(rng, n_docs, int(rng.integers(1, 50)), 4, contiguous)
    n_rows = int(off[-1])
    idx = ffx.DeviceIndex(M, capacity=n_rows, row_kind=ffx.ROWS_PQ_U8)
    idx.stage(0, rng.integers(0, Ks, (n_rows, M)).astype(np.uint8))
    idx.set_docs(off, None if contiguous else rows)
    D = M * Ds
    R = np.linalg.qr(rng.standard_normal((D, D)))[0].astype(np.float32) if rng.integers(0, 2) else None
    idx.set_pq(rng.standard_normal((M, Ks, Ds)).astype(np.float32), R)
    try:
        for _ in range(3):
            mode = int(rng.choice([fo.MODE_PASSAGE, fo.MODE_MAXP, fo.MODE_FIRSTP, fo.MODE_AVEP]))
            pool = n_rows if mode == fo.MODE_PASSAGE else n_docs
            nq = int(rng.choice([1, 9, 149, 200]))
            cnts = rng.integers(0, int(rng.choice([1, 40, 300])) + 1, nq)
            q_off = np.concatenate([[0], np.cumsum(cnts)]).astype(np.int64)
            cand = rng.integers(0, pool, int(q_off[-1])).astype(np.int32)
            qv = rng.standard_normal((nq, D)).astype(np.float32)
            lex = rng.uniform(0, 20, len(cand)).astype(np.float32)
            k = int(rng.choice([1, 10, 64, 301]))
            ref = None
            for adc in (1, 2, 3):
                ffx.set_option("adc", adc)
                out = idx.rerank_host(mode, qv, q_off, cand, lex, 0.3, k, want_ff=True, want_int=True)
                assert (bits(out["int"]) == bits(fo.interpolate_f32(lex, out["ff"], 0.3))).all()
                ts, tp = fo.topk_per_query(q_off, out["int"], k)
                assert (out["topk_pos"] == tp).all() and (bits(out["topk_score"]) == bits(ts)).all(), (seed, M, Ks, Ds, mode, nq, k, adc)
                ref = out["ff"] if ref is None else ref
                assert np.allclose(out["ff"], ref, rtol=1e-4, atol=1e-4 * np.sqrt(D) * 4), (seed, M, adc)
    finally:
        ffx.set_option("adc", 0)
        idx.close()
