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

DIMS = [768, 384, 1024, 512, 3072, 4096, 256, 128, 192, 64, 100, 33, 8]


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
    dim = int(rng.choice([768, 384, 1024, 640, 2560, 256, 100]))
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


@pytest.mark.parametrize("seed", range(12))
def test_random_shards_equal_the_whole(ffx, seed):
    """Doc-id-range shards on random shapes (fp32 and PQ indexes, 2..6 shards, fused and tiled
    launches): per-pair scores written once across shards and the merged per-shard top-k lists
    equal the unsharded index's — bit for bit on fp32 indexes."""
    import torch

    from fast_forward.sharded import plan_doc_shards

    rng = np.random.default_rng(7000 + seed)
    pq = bool(seed % 3 == 2)
    n_docs = int(rng.integers(10, 600))
    cnt = rng.integers(1, int(rng.choice([2, 9, 40])) + 1, n_docs)
    off = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
    n_rows = int(off[-1])
    if pq:
        M, Ks, Ds = int(rng.choice([96, 64, 8])), 64, 4
        dim, D = M, M * Ds
        rows_data = rng.integers(0, Ks, (n_rows, M)).astype(np.uint8)
        cw = rng.standard_normal((M, Ks, Ds)).astype(np.float32)
        R = np.linalg.qr(rng.standard_normal((D, D)))[0].astype(np.float32)
        kind = ffx.ROWS_PQ_U8
    else:
        dim = D = int(rng.choice([768, 384, 100]))
        rows_data = rng.standard_normal((n_rows, dim)).astype(np.float32)
        kind = ffx.ROWS_F32

    def build(r0, r1, d0, d1):
        ix = ffx.DeviceIndex(dim, capacity=max(int(r1 - r0), 1), row_kind=kind)
        ix.stage(0, rows_data[r0:r1])
        ix.set_docs(off[d0:d1 + 1] - r0)
        if pq:
            ix.set_pq(cw, R)
        return ix

    whole = build(0, n_rows, 0, n_docs)
    S = int(rng.integers(2, 7))
    bounds = plan_doc_shards(cnt, S)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
    shards = []
    for s in range(S):
        lo, hi = int(bounds[s]), int(bounds[s + 1])
        if hi == lo:
            shards.append(None)
            continue
        sh = build(int(off[lo]), int(off[hi]), lo, hi)
        sh.set_shard(lo, n_docs, int(off[lo]), n_rows)
        shards.append(sh)
    try:
        for _ in range(3):
            mode = int(rng.choice([fo.MODE_PASSAGE, fo.MODE_MAXP, fo.MODE_FIRSTP, fo.MODE_AVEP]))
            pool = n_rows if mode == fo.MODE_PASSAGE else n_docs
            nq = int(rng.choice([3, 160, 320]))
            cnts = rng.integers(0, int(rng.choice([5, 120])) + 1, nq)
            if cnts.max() == 0:
                cnts[0] = 1
            q_off = np.concatenate([[0], np.cumsum(cnts)]).astype(np.int64)
            cand = rng.integers(0, pool, int(q_off[-1])).astype(np.int32)
            lex = (rng.integers(0, 30, len(cand)) / 2).astype(np.float32)
            qv = rng.standard_normal((nq, D)).astype(np.float32)
            k, alpha = int(rng.choice([1, 7, 50])), float(rng.choice([0.1, 1.0]))
            want = whole.rerank_host(mode, qv, q_off, cand, lex, alpha, k, want_ff=True)
            d_qv, d_off, d_cand, d_lex = dev(qv), dev(q_off), dev(cand), dev(lex)
            sh_s = torch.full((S, nq, k), float("-inf"), dtype=torch.float32, device="cuda")
            sh_p = torch.full((S, nq, k), -1, dtype=torch.int32, device="cuda")
            ff = torch.zeros(len(cand), dtype=torch.float32, device="cuda")
            torch.cuda.synchronize()
            for s, sh in enumerate(shards):
                if sh is None:
                    continue
                sh.rerank_device(mode, d_qv.data_ptr(), nq, d_off.data_ptr(), d_cand.data_ptr(), d_lex.data_ptr(),
                                 alpha, k, int(cnts.max()), ff.data_ptr(), 0, sh_s[s].data_ptr(), sh_p[s].data_ptr())
                sh.sync()
            o_s = torch.empty((nq, k), dtype=torch.float32, device="cuda")
            o_p = torch.empty((nq, k), dtype=torch.int32, device="cuda")
            ffx.merge_topk(0, sh_s.data_ptr(), sh_p.data_ptr(), S, nq, k, o_s.data_ptr(), o_p.data_ptr())
            torch.cuda.synchronize()
            tag = (seed, pq, dim, S, mode, nq, k, alpha)
            got_ff = ff.cpu().numpy()
            if pq:
                # the ADC sum order of a row depends on the lane that scores it, and a shard skips
                # foreign pairs: scores agree to reassociation, the merged lists must be exactly
                # the ranking of the scores the shards produced
                assert np.allclose(got_ff, want["ff"], rtol=1e-4, atol=1e-3), tag
                ts, tp = fo.topk_per_query(q_off, fo.interpolate_f32(lex, got_ff, alpha), k)
            else:
                assert (bits(got_ff) == bits(want["ff"])).all(), tag
                ts, tp = want["topk_score"], want["topk_pos"]
            assert (o_p.cpu().numpy() == tp).all(), tag
            assert (bits(o_s.cpu().numpy()) == bits(ts)).all(), tag
    finally:
        whole.close()
        for sh in shards:
            if sh is not None:
                sh.close()
