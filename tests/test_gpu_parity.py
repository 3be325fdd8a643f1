"""GPU parity tests proper: the CUDA path, called through the C ABI (libffx.so), against the
oracle and against the golden vectors written by the unmodified reference.

Bar: bit-exact for ids / candidate sets / positions and for fp32 scores (N1/N2/N3 make the
fp32 path reproducible bit-for-bit); the PQ/OPQ ADC path reorders arithmetic and is held to
rtol 1e-5 + atol 1e-5*|q||d| (stated in the test)."""

import ctypes

import numpy as np
import pytest

import ff_oracle as fo

pytestmark = pytest.mark.gpu

MODES = {"PASSAGE": fo.MODE_PASSAGE, "MAXP": fo.MODE_MAXP, "FIRSTP": fo.MODE_FIRSTP,
         "AVEP": fo.MODE_AVEP}


def bits(x):
    return np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)


@pytest.fixture(scope="module")
def ffx():
    import __graft_entry__ as g

    g.build()
    from fast_forward import _ffx

    assert _ffx.device_count() >= 1, "gpu tests need a CUDA device"
    return _ffx


@pytest.fixture(params=["ldg", "tma"], autouse=True)
def kernel_variant(request, ffx):
    """Every test of this file runs against both scoring kernels: the register-staged
    ffx_score_kernel and the TMA-staged ffx_score_tma_kernel (the default)."""
    ffx.set_option("kernel", 1 if request.param == "ldg" else 2)
    yield request.param
    ffx.set_option("kernel", 0)


def c_scores(oracle_c, vec, u_off, u_rows, pair_q, pair_u, qv, mode):
    vec = np.ascontiguousarray(vec, np.float32)
    qv = np.ascontiguousarray(qv, np.float32)
    u_off = np.ascontiguousarray(u_off, np.int64)
    u_rows = np.ascontiguousarray(u_rows, np.int64)
    pair_q = np.ascontiguousarray(pair_q, np.int64)
    pair_u = np.ascontiguousarray(pair_u, np.int64)
    out = np.empty(len(pair_q), np.float32)
    P = ctypes.c_void_p
    oracle_c.ffo_score_pairs(P(vec.ctypes.data), ctypes.c_int64(vec.shape[1]), P(u_off.ctypes.data),
                             P(u_rows.ctypes.data), P(pair_q.ctypes.data), P(pair_u.ctypes.data),
                             ctypes.c_int64(len(pair_q)), P(qv.ctypes.data), ctypes.c_int(mode),
                             P(out.ctypes.data))
    return out


def units_for_mode(doc_off, doc_rows, n_rows, mode):
    """Mode-resolved CSR for the oracle (index/util.py:29-41)."""
    if mode == fo.MODE_PASSAGE:
        return np.arange(n_rows + 1), np.arange(n_rows)
    if mode == fo.MODE_FIRSTP:
        return np.arange(len(doc_off)), doc_rows[doc_off[:-1]]
    return doc_off, doc_rows


def make_corpus(rng, n_docs, max_psg, dim, contiguous=True):
    cnt = rng.integers(1, max_psg + 1, n_docs)
    off = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
    n_rows = int(off[-1])
    rows = np.arange(n_rows, dtype=np.int64)
    if not contiguous:  # documents own scattered rows, listed in increasing (insertion) order
        owner = np.repeat(np.arange(n_docs), cnt)
        rng.shuffle(owner)
        order = np.argsort(owner, kind="stable")
        rows = order.astype(np.int64)
    vec = rng.standard_normal((n_rows, dim)).astype(np.float32)
    return off, rows, vec


def make_pairs(rng, nq, pool, lo, hi):
    cnts = rng.integers(lo, hi + 1, nq)
    q_off = np.concatenate([[0], np.cumsum(cnts)]).astype(np.int64)
    cand = np.concatenate([rng.choice(pool, c, replace=False) for c in cnts] + [np.zeros(0, np.int64)])
    return q_off, cand.astype(np.int32), np.repeat(np.arange(nq), cnts)


# ------------------------------------------------------------------------------------------
# golden vectors of the unmodified reference, through the C ABI
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("key", ["s11", "s12", "s13", "s14"])
@pytest.mark.parametrize("mode", ["MAXP", "AVEP", "FIRSTP", "PASSAGE"])
def test_golden_reference_vectors(ffx, golden_random, key, mode):
    meta, arrays = golden_random
    case = meta[key]
    g = case["modes"][mode]
    vec, qvecs = arrays[f"{key}/vectors"], arrays[f"{key}/qvecs"]
    doc_names = list(dict.fromkeys(case["doc_ids"]))
    doc_ord = {d: i for i, d in enumerate(doc_names)}
    lists = [[] for _ in doc_names]
    for row, d in enumerate(case["doc_ids"]):
        lists[doc_ord[d]].append(row)
    doc_off = np.concatenate([[0], np.cumsum([len(x) for x in lists])])
    doc_rows = np.concatenate(lists)
    psg_row = {p: i for i, p in enumerate(case["psg_ids"])}

    idx = ffx.DeviceIndex(case["dim"], capacity=len(vec))
    half = len(vec) // 2
    idx.stage(0, vec[:half])
    idx.stage(half, vec[half:])
    idx.set_docs(doc_off, doc_rows)
    assert idx.has_fast_path  # lane-major plan (384, 768, 1024) or the tree-as-data kernel (100)

    # integer-code the first-stage ranking: queries in order of appearance, candidates of a
    # query in ascending id order (the order the reference's outer merge leaves ties in)
    fs = g["first_stage"]
    q_names = list(dict.fromkeys(fs["q_id"]))
    per_q = {q: [] for q in q_names}
    for q, i, s in zip(fs["q_id"], fs["id"], fs["score"]):
        per_q[q].append((i, s))
    cand, lex, q_off, names = [], [], [0], []
    for q in q_names:
        for i, s in sorted(per_q[q]):
            cand.append(psg_row[i] if mode == "PASSAGE" else doc_ord[i])
            lex.append(s)
            names.append((q, i))
        q_off.append(len(cand))
    qv = np.stack([qvecs[int(q[1:])] for q in q_names])
    k = case["cutoff"]
    out = idx.rerank_host(MODES[mode], qv, q_off, cand, lex, case["alpha"], k, want_ff=True, want_int=True)

    want_ff = {(q, i): b for q, i, b in zip(g["ff"]["q_id"], g["ff"]["id"], g["ff"]["score_bits"])}
    got_ff = dict(zip(names, bits(out["ff"]).tolist()))
    assert got_ff == want_ff
    want_int = {(q, i): b for q, i, b in zip(g["interpolated"]["q_id"], g["interpolated"]["id"],
                                             g["interpolated"]["score_bits"])}
    assert dict(zip(names, bits(out["int"]).tolist())) == want_int
    # cut: same rows, same order, same bits — per query, in the reference's frame order
    got = []
    for qi, q in enumerate(q_names):
        for j in range(k):
            p = out["topk_pos"][qi, j]
            if p >= 0:
                got.append((q, names[q_off[qi] + p][1], int(bits(out["topk_score"][qi, j:j + 1])[0])))
    want = list(zip(g["cut"]["q_id"], g["cut"]["id"], g["cut"]["score_bits"]))
    order = {q: n for n, q in enumerate(dict.fromkeys(g["cut"]["q_id"]))}
    got.sort(key=lambda t: order[t[0]])  # stable: keeps our within-query order
    assert got == want
    idx.close()


def test_reference_known_answers(ffx, golden_kat):
    """reference tests/test_index.py:135-200 fixtures (5x5 lower-triangular, d0 has 2 rows)."""
    V = np.tril(np.ones((5, 5), dtype=np.float32))
    idx = ffx.DeviceIndex(5, capacity=5)
    idx.stage(0, V)
    idx.set_docs([0, 2, 3, 4, 5])
    qv = np.ones((2, 5), np.float32)
    q_off, cand = [0, 4, 8], [0, 1, 2, 3] * 2
    for mode, d0 in (("MAXP", 2.0), ("FIRSTP", 1.0), ("AVEP", 1.5)):
        ff = idx.rerank_host(MODES[mode], qv, q_off, cand)["ff"]
        assert ff.tolist() == [d0, 3.0, 4.0, 5.0] * 2
        want = golden_kat[f"full/{mode}"]
        by = {(q, i): s for q, i, s in zip(want["q_id"], want["id"], want["score"])}
        assert [by[("q1", f"d{c}")] for c in range(4)] == ff[:4].tolist()
    ff = idx.rerank_host(fo.MODE_PASSAGE, qv, [0, 5, 10], list(range(5)) * 2)["ff"]
    assert ff.tolist() == [1.0, 2.0, 3.0, 4.0, 5.0] * 2
    idx.close()


# ------------------------------------------------------------------------------------------
# seeded random problems against the oracle
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dim", [768, 384, 512, 640, 896, 1024, 1536, 2048, 2560, 3072, 3584, 4096, 64, 96, 128, 192,
                                 256, 100, 5, 130, 1, 7, 8, 9, 50, 200, 260, 300, 312, 1000, 1280, 2000, 4000, 4100])
@pytest.mark.parametrize("contiguous", [True, False])
def test_scores_bit_exact_all_modes(ffx, oracle_c, dim, contiguous):
    rng = np.random.default_rng(dim * 2 + contiguous)
    n_docs = 400
    off, rows, vec = make_corpus(rng, n_docs, 9, dim, contiguous)
    idx = ffx.DeviceIndex(dim, capacity=len(vec))
    idx.stage(0, vec)
    idx.set_docs(off, rows)
    nq = 7
    qv = rng.standard_normal((nq, dim)).astype(np.float32)
    for name, mode in MODES.items():
        pool = len(vec) if mode == fo.MODE_PASSAGE else n_docs
        q_off, cand, pair_q = make_pairs(rng, nq, pool, 0, 150)
        lex = (rng.integers(0, 40, len(cand)) / 2).astype(np.float32)
        out = idx.rerank_host(mode, qv, q_off, cand, lex, 0.1, 12, want_ff=True, want_int=True)
        u_off, u_rows = units_for_mode(off, rows, len(vec), mode)
        ff = c_scores(oracle_c, vec, u_off, u_rows, pair_q, cand, qv, mode)
        assert (bits(out["ff"]) == bits(ff)).all(), name
        it = fo.interpolate_f32(lex, ff, 0.1)
        assert (bits(out["int"]) == bits(it)).all(), name
        ts, tp = fo.topk_per_query(q_off, it, 12)
        assert (out["topk_pos"] == tp).all(), name
        assert (bits(out["topk_score"]) == bits(ts)).all(), name
    idx.close()


def test_numpy_oracle_agrees_with_c_oracle_and_gpu(ffx, oracle_c):
    """The numpy restatement (np.sum(q*d, axis=1) itself) on a mid-size problem."""
    rng = np.random.default_rng(77)
    off, rows, vec = make_corpus(rng, 300, 6, 768, True)
    idx = ffx.DeviceIndex(768, capacity=len(vec))
    idx.stage(0, vec)
    idx.set_docs(off)
    qv = rng.standard_normal((5, 768)).astype(np.float32)
    q_off, cand, pair_q = make_pairs(rng, 5, 300, 50, 120)
    for mode in (fo.MODE_MAXP, fo.MODE_AVEP):
        ff_np = fo.score_pairs(vec, off, rows, pair_q, cand, qv, mode)
        ff_c = c_scores(oracle_c, vec, off, rows, pair_q, cand, qv, mode)
        ff_gpu = idx.rerank_host(mode, qv, q_off, cand)["ff"]
        assert (bits(ff_np) == bits(ff_c)).all() and (bits(ff_gpu) == bits(ff_np)).all()
    idx.close()


@pytest.mark.parametrize("mode", ["MAXP", "AVEP", "PASSAGE"])
def test_fused_topk_path_many_queries(ffx, oracle_c, mode):
    """nq >= 2 CTAs/SM x SMs switches to the fused score+interpolate+top-k kernel."""
    rng = np.random.default_rng(5)
    m = MODES[mode]
    off, rows, vec = make_corpus(rng, 2000, 8, 768, True)
    idx = ffx.DeviceIndex(768, capacity=len(vec))
    idx.stage(0, vec)
    idx.set_docs(off)
    nq = 320
    pool = len(vec) if m == fo.MODE_PASSAGE else 2000
    qv = rng.standard_normal((nq, 768)).astype(np.float32)
    q_off, cand, pair_q = make_pairs(rng, nq, pool, 0, 300)
    lex = (rng.integers(0, 8, len(cand)) * 4).astype(np.float32)  # heavy ties with alpha=1
    u_off, u_rows = units_for_mode(off, rows, len(vec), m)
    ff = c_scores(oracle_c, vec, u_off, u_rows, pair_q, cand, qv, m)
    for alpha, k in ((0.1, 100), (1.0, 37), (0.0, 512)):
        out = idx.rerank_host(m, qv, q_off, cand, lex, alpha, k, want_ff=True, want_int=True)
        it = fo.interpolate_f32(lex, ff, alpha)
        ts, tp = fo.topk_per_query(q_off, it, k)
        assert (bits(out["ff"]) == bits(ff)).all()
        assert (bits(out["int"]) == bits(it)).all()
        assert (out["topk_pos"] == tp).all()
        assert (bits(out["topk_score"]) == bits(ts)).all()
        # top-k only (no per-pair outputs) must give the same lists
        out2 = idx.rerank_host(m, qv, q_off, cand, lex, alpha, k, want_ff=False, want_int=False)
        assert (out2["topk_pos"] == tp).all() and (bits(out2["topk_score"]) == bits(ts)).all()
    idx.close()


@pytest.mark.parametrize("dim", [2560, 3072, 3584, 4096, 64, 96, 128, 192, 256])
def test_long_and_short_rows_fused_path(ffx, oracle_c, dim):
    """D >= 2560: 10-16 KB rows streamed from the TMA ring against a shared-memory copy of the
    query vector (one numpy leaf per lane); D <= 256: a quarter / half warp per row, 4 / 2 rows
    per warp step.  Both in the fused one-CTA-per-query form."""
    rng = np.random.default_rng(dim)
    off, rows, vec = make_corpus(rng, 300, 6, dim, True)
    idx = ffx.DeviceIndex(dim, capacity=len(vec))
    idx.stage(0, vec)
    assert (idx.read_rows([0, len(vec) - 1]) == vec[[0, len(vec) - 1]]).all()
    idx.set_docs(off)
    nq = 310
    qv = rng.standard_normal((nq, dim)).astype(np.float32)
    for mode in (fo.MODE_MAXP, fo.MODE_AVEP, fo.MODE_PASSAGE):
        pool = len(vec) if mode == fo.MODE_PASSAGE else 300
        q_off, cand, pair_q = make_pairs(rng, nq, pool, 0, 90)
        lex = (rng.integers(0, 8, len(cand)) * 4).astype(np.float32)
        u_off, u_rows = units_for_mode(off, rows, len(vec), mode)
        ff = c_scores(oracle_c, vec, u_off, u_rows, pair_q, cand, qv, mode)
        out = idx.rerank_host(mode, qv, q_off, cand, lex, 0.2, 20, want_ff=True, want_int=True)
        it = fo.interpolate_f32(lex, ff, 0.2)
        ts, tp = fo.topk_per_query(q_off, it, 20)
        assert (bits(out["ff"]) == bits(ff)).all()
        assert (out["topk_pos"] == tp).all() and (bits(out["topk_score"]) == bits(ts)).all()
    idx.close()


@pytest.mark.parametrize("dim", [5, 50, 100, 130, 200, 260, 300, 1000, 1280, 2000, 4000, 64, 128, 192, 384, 512])
def test_packed_kernel_fused_and_tiled(ffx, oracle_c, dim):
    """ffx_score_packed_kernel under its default options.  Dimensions without a uniform numpy tree
    (TreeDot, the tree as data: leaves of different lengths and depths, a tail after the last leaf,
    row stride padded to 16 bytes; 4 / 8 / 16 / 32 lanes per row) and the lane-major rows of up to
    512 elements (LaneMajorDot: 2 .. 16 lanes per row, 16 .. 2 rows per warp step): fused
    one-CTA-per-query launches, tiled launches, scattered documents, ring depths and batch sizes —
    bit for bit against the plain-C oracle; rows read back unchanged."""
    kernel_name = "ffx_score_packed_kernel<ffx::" + ("LaneMajorDot<" if dim in (64, 128, 192, 384, 512) else "TreeDot<")
    ffx.set_option("kernel", 0)  # the fixture's kernel = 1 / 2 would pick the whole-warp kernels at D = 384 / 512
    rng = np.random.default_rng(dim)
    for contiguous in (True, False):
        off, rows, vec = make_corpus(rng, 500, 7, dim, contiguous)
        idx = ffx.DeviceIndex(dim, capacity=len(vec))
        idx.stage(0, vec[:100])
        idx.stage(100, vec[100:])
        assert idx.has_fast_path
        some = rng.choice(len(vec), 50, replace=False)
        assert (idx.read_rows(some) == vec[some]).all()
        idx.set_docs(off, rows)
        nq = 310
        qv = rng.standard_normal((nq, dim)).astype(np.float32)
        for mode in (fo.MODE_MAXP, fo.MODE_AVEP, fo.MODE_FIRSTP, fo.MODE_PASSAGE):
            pool = len(vec) if mode == fo.MODE_PASSAGE else 500
            q_off, cand, pair_q = make_pairs(rng, nq, pool, 0, 120)
            lex = (rng.integers(0, 8, len(cand)) * 4).astype(np.float32)
            u_off, u_rows = units_for_mode(off, rows, len(vec), mode)
            ff = c_scores(oracle_c, vec, u_off, u_rows, pair_q, cand, qv, mode)
            it = fo.interpolate_f32(lex, ff, 0.2)
            ts, tp = fo.topk_per_query(q_off, it, 20)
            for stages, batch in ((0, 0), (2, 5), (7, 32)):
                ffx.set_option("tma_stages", stages)
                ffx.set_option("batch", batch)
                try:
                    for sub in (nq, 9):  # fused and tiled launches
                        n_sub = int(q_off[sub])
                        out = idx.rerank_host(mode, qv[:sub], q_off[:sub + 1], cand[:n_sub], lex[:n_sub], 0.2, 20,
                                              want_ff=True, want_int=True)
                        assert kernel_name in ffx.last_kernel()
                        assert ("true>" in ffx.last_kernel()) == (sub == nq)
                        assert (bits(out["ff"]) == bits(ff[:n_sub])).all(), (mode, stages, batch, sub)
                        assert (bits(out["int"]) == bits(it[:n_sub])).all()
                        assert (out["topk_pos"] == tp[:sub]).all() and (bits(out["topk_score"]) == bits(ts[:sub])).all()
                finally:
                    ffx.set_option("tma_stages", 0)
                    ffx.set_option("batch", 0)
        idx.close()


@pytest.mark.parametrize("stages,batch", [(2, 5), (3, 32), (16, 1)])
def test_tma_ring_and_batch_shapes(ffx, oracle_c, stages, batch):
    """The ring depth and the candidate batch are tuning knobs: results must not depend on them."""
    rng = np.random.default_rng(stages * 100 + batch)
    off, rows, vec = make_corpus(rng, 1500, 40, 768, True)
    idx = ffx.DeviceIndex(768, capacity=len(vec))
    idx.stage(0, vec)
    idx.set_docs(off)
    nq = 300
    qv = rng.standard_normal((nq, 768)).astype(np.float32)
    ffx.set_option("kernel", 2)
    ffx.set_option("tma_stages", stages)
    ffx.set_option("batch", batch)
    try:
        for mode in (fo.MODE_MAXP, fo.MODE_AVEP, fo.MODE_FIRSTP, fo.MODE_PASSAGE):
            pool = len(vec) if mode == fo.MODE_PASSAGE else 1500
            q_off, cand, pair_q = make_pairs(rng, nq, pool, 0, 90)
            lex = rng.uniform(0, 20, len(cand)).astype(np.float32)
            u_off, u_rows = units_for_mode(off, rows, len(vec), mode)
            ff = c_scores(oracle_c, vec, u_off, u_rows, pair_q, cand, qv, mode)
            ts, tp = fo.topk_per_query(q_off, fo.interpolate_f32(lex, ff, 0.3), 20)
            for sub in (nq, 7):  # fused and tiled launches
                o = q_off[:sub + 1]
                out = idx.rerank_host(mode, qv[:sub], o, cand[:o[-1]], lex[:o[-1]], 0.3, 20, want_ff=True)
                assert (bits(out["ff"]) == bits(ff[:o[-1]])).all()
                assert (out["topk_pos"] == tp[:sub]).all() and (bits(out["topk_score"]) == bits(ts[:sub])).all()
    finally:
        ffx.set_option("tma_stages", 0)
        ffx.set_option("batch", 0)
    idx.close()


def test_edge_cases(ffx):
    rng = np.random.default_rng(1)
    vec = rng.standard_normal((64, 768)).astype(np.float32)
    idx = ffx.DeviceIndex(768, capacity=64)
    idx.stage(0, vec)
    idx.set_docs(np.arange(0, 65, 2))
    qv = rng.standard_normal((3, 768)).astype(np.float32)
    # no queries / no pairs / empty query in the middle / k larger than the candidate count
    assert idx.rerank_host(fo.MODE_MAXP, qv[:0], [0], [], k=0)["ff"].shape == (0,)
    out = idx.rerank_host(fo.MODE_MAXP, qv, [0, 0, 0, 0], [], None, 0.0, 4)
    assert (out["topk_pos"] == -1).all() and np.isneginf(out["topk_score"]).all()
    out = idx.rerank_host(fo.MODE_MAXP, qv, [0, 3, 3, 5], [1, 2, 3, 4, 5], None, 0.0, 4)
    assert (out["topk_pos"][1] == -1).all() and (out["topk_pos"][0, 3] == -1)
    assert sorted(out["topk_pos"][0, :3].tolist()) == [0, 1, 2]
    assert (out["topk_pos"][2, 2:] == -1).all()
    # out-of-range candidates never reach the kernel
    with pytest.raises(ffx.FFXError):
        idx.rerank_host(fo.MODE_MAXP, qv, [0, 1, 1, 1], [32])
    with pytest.raises(ffx.FFXError):
        idx.rerank_host(fo.MODE_PASSAGE, qv, [0, 1, 1, 1], [64])
    with pytest.raises(ffx.FFXError):
        idx.rerank_host(fo.MODE_PASSAGE, qv, [0, 1, 1, 1], [-1])
    # a long document (more rows than a warp) and a very long candidate list (global-key top-k)
    idx.set_docs([0, 40, 64])
    ff = idx.rerank_host(fo.MODE_AVEP, qv, [0, 2, 2, 2], [0, 1])["ff"]
    s = (vec.astype(np.float32) @ qv[0])
    want = fo.score_pairs(vec, np.array([0, 40, 64]), np.arange(64), [0, 0], [0, 1], qv, fo.MODE_AVEP)
    assert (bits(ff) == bits(want)).all() and np.allclose(ff, [s[:40].mean(), s[40:].mean()], rtol=1e-4)
    idx.close()


def test_large_candidate_lists_use_global_keys(ffx, oracle_c):
    rng = np.random.default_rng(8)
    vec = rng.standard_normal((30000, 384)).astype(np.float32)
    idx = ffx.DeviceIndex(384, capacity=len(vec))
    idx.stage(0, vec)
    qv = rng.standard_normal((2, 384)).astype(np.float32)
    q_off = np.array([0, 20000, 20007])
    cand = np.concatenate([rng.permutation(30000)[:20000], np.arange(7)]).astype(np.int32)
    lex = np.round(rng.standard_normal(len(cand)) * 3).astype(np.float32)
    out = idx.rerank_host(fo.MODE_PASSAGE, qv, q_off, cand, lex, 0.9, 1000, want_ff=True)
    ff = c_scores(oracle_c, vec, np.arange(30001), np.arange(30000), np.repeat([0, 1], [20000, 7]),
                  cand, qv, fo.MODE_PASSAGE)
    ts, tp = fo.topk_per_query(q_off, fo.interpolate_f32(lex, ff, 0.9), 1000)
    assert (bits(out["ff"]) == bits(ff)).all()
    assert (out["topk_pos"] == tp).all() and (bits(out["topk_score"]) == bits(ts)).all()
    idx.close()


# ------------------------------------------------------------------------------------------
# storage
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dim", [768, 100, 2048, 3072, 128, 256])
def test_stage_read_roundtrip_and_growth(ffx, dim):
    rng = np.random.default_rng(dim)
    vec = rng.standard_normal((5000, dim)).astype(np.float32)
    idx = ffx.DeviceIndex(dim, capacity=1000)
    idx.stage(0, vec[:1000])
    idx.reserve(5000)  # index/memory.py:103-108 growth keeps contents
    idx.stage(1000, vec[1000:])
    assert len(idx) == 5000 and idx.capacity == 5000
    rows = rng.integers(0, 5000, 700)
    assert (idx.read_rows(rows) == vec[rows]).all()
    assert idx.read_rows([]).shape == (0, dim)
    with pytest.raises(ffx.FFXError):
        idx.read_rows([5000])
    with pytest.raises(ffx.FFXError):
        idx.stage(4999, vec[:2])
    idx.close()


def test_staging_larger_than_the_pinned_buffer(ffx):
    rng = np.random.default_rng(3)
    vec = rng.standard_normal((40000, 768)).astype(np.float32)  # 123 MB > 2 x 32 MB buffers
    idx = ffx.DeviceIndex(768, capacity=len(vec))
    idx.stage(0, vec)
    rows = np.concatenate([[0, 39999], rng.integers(0, 40000, 300)])
    assert (idx.read_rows(rows) == vec[rows]).all()
    idx.close()


@pytest.mark.parametrize("nq", [4, 320])
def test_short_lists_out_of_long_candidate_lists(ffx, oracle_c, nq):
    """k << candidates takes the radix-select path (select_largest_keys: the k-th largest key by
    8-bit digits, stable compaction, then a sort of only next_pow2(k) keys) — in the fused
    kernel's epilogue (nq = 320) and in ffx_topk_kernel (nq = 4).  Coarse lexical scores with
    alpha = 1 make thousands of exactly equal scores: the cut must fall by position."""
    rng = np.random.default_rng(nq)
    off, rows, vec = make_corpus(rng, 6000, 3, 384, True)
    idx = ffx.DeviceIndex(384, capacity=len(vec))
    idx.stage(0, vec)
    idx.set_docs(off)
    qv = rng.standard_normal((nq, 384)).astype(np.float32)
    q_off, cand, pair_q = make_pairs(rng, nq, 6000, 1100, 5000)
    ff = c_scores(oracle_c, vec, off, rows, pair_q, cand, qv, fo.MODE_MAXP)
    for alpha, lex in ((1.0, rng.integers(0, 5, len(cand)).astype(np.float32)),
                       (0.3, rng.uniform(0, 20, len(cand)).astype(np.float32))):
        it = fo.interpolate_f32(lex, ff, alpha)
        for k in (1, 7, 100, 256, 275):
            out = idx.rerank_host(fo.MODE_MAXP, qv, q_off, cand, lex, alpha, k, want_ff=False, want_int=False)
            ts, tp = fo.topk_per_query(q_off, it, k)
            assert (out["topk_pos"] == tp).all(), (alpha, k)
            assert (bits(out["topk_score"]) == bits(ts)).all(), (alpha, k)
    idx.close()


# ------------------------------------------------------------------------------------------
# PQ / OPQ asymmetric distance
# ------------------------------------------------------------------------------------------
def adc_case(ffx, rng, M, Ks, Ds, rotate, contiguous, n_docs=500, max_psg=7):
    D = M * Ds
    off, rows, _ = make_corpus(rng, n_docs, max_psg, 4, contiguous)
    n_rows = int(off[-1])
    codes = rng.integers(0, Ks, (n_rows, M)).astype(np.uint8)
    cw = rng.standard_normal((M, Ks, Ds)).astype(np.float32)
    R = np.linalg.qr(rng.standard_normal((D, D)))[0].astype(np.float32) if rotate else None
    idx = ffx.DeviceIndex(M, capacity=n_rows, row_kind=ffx.ROWS_PQ_U8)
    idx.stage(0, codes)
    assert (idx.read_rows([0, n_rows - 1]) == codes[[0, n_rows - 1]]).all()
    idx.set_docs(off, None if contiguous else rows)
    idx.set_pq(cw, R)
    dec = fo.pq_decode(codes, cw) if R is None else fo.opq_decode(codes, cw, R)
    return idx, off, rows, n_rows, dec


def check_adc(idx, rng, off, rows, n_rows, dec, qv, n_docs, lo, hi, k, modes=MODES):
    row_norm = np.linalg.norm(dec, axis=1)
    nq = len(qv)
    for name, mode in modes.items():
        pool = n_rows if mode == fo.MODE_PASSAGE else n_docs
        q_off, cand, pair_q = make_pairs(rng, nq, pool, lo, hi)
        lex = rng.uniform(0, 20, len(cand)).astype(np.float32)
        out = idx.rerank_host(mode, qv, q_off, cand, lex, 0.2, k, want_ff=True, want_int=True)
        u_off, u_rows = units_for_mode(off, rows, n_rows, mode)
        want = fo.score_pairs(dec.astype(np.float64), u_off, u_rows, pair_q, cand, qv.astype(np.float64), mode)
        scale = np.array([row_norm[u_rows[u_off[c]:u_off[c + 1]]].max() for c in cand]) * \
            np.linalg.norm(qv, axis=1)[pair_q]
        err = np.abs(out["ff"] - want)
        assert (err <= 1e-5 * np.abs(want) + 1e-5 * scale).all(), name
        it = fo.interpolate_f32(lex, out["ff"], 0.2)
        assert (bits(out["int"]) == bits(it)).all()
        ts, tp = fo.topk_per_query(q_off, it, k)
        assert (out["topk_pos"] == tp).all() and (bits(out["topk_score"]) == bits(ts)).all()
        # top-k only: same lists without the per-pair outputs
        out2 = idx.rerank_host(mode, qv, q_off, cand, lex, 0.2, k, want_ff=False, want_int=False)
        assert (out2["topk_pos"] == tp).all() and (bits(out2["topk_score"]) == bits(ts)).all()


@pytest.mark.parametrize("adc_kernel", [1, 2, 3])
@pytest.mark.parametrize("rotate", [False, True])
@pytest.mark.parametrize("M,Ks,Ds", [(96, 256, 8), (64, 256, 4), (128, 256, 2), (32, 50, 4), (80, 100, 3),
                                     (8, 16, 4), (12, 256, 3)])
def test_adc_matches_decode_then_dot(ffx, M, Ks, Ds, rotate, adc_kernel):
    """quantizer/nanopq.py:43-44,111-112 + index/base.py:292-303.  ADC sums M LUT entries
    instead of D products, so scores differ from decode-then-dot by reassociation only:
    tolerance rtol 1e-5 (north_star) + atol 1e-5 * |q| * |d|.  All three ADC kernels: generic
    thread-per-row (adc=1), warp-per-row with conflict-free tables for M = 64..128 (adc=2) and
    XOR-swizzled thread-per-row for M % 32 == 0 (adc=3, the default where it applies)."""
    ffx.set_option("adc", adc_kernel)
    try:
        rng = np.random.default_rng(M + Ks)
        for contiguous in (True, False):
            idx, off, rows, n_rows, dec = adc_case(ffx, rng, M, Ks, Ds, rotate, contiguous)
            qv = rng.standard_normal((6, M * Ds)).astype(np.float32)
            check_adc(idx, rng, off, rows, n_rows, dec, qv, 500, 10, 200, 10)
            idx.close()
    finally:
        ffx.set_option("adc", 0)


@pytest.mark.parametrize("adc_kernel", [2, 3])
@pytest.mark.parametrize("M,Ks,Ds", [(96, 256, 8), (64, 64, 4)])
def test_adc_fused_topk_many_queries(ffx, M, Ks, Ds, adc_kernel):
    """nq >= #SMs switches the fast ADC kernels to their fused score+interpolate+top-k form
    (one CTA per query, sort keys built over the dead tables); documents of up to 70 passages
    cross the 32-row blocks of the kernels."""
    ffx.set_option("adc", adc_kernel)
    try:
        rng = np.random.default_rng(M)
        idx, off, rows, n_rows, dec = adc_case(ffx, rng, M, Ks, Ds, True, True, n_docs=700, max_psg=70)
        qv = rng.standard_normal((200, M * Ds)).astype(np.float32)
        check_adc(idx, rng, off, rows, n_rows, dec, qv, 700, 0, 300, 64,
                  modes={"MAXP": fo.MODE_MAXP, "AVEP": fo.MODE_AVEP, "PASSAGE": fo.MODE_PASSAGE})
        idx.close()
    finally:
        ffx.set_option("adc", 0)


@pytest.mark.parametrize("adc_kernel", [2, 3])
@pytest.mark.parametrize("lo,hi", [(4200, 5000), (2100, 4000)])
def test_adc_fused_long_lists_use_the_register_sort(ffx, adc_kernel, lo, hi):
    """Fused ADC kernels with thousands of candidates per query sort their keys with the
    register / shuffle / shared-memory hybrid bitonic network (n = 4, 8 or 16 keys per thread).
    Lexical ties are heavy (alpha close to 1), so the tie-by-position rule is exercised; the
    ranked lists are checked against the oracle's ordering of the kernel's own scores."""
    ffx.set_option("adc", adc_kernel)
    try:
        rng = np.random.default_rng(lo)
        M, Ks, Ds = 64, 16, 2
        idx, off, rows, n_rows, dec = adc_case(ffx, rng, M, Ks, Ds, False, True, n_docs=6000, max_psg=3)
        nq = 150
        qv = rng.standard_normal((nq, M * Ds)).astype(np.float32)
        q_off, cand, pair_q = make_pairs(rng, nq, 6000, lo, hi)
        lex = rng.integers(0, 6, len(cand)).astype(np.float32)
        for alpha, k in ((0.999, int(np.diff(q_off).max())), (1.0, 1000), (0.1, 10)):
            out = idx.rerank_host(fo.MODE_MAXP, qv, q_off, cand, lex, alpha, k, want_ff=True, want_int=True)
            assert (bits(out["int"]) == bits(fo.interpolate_f32(lex, out["ff"], alpha))).all()
            ts, tp = fo.topk_per_query(q_off, out["int"], k)
            assert (out["topk_pos"] == tp).all() and (bits(out["topk_score"]) == bits(ts)).all()
        idx.close()
    finally:
        ffx.set_option("adc", 0)


@pytest.mark.parametrize("lo,hi", [(2048, 2400), (4700, 5000), (5900, 6145), (7000, 8192), (9000, 12000)])
def test_long_lists_sorted_by_the_radix_epilogue(ffx, oracle_c, lo, hi):
    """Lists of >= 2048 candidates whose every pair is ranked, with k above a quarter of the list,
    are ordered by the stable LSD radix sort on the score half of the keys (fused fp32 kernel
    and the stand-alone top-k kernel); lists that leave no room for its counters, longer than
    its keys-per-thread limit, or with k small, keep the bitonic / select paths.  Heavy ties
    (8 distinct lexical scores, alpha = 1) pin the tie-by-position rule."""
    rng = np.random.default_rng(lo)
    off, rows, vec = make_corpus(rng, 13000, 2, 384, True)
    idx = ffx.DeviceIndex(384, capacity=len(vec))
    idx.stage(0, vec)
    idx.set_docs(off)
    nq = 300
    qv = rng.standard_normal((nq, 384)).astype(np.float32)
    q_off, cand, pair_q = make_pairs(rng, nq, 13000, lo, hi)
    lex = (rng.integers(0, 8, len(cand)) * 0.5).astype(np.float32)
    u_off, u_rows = units_for_mode(off, rows, len(vec), fo.MODE_MAXP)
    ff = c_scores(oracle_c, vec, u_off, u_rows, pair_q, cand, qv, fo.MODE_MAXP)
    longest = int(np.diff(q_off).max())
    for alpha, k in ((1.0, longest), (0.3, longest), (1.0, longest // 2), (0.3, 100)):
        it = fo.interpolate_f32(lex, ff, alpha)
        ts, tp = fo.topk_per_query(q_off, it, k)
        out = idx.rerank_host(fo.MODE_MAXP, qv, q_off, cand, lex, alpha, k, want_ff=False, want_int=False)
        assert (out["topk_pos"] == tp).all() and (bits(out["topk_score"]) == bits(ts)).all()
        alone = idx.interpolate_topk_host(lex, ff, q_off, alpha, k)
        assert (alone["topk_pos"] == tp).all() and (bits(alone["topk_score"]) == bits(ts)).all()
    idx.close()


@pytest.mark.parametrize("M,Ks,Ds", [(96, 256, 8), (32, 64, 4), (64, 256, 16), (64, 100, 8)])
def test_adc_tables_are_the_same_from_all_three_builders(ffx, M, Ks, Ds):
    """The per-query tables of the XOR-swizzled ADC kernel come from a tiled kernel (query slices
    broadcast), a thread-per-entry kernel (shapes the tiles do not cover, here Ks = 100) or are
    built inside the scoring kernel: same fmaf chain per entry, so the scores agree bit for bit."""
    rng = np.random.default_rng(M + Ks + Ds)
    off, rows, _ = make_corpus(rng, 400, 6, 4, True)
    n_rows = int(off[-1])
    idx = ffx.DeviceIndex(M, capacity=n_rows, row_kind=ffx.ROWS_PQ_U8)
    idx.stage(0, rng.integers(0, Ks, (n_rows, M)).astype(np.uint8))
    idx.set_docs(off)
    D = M * Ds
    R = np.linalg.qr(rng.standard_normal((D, D)))[0].astype(np.float32)
    idx.set_pq(rng.standard_normal((M, Ks, Ds)).astype(np.float32), R)
    nq = 333  # not a multiple of the 32 queries a table CTA walks
    qv = rng.standard_normal((nq, D)).astype(np.float32)
    q_off, cand, _ = make_pairs(rng, nq, 400, 1, 90)
    lex = rng.uniform(0, 20, len(cand)).astype(np.float32)
    outs = []
    try:
        for how in (0, 1, 2):
            ffx.set_option("adc_lut", how)
            for mode in (fo.MODE_MAXP, fo.MODE_AVEP):
                outs.append(idx.rerank_host(mode, qv, q_off, cand, lex, 0.2, 30, want_ff=True, want_int=True))
                assert "ffx_adc_xor_kernel" in ffx.last_kernel()
    finally:
        ffx.set_option("adc_lut", 0)
        idx.close()
    for a, b in ((0, 2), (0, 4), (1, 3), (1, 5)):
        assert (bits(outs[a]["ff"]) == bits(outs[b]["ff"])).all() and (outs[a]["topk_pos"] == outs[b]["topk_pos"]).all()


def test_adc_fused_launches_are_bit_reproducible(ffx):
    """The XOR ADC kernel refills a warp's code slots with bulk copies (async proxy) right after
    the lanes have loaded their rows through the generic proxy; without a cross-proxy fence a copy
    could overtake a queued load (about one wrong row in 10^6 under a saturated LSU).  Same
    inputs, repeated launches at a size where that showed: every per-pair score and every list
    must come out bit-identical, through the device launch and the pipelined host path."""
    rng = np.random.default_rng(11)
    M, Ks, Ds = 96, 256, 8
    n_docs = 120_000
    cnt = rng.integers(1, 13, n_docs)
    off = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
    idx = ffx.DeviceIndex(M, capacity=int(off[-1]), row_kind=ffx.ROWS_PQ_U8)
    idx.stage(0, rng.integers(0, Ks, (int(off[-1]), M), dtype=np.uint8))
    idx.set_docs(off)
    idx.set_pq(rng.standard_normal((M, Ks, Ds)).astype(np.float32), None)
    nq, C = 600, 5000
    qv = rng.standard_normal((nq, M * Ds)).astype(np.float32)
    q_off = (np.arange(nq + 1, dtype=np.int64) * C)
    cand = np.concatenate([rng.choice(n_docs, C, replace=False) for _ in range(nq)]).astype(np.int32)
    lex = (rng.random(nq * C) * 20).astype(np.float32)
    first = idx.rerank_host(fo.MODE_AVEP, qv, q_off, cand, lex, 0.1, C, want_ff=False, want_int=True)
    ts, tp = fo.topk_per_query(q_off, first["int"], C)
    assert (first["topk_pos"] == tp).all() and (bits(first["topk_score"]) == bits(ts)).all()
    for _ in range(3):
        again = idx.rerank_host(fo.MODE_AVEP, qv, q_off, cand, lex, 0.1, C, want_ff=False, want_int=True)
        assert (bits(again["int"]) == bits(first["int"])).all()
        assert (again["topk_pos"] == first["topk_pos"]).all()
        lists_only = idx.rerank_host(fo.MODE_AVEP, qv, q_off, cand, lex, 0.1, C, want_ff=False, want_int=False)
        assert (lists_only["topk_pos"] == first["topk_pos"]).all()
    idx.close()


def test_host_pipeline_matches_single_launch_opq(ffx):
    """ffx_rerank_host cuts many queries into chunks that run on alternating streams; every
    chunk must rotate its OPQ queries into its own buffer (a shared one is overwritten by the
    next chunk's rotation while this chunk's tables are still being built).  All lists of the
    pipelined host path are compared with one single launch over device buffers."""
    import torch

    rng = np.random.default_rng(77)
    M, Ks, Ds = 32, 16, 2
    idx, off, rows, n_rows, dec = adc_case(ffx, rng, M, Ks, Ds, True, True, n_docs=2000, max_psg=4)
    nq, C, k = 2400, 450, 450
    qv = rng.standard_normal((nq, M * Ds)).astype(np.float32)
    cand = np.concatenate([rng.choice(2000, C, replace=False) for _ in range(nq)]).astype(np.int32)
    q_off = np.arange(nq + 1, dtype=np.int64) * C
    lex = rng.uniform(0, 20, nq * C).astype(np.float32)
    host = idx.rerank_host(fo.MODE_AVEP, qv, q_off, cand, lex, 0.3, k, want_ff=True)
    dev = torch.device("cuda", 0)
    t = {n: torch.from_numpy(a).to(dev) for n, a in (("qv", qv), ("off", q_off), ("cand", cand), ("lex", lex))}
    ff = torch.zeros(nq * C, device=dev)
    ts = torch.empty((nq, k), device=dev)
    tp = torch.empty((nq, k), device=dev, dtype=torch.int32)
    torch.cuda.synchronize()
    idx.rerank_device(fo.MODE_AVEP, t["qv"].data_ptr(), nq, t["off"].data_ptr(), t["cand"].data_ptr(),
                      t["lex"].data_ptr(), 0.3, k, C, ff.data_ptr(), 0, ts.data_ptr(), tp.data_ptr())
    idx.sync()
    assert (bits(host["ff"]) == bits(ff.cpu().numpy())).all()
    assert (host["topk_pos"] == tp.cpu().numpy()).all()
    assert (bits(host["topk_score"]) == bits(ts.cpu().numpy())).all()
    idx.close()


# ------------------------------------------------------------------------------------------
# shard merge (doc-id-range sharded corpora)
# ------------------------------------------------------------------------------------------
def test_merge_topk_equals_global_topk(ffx):
    import torch

    rng = np.random.default_rng(21)
    nq, C, k, S = 50, 600, 64, 4
    scores = np.round(rng.standard_normal((nq, C)) * 4).astype(np.float32) / 2  # ties across shards
    q_off = np.arange(nq + 1) * C
    ws, wp = fo.topk_per_query(q_off, scores.ravel(), k)
    sh_s = np.full((S, nq, k), -np.inf, np.float32)
    sh_p = np.full((S, nq, k), -1, np.int32)
    owner = rng.integers(0, S, (nq, C))
    for s in range(S):
        for q in range(nq):
            pos = np.nonzero(owner[q] == s)[0]
            order = pos[np.argsort(-scores[q, pos], kind="stable")][:k]
            sh_s[s, q, :len(order)] = scores[q, order]
            sh_p[s, q, :len(order)] = order
    d_s, d_p = torch.from_numpy(sh_s).cuda(), torch.from_numpy(sh_p).cuda()
    o_s = torch.empty((nq, k), dtype=torch.float32, device="cuda")
    o_p = torch.empty((nq, k), dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    ffx.merge_topk(0, d_s.data_ptr(), d_p.data_ptr(), S, nq, k, o_s.data_ptr(), o_p.data_ptr())
    torch.cuda.synchronize()
    assert (o_p.cpu().numpy() == wp).all() and (bits(o_s.cpu().numpy()) == bits(ws)).all()
