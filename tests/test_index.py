"""The `fast_forward` drop-in API on the GPU — the reference's backend-parametrised index
suite (tests/test_index.py:49-441 there) re-stated for the HBM-resident back-ends, with the
expected frames taken from the unmodified reference (tests/golden/)."""

import itertools

import numpy as np
import pandas as pd
import pytest

import ff_oracle as fo

pytestmark = pytest.mark.gpu

QUERIES = {"q1": "query 1", "q2": "query 2"}
DOC = ["d0", "d0", "d1", "d2", "d3"]
PSG = ["p0", "p1", "p2", "p3", "p4"]
V = np.tril(np.ones((5, 5), dtype=np.int64))
DOC_RUN = {"q1": {"d0": 100, "d1": 2, "d2": 3, "d3": 200}, "q2": {"d0": 400, "d1": 5, "d2": 6, "d3": 800}}
PSG_RUN = {"q1": {"p0": 100, "p1": 2, "p2": 3, "p3": 4, "p4": 5},
           "q2": {"p0": 500, "p1": 6, "p2": 7, "p3": 8, "p4": 9}}


@pytest.fixture(scope="module")
def ff():
    import __graft_entry__ as g

    g.build()
    import fast_forward

    return fast_forward


@pytest.fixture(scope="module")
def api(ff):
    from fast_forward.encoder import LambdaEncoder, TableEncoder
    from fast_forward.index import InMemoryIndex, Mode
    from fast_forward.quantizer import NanoOPQ, NanoPQ

    class Api:
        pass

    a = Api()
    a.Ranking, a.InMemoryIndex, a.Mode = ff.Ranking, InMemoryIndex, Mode
    a.LambdaEncoder, a.TableEncoder, a.NanoPQ, a.NanoOPQ = LambdaEncoder, TableEncoder, NanoPQ, NanoOPQ
    a.ones = LambdaEncoder(lambda _: np.array([1, 1, 1, 1, 1]))
    a.new = lambda **kw: InMemoryIndex(**kw)
    return a


def same_frame(r, want):
    df = r._df
    assert df["q_id"].tolist() == want["q_id"] and df["id"].tolist() == want["id"]
    assert df["score"].to_numpy().astype(np.float32).view(np.uint32).tolist() == want["score_bits"]


def fill(index, which):
    if which == "full":
        index.add(V, doc_ids=DOC, psg_ids=PSG)
    else:  # reference tests/test_index.py:58-69: ids partially missing, docs split over adds
        index.add(V, doc_ids=[None, None] + DOC[2:], psg_ids=PSG[:-2] + [None, None])
        index.add(V[:2], doc_ids=DOC[:2])
        index.add(V[-2:], psg_ids=PSG[-2:])
    return index


def assert_same_vectors(vecs, ids, want_vecs, want_ids):
    """order-insensitive, like the reference helper (tests/test_index.py:667-683)."""
    assert len(ids) == len(want_ids) == len(vecs)
    # the store keeps float32: the expected rows are the float32 roundings of the input
    got = sorted(zip(ids, map(tuple, np.asarray(vecs, np.float32).tolist())))
    want = sorted(zip(want_ids, map(tuple, np.asarray(want_vecs, np.float32).tolist())))
    assert got == want


@pytest.mark.parametrize("which", ["full", "partial"])
def test_modes_match_the_reference(api, golden_kat, which):
    index = fill(api.new(query_encoder=api.ones), which)
    doc_rank = api.Ranking.from_run(DOC_RUN, queries=QUERIES)
    psg_rank = api.Ranking.from_run(PSG_RUN, queries=QUERIES)
    for mode in (api.Mode.MAXP, api.Mode.FIRSTP, api.Mode.AVEP):
        index.mode = mode
        out = index(doc_rank)
        assert out.has_queries and out.name == "fast-forward"
        same_frame(out, golden_kat[f"{which}/{mode.name}"])
    index.mode = api.Mode.PASSAGE
    same_frame(index(psg_rank), golden_kat[f"{which}/PASSAGE"])


def test_properties(api):
    full = fill(api.new(query_encoder=api.ones), "full")
    partial = fill(api.new(query_encoder=api.ones), "partial")
    assert full.doc_ids == set(DOC) and full.psg_ids == set(PSG) and len(full) == 5 and full.dim == 5
    assert partial.doc_ids == set(DOC) and partial.psg_ids == set(PSG) and len(partial) == 9
    docs_only = api.new()
    docs_only.add(V, doc_ids=DOC)
    assert docs_only.psg_ids == set() and docs_only.doc_ids == set(DOC)
    empty = api.new()
    assert len(empty) == 0 and empty.dim is None and empty.doc_ids == set()


def test_add_and_retrieve_while_growing(api):
    index = api.new(init_size=32, alloc_size=32)
    rng = np.random.default_rng(0)
    data = rng.normal(size=(80, 16))
    doc_ids = [f"doc_{i // 2}" for i in range(80)]
    psg_ids = [f"psg_{i}" for i in range(80)]
    for lo, hi in [(0, 8), (8, 24), (24, 80)]:
        index.add(data[lo:hi], doc_ids=doc_ids[lo:hi], psg_ids=psg_ids[lo:hi])
        assert len(index) == hi
        index.mode = api.Mode.PASSAGE
        assert_same_vectors(*index._get_vectors(psg_ids[lo:hi]), data[lo:hi], psg_ids[lo:hi])
        index.mode = api.Mode.MAXP
        docs = [f"doc_{i}" for i in range(lo // 2, hi // 2)]
        assert_same_vectors(*index._get_vectors(docs), data[lo:hi], doc_ids[lo:hi])
        index.mode = api.Mode.FIRSTP
        assert_same_vectors(*index._get_vectors(docs), data[lo:hi:2], doc_ids[lo:hi:2])
    index.consolidate()
    index.mode = api.Mode.PASSAGE
    assert_same_vectors(*index._get_vectors(psg_ids), data, psg_ids)
    vecs, ids = index._get_vectors([])
    assert len(vecs) == 0 and ids == []


def test_errors(api):
    idx = api.new()
    with pytest.raises(ValueError):
        idx.add(V, doc_ids=None, psg_ids=None)
    with pytest.raises(ValueError):
        idx.add(V, doc_ids=DOC[:-2])
    with pytest.raises(ValueError):
        idx.add(V, psg_ids=PSG[:-2])
    with pytest.raises(ValueError):
        idx.add(V, doc_ids=[None] + DOC[1:], psg_ids=[None] + PSG[1:])
    idx.add(V[:1], psg_ids=PSG[:1])
    with pytest.raises(RuntimeError):
        idx.add(V[:1], psg_ids=PSG[:1])
    assert len(idx) == 1  # nothing was added by the failing call
    with pytest.raises(RuntimeError):
        idx.encode_queries(["test"])
    wrong = api.new()
    wrong.add(np.array([[0, 0], [1, 1]]), doc_ids=["d1", "d2"])
    with pytest.raises(ValueError):
        wrong.add(np.array([[0, 0, 0], [1, 1, 1]]), doc_ids=["d3", "d4"])

    full = fill(api.new(query_encoder=api.ones), "full")
    with pytest.raises(ValueError):
        full(api.Ranking.from_run(DOC_RUN))  # no queries attached
    rank = api.Ranking.from_run(DOC_RUN, queries=QUERIES)
    with pytest.raises(ValueError):
        full(rank, early_stopping=10, early_stopping_alpha=None, early_stopping_depths=(1,))
    with pytest.raises(ValueError):
        full(rank, early_stopping=10, early_stopping_alpha=0.5, early_stopping_depths=None)
    pq = api.NanoPQ(2, 8)
    pq.fit(np.random.default_rng(0).normal(size=(16, 16)).astype(np.float32))
    with pytest.raises(RuntimeError):
        full.quantizer = pq
    with pytest.raises(IndexError, match="ID dx not found in the index."):
        full(api.Ranking.from_run({"q1": {"d0": 100, "dx": 2}}, queries=QUERIES))
    full.mode = api.Mode.PASSAGE
    with pytest.raises(IndexError):
        full(rank)  # document ids are not passage ids


def test_batch_size(api, golden_kat):
    full = fill(api.new(query_encoder=api.ones), "full")
    g = golden_kat["batch/input"]
    run = {}
    for q, i, s in zip(g["q_id"], g["id"], g["score"]):
        run.setdefault(q, {})[i] = s
    r = api.Ranking.from_run(run, queries={f"q{n}": f"query {n}" for n in range(1, 6)})
    expected = full(r)
    same_frame(expected, golden_kat["batch/none"])
    for bs in (1, 2, 5, 10):  # 5 divides the query count: the reference crashes, we do not
        assert full(r, batch_size=bs) == expected
    same_frame(full(r, batch_size=2), golden_kat["batch/2"])


def test_early_stopping(api, golden_kat):
    es = api.new(query_encoder=api.LambdaEncoder(lambda q: np.array([10, 10])), mode=api.Mode.PASSAGE)
    es.add(np.stack([[1, 0], [1, 1]] * 10), psg_ids=[f"p{i}" for i in range(20)])
    r = api.Ranking(pd.DataFrame([{"q_id": q, "query": q, "id": f"p{i}", "score": i}
                                  for i in range(20) for q in ("q1", "q2")]))
    out = es(r, early_stopping=5, early_stopping_alpha=0.5, early_stopping_depths=(2, 5, 10, 20))
    same_frame(out, golden_kat["es/5_0.5_2-5-10-20"])
    out = es(r, early_stopping=3, early_stopping_alpha=0.2, early_stopping_depths=(4, 8, 20))
    same_frame(out, golden_kat["es/3_0.2_4-8-20"])


@pytest.mark.parametrize("key", ["es21", "es22"])
def test_early_stopping_random_golden(api, golden_es, key):
    """Seeded early-stopping cases against the reference's frames: es21 (D=768) runs the whole
    depth walk in one launch (ffx_rerank_early_stop), es22 (D=100, a dimension without a uniform
    summation tree: ffx_score_packed_kernel with TreeDot) as a stream-ordered sequence of launches per depth."""
    from fast_forward import _ffx

    meta, arrays = golden_es
    case = meta[key]
    vec, qvecs = arrays[f"{key}/vectors"], arrays[f"{key}/qvecs"]
    enc = api.TableEncoder({f"text {i}": qvecs[i] for i in range(len(qvecs))})
    index = api.new(query_encoder=enc, init_size=len(vec))
    index.add(vec, doc_ids=case["doc_ids"], psg_ids=case["psg_ids"])
    assert index._device().has_fast_path  # every dimension up to 4096 has a warp-per-row kernel
    for mode in api.Mode:
        entry = case["modes"][mode.name]
        fs = entry["first_stage"]
        first = api.Ranking(pd.DataFrame({"q_id": fs["q_id"], "id": fs["id"], "score": fs["score"]}),
                            queries=case["queries"])
        index.mode = mode
        for st in entry["settings"]:
            before = _ffx.launch_count()
            out = index(first, early_stopping=st["cutoff"], early_stopping_alpha=st["alpha"],
                        early_stopping_depths=tuple(st["depths"]))
            same_frame(out, st["out"])
            if key == "es21":
                assert _ffx.launch_count() - before == 1  # the whole walk is one kernel
            # batches of queries walk independently
            assert index(first, early_stopping=st["cutoff"], early_stopping_alpha=st["alpha"],
                         early_stopping_depths=tuple(st["depths"]), batch_size=4) == out


@pytest.mark.parametrize("dim", [128, 3072])
def test_early_stopping_falls_back_to_the_host_walk(api, dim):
    """Dimensions whose rows are scored by the TMA-staged kernel only (short and long rows) have
    no one-launch depth walk: ffx_rerank_early_stop answers FFX_ERR_UNSUPPORTED and the shell
    walks the depths itself — same rows, same scores as the oracle's restatement predicts from
    the full scores."""
    rng = np.random.default_rng(dim)
    n_docs, nq, C = 300, 12, 120
    vec = rng.standard_normal((n_docs, dim)).astype(np.float32)
    qv = rng.standard_normal((nq, dim)).astype(np.float32)
    index = api.new(query_encoder=api.TableEncoder({f"t{i}": qv[i] for i in range(nq)}), mode=api.Mode.PASSAGE)
    index.add(vec, psg_ids=[f"p{i}" for i in range(n_docs)])
    steep = rng.uniform(0.93, 0.999, nq)  # first-stage scores fall off at a different rate per query
    rows = [{"q_id": f"q{q:02d}", "query": f"t{q}", "id": f"p{int(p)}",
             "score": float(np.float32(3.0 * np.sqrt(dim)) * np.float32(steep[q]) ** r)}
            for q in range(nq) for r, p in enumerate(rng.choice(n_docs, C, replace=False))]
    first = api.Ranking(pd.DataFrame(rows))
    full = index(first)
    es = index(first, early_stopping=5, early_stopping_alpha=0.5, early_stopping_depths=(10, 30, 120))
    src = first._df
    ff_of = {(q, i): s for q, i, s in zip(full._df["q_id"], full._df["id"], full._df["score"])}
    ff = np.array([ff_of[(q, i)] for q, i in zip(src["q_id"], src["id"])], np.float32)
    q_off = np.arange(nq + 1) * C
    want = fo.early_stopping_depth(q_off, src["score"].to_numpy(), ff, 0.5, 5, (10, 30, 120))
    assert 1 < len(set(want.tolist()))  # queries stop at different depths
    got = es._df.groupby("q_id").size()
    assert [int(got[q]) for q in pd.unique(src["q_id"])] == want.tolist()
    es_of = {(q, i): s for q, i, s in zip(es._df["q_id"], es._df["id"], es._df["score"])}
    assert all(ff_of[key] == s for key, s in es_of.items())


def test_iteration(api):
    for kw in ({"init_size": 2, "alloc_size": 2}, {"init_size": 5}):
        index = api.new(**kw)
        index.add(V, doc_ids=DOC, psg_ids=PSG)
        for bs in (1, 3, 5, 10):
            vecs, docs, psgs = zip(*index.batch_iter(bs))
            assert np.array_equal(np.concatenate(vecs), V)
            assert list(itertools.chain.from_iterable(docs)) == DOC
            assert list(itertools.chain.from_iterable(psgs)) == PSG
        assert [d for _, d, _ in index] == DOC


@pytest.mark.parametrize("key", ["s11", "s12", "s13", "s14"])
def test_random_golden_through_the_public_api(api, golden_random, key):
    """index(ranking) -> interpolate -> cut, compared frame by frame with the reference, via
    the plain three-call idiom (README.md:48-51 of the reference) AND the fused `rerank`."""
    meta, arrays = golden_random
    case = meta[key]
    vec, qvecs = arrays[f"{key}/vectors"], arrays[f"{key}/qvecs"]
    enc = api.TableEncoder({f"text {i}": qvecs[i] for i in range(len(qvecs))})
    index = api.new(query_encoder=enc, init_size=16, alloc_size=64)
    sp = case["split"]
    index.add(vec[:sp], doc_ids=case["doc_ids"][:sp], psg_ids=case["psg_ids"][:sp])
    index.add(vec[sp:], doc_ids=case["doc_ids"][sp:], psg_ids=case["psg_ids"][sp:])
    for mode in api.Mode:
        g = case["modes"][mode.name]
        fs = g["first_stage"]
        first = api.Ranking(pd.DataFrame({"q_id": fs["q_id"], "id": fs["id"], "score": fs["score"]}),
                            queries=case["queries"])
        same_frame(first, fs)
        index.mode = mode
        out = index(first)
        same_frame(out, g["ff"])
        inter = first.interpolate(out, case["alpha"])
        assert out._origin is not None  # the GPU interpolate path was taken
        same_frame(inter, g["interpolated"])
        same_frame(inter.cut(case["cutoff"]), g["cut"])
        same_frame(index.rerank(first, case["alpha"]), g["interpolated"])
        same_frame(index.rerank(first, case["alpha"], cutoff=case["cutoff"]), g["cut"])
        # alpha = 1 makes the interpolated score the coarse lexical score: long runs of equal
        # scores straddle every cut, and the reference keeps the smaller ids (its outer merge
        # sorts the keys).  The fused path must cut exactly there too.
        for cut in (1, 3, case["cutoff"], 12):
            want = first.interpolate(api.Ranking(out._df), 1.0).cut(cut)  # generic (merge) route
            got = index.rerank(first, 1.0, cutoff=cut)
            assert got._df["id"].tolist() == want._df["id"].tolist()
            assert got._df["score"].tolist() == want._df["score"].tolist()
        # a ranking that did not come from this index takes the generic (merge) route
        same_frame(first.interpolate(api.Ranking(out._df), case["alpha"]), g["interpolated"])


def test_quantized_index(api):
    """reference tests/test_index.py:389-399 (shapes) + scores from codes == decode-then-dot
    (rtol 1e-5 + atol 1e-5*|q||d|; ADC reorders the sum)."""
    rng = np.random.default_rng(4)
    for cls, kw in ((api.NanoPQ, {"iter": 3}), (api.NanoOPQ, {"pq_iter": 3, "rotation_iter": 2})):
        pq = cls(4, 16)
        pq.fit(rng.normal(size=(64, 16)).astype(np.float32), **kw)
        qv = rng.normal(size=(2, 16)).astype(np.float32)
        index = api.new(query_encoder=api.TableEncoder({"query 1": qv[0], "query 2": qv[1]}), quantizer=pq)
        x = rng.normal(size=(5, 16)).astype(np.float32)
        index.add(x, doc_ids=DOC, psg_ids=PSG)
        assert index._get_internal_dim() == 4 and index.dim == 16
        assert all(v.shape == (16,) for v, _, _ in index)
        index.mode = api.Mode.MAXP
        codes, ids = index._get_vectors(sorted(set(DOC)))
        assert codes.shape == (5, 4) and codes.dtype == np.uint8
        dec = pq.decode(pq.encode(x)).astype(np.float64)
        for mode, red in ((api.Mode.MAXP, np.max), (api.Mode.AVEP, np.mean), (api.Mode.FIRSTP, lambda a: a[0])):
            index.mode = mode
            out = index(api.Ranking.from_run(DOC_RUN, queries=QUERIES))
            for qi, q in enumerate(("q1", "q2")):
                s = dec @ qv[qi].astype(np.float64)
                want = {"d0": red(s[:2]), "d1": s[2], "d2": s[3], "d3": s[4]}
                tol = 1e-5 * np.linalg.norm(qv[qi]) * np.linalg.norm(dec, axis=1).max()
                for d, w in want.items():
                    assert abs(out[q][d] - w) <= 1e-5 * abs(w) + tol


def test_mid_size_random_vs_oracle(api):
    """A few thousand pairs through the full API against the numpy oracle, bit for bit."""
    rng = np.random.default_rng(12)
    n_docs, dim, nq, C = 500, 768, 12, 300
    cnt = rng.integers(1, 8, n_docs)
    off = np.concatenate([[0], np.cumsum(cnt)])
    vec = rng.standard_normal((off[-1], dim)).astype(np.float32)
    doc_ids = [f"D{d}" for d in np.repeat(np.arange(n_docs), cnt)]
    qv = rng.standard_normal((nq, dim)).astype(np.float32)
    index = api.new(query_encoder=api.TableEncoder({f"t{i}": qv[i] for i in range(nq)}), init_size=len(vec))
    index.add(vec, doc_ids=doc_ids)
    run = {f"q{i}": {f"D{d}": float(np.float32(rng.uniform(0, 20))) for d in rng.choice(n_docs, C, replace=False)}
           for i in range(nq)}
    first = api.Ranking.from_run(run, queries={f"q{i}": f"t{i}" for i in range(nq)})
    for mode, m in ((api.Mode.MAXP, fo.MODE_MAXP), (api.Mode.AVEP, fo.MODE_AVEP)):
        index.mode = mode
        out = index(first)
        df = first._df
        pair_q = df["q_id"].str[1:].astype(int).to_numpy()
        pair_d = df["id"].str[1:].astype(int).to_numpy()
        want = fo.score_pairs(vec, off, np.arange(off[-1]), pair_q, pair_d, qv, m)
        got = out._df.set_index(["q_id", "id"])["score"].reindex(list(zip(df["q_id"], df["id"]))).to_numpy()
        assert (got.view(np.uint32) == want.view(np.uint32)).all()
        inter = first.interpolate(out, 0.1).cut(50)
        s = fo.interpolate_f32(df["score"].to_numpy(), want, 0.1)
        ts, tp = fo.topk_per_query(np.arange(nq + 1) * C, s, 50)
        assert (inter._df["score"].to_numpy().view(np.uint32) == ts.ravel().view(np.uint32)).all()
        assert index.rerank(first, 0.1, cutoff=50) == inter


def test_candidates_are_cached_on_the_ranking_and_follow_the_index(api):
    """A ranking's ids are resolved against an index once (per distinct id), the integer candidates
    stay on the ranking's columns, and a later `add` (new documents, a document extended by more
    rows) is seen: the cache is keyed by the index's contents version.  Results stay those of the
    reference semantics (index/util.py:29-41): IndexError for an id the index does not hold yet."""
    rng = np.random.default_rng(3)
    dim = 384
    vec = rng.standard_normal((60, dim)).astype(np.float32)
    qv = {f"query {i}": rng.standard_normal(dim).astype(np.float32) for i in range(4)}
    index = api.InMemoryIndex(api.LambdaEncoder(lambda t: qv[t]), mode=api.Mode.MAXP)
    index.add(vec[:40], doc_ids=[f"d{i // 2}" for i in range(40)])
    frame = pd.DataFrame({"q_id": np.repeat([f"q{i}" for i in range(4)], 10),
                          "id": [f"d{(7 * i) % 20}" for i in range(40)], "score": rng.uniform(0, 5, 40).astype(np.float32)})
    queries = {f"q{i}": f"query {i}" for i in range(4)}
    r = api.Ranking(frame, queries=queries)
    first = index.rerank(r, 0.3, 5)
    cols = r._cols
    assert cols is not None and len(cols._cand) == 1
    cached = next(iter(cols._cand.values()))
    second = index.rerank(r, 0.3, 5)
    assert next(iter(cols._cand.values())) is cached  # no second resolve
    pd.testing.assert_frame_equal(first._df, second._df)
    pd.testing.assert_frame_equal(first._df, r.interpolate(index(r), 0.3).cut(5)._df)
    # other mode, other cache entry; back again
    index.mode = api.Mode.FIRSTP
    fp = index.rerank(r, 0.3, 5)
    index.mode = api.Mode.MAXP
    assert not fp._df.equals(first._df)
    # a ranking naming a document the index does not hold yet
    frame2 = pd.concat([frame, pd.DataFrame({"q_id": ["q0"], "id": ["d25"], "score": np.float32([9.0])})])
    r2 = api.Ranking(frame2, queries=queries)
    with pytest.raises(IndexError, match="ID d25 not found in the index."):
        index.rerank(r2, 0.3, 5)
    index.add(vec[40:], doc_ids=[f"d{20 + i // 2}" for i in range(18)] + ["d0", "d0"])  # d25 arrives, d0 grows
    third = index.rerank(r2, 0.3, 5)
    assert "d25" in third._df["id"].tolist()
    grown = index.rerank(r, 0.3, 5)  # d0 has four rows now: its MAXP score may only go up
    old, new = index(r)["q0"], None
    ref_index = api.InMemoryIndex(api.LambdaEncoder(lambda t: qv[t]), mode=api.Mode.MAXP)
    ref_index.add(vec, doc_ids=[f"d{i // 2}" for i in range(40)] + [f"d{20 + i // 2}" for i in range(18)] + ["d0", "d0"])
    pd.testing.assert_frame_equal(grown._df, ref_index.rerank(api.Ranking(frame, queries=queries), 0.3, 5)._df)
    assert old == ref_index(api.Ranking(frame, queries=queries))["q0"] and new is None


def test_early_stopping_only_fails_for_ids_it_scores(api):
    """The reference resolves ids depth by depth (index/base.py:373 -> index/util.py:38-39): an id
    the index does not hold raises IndexError only when its row is reached.  Here all ids are coded
    up front, so unknown ones are checked against the scored prefixes afterwards."""
    rng = np.random.default_rng(11)
    dim = 384
    vec = rng.standard_normal((200, dim)).astype(np.float32)
    index = api.InMemoryIndex(api.LambdaEncoder(lambda t: np.ones(dim, np.float32)), mode=api.Mode.MAXP)
    index.add(vec, doc_ids=[f"d{i}" for i in range(200)])
    # first-stage scores fall fast: with alpha = 1 a query stops after the first interval
    ids = [f"d{i}" for i in range(60)]
    frame = pd.DataFrame({"q_id": ["q1"] * 60, "id": ids, "score": np.linspace(100, 1, 60).astype(np.float32)})
    late = frame.copy()
    late.loc[45, "id"] = "unknown-late"
    early = frame.copy()
    early.loc[3, "id"] = "unknown-early"
    kw = dict(early_stopping=5, early_stopping_alpha=1.0, early_stopping_depths=[10, 30, 60])
    want = index(api.Ranking(frame, queries=QUERIES), **kw)
    got = index(api.Ranking(late, queries=QUERIES), **kw)  # the walk stops at depth 10: row 45 is never looked up
    assert len(want._df) == 10
    pd.testing.assert_frame_equal(got._df, want._df)
    with pytest.raises(IndexError, match="ID unknown-early not found in the index."):
        index(api.Ranking(early, queries=QUERIES), **kw)
    with pytest.raises(IndexError, match="ID unknown-late not found in the index."):  # alpha 0: every depth is walked
        index(api.Ranking(late, queries=QUERIES), early_stopping=5, early_stopping_alpha=0.0,
              early_stopping_depths=[10, 30, 60])
    with pytest.raises(IndexError, match="ID unknown-late not found in the index."):  # no early stopping: all rows
        index(api.Ranking(late, queries=QUERIES))


def test_one_very_long_list_among_many_short_ones(api, monkeypatch):
    """Memory follows the number of pairs, not queries x longest list: with skewed list lengths
    the order comes from a host radix sort over the device's semantic scores instead of dense
    [nq, widest] list matrices.  Same frames as the dense route."""
    from fast_forward.index.base import Index

    rng = np.random.default_rng(12)
    dim, n_docs, nq = 64, 30_000, 1500
    vec = rng.standard_normal((n_docs, dim)).astype(np.float32)
    qv = rng.standard_normal((nq, dim)).astype(np.float32)
    index = api.InMemoryIndex(api.TableEncoder({f"text {i}": qv[i] for i in range(nq)}), mode=api.Mode.MAXP)
    index.add(vec, doc_ids=[f"d{i}" for i in range(n_docs)])
    q_ids = np.concatenate([np.repeat("q0", n_docs), np.repeat([f"q{i}" for i in range(1, nq)], 3)])
    ids = np.concatenate([np.arange(n_docs), rng.integers(0, n_docs - 2, nq - 1).repeat(3) + np.tile([0, 1, 2], nq - 1)])
    frame = pd.DataFrame({"q_id": q_ids, "id": [f"d{i}" for i in ids], "score": (rng.integers(0, 50, len(ids)) * 0.25).astype(np.float32)})
    r = api.Ranking(frame, queries={f"q{i}": f"text {i}" for i in range(nq)})
    assert Index._lists_are_skewed(nq, n_docs, len(frame))
    skewed = (index(r), index.rerank(r, 0.3, 10), index.rerank(r, 0.3), r.interpolate(index(r), 0.3).cut(7))
    monkeypatch.setattr(Index, "_lists_are_skewed", staticmethod(lambda *a: False))
    dense = (index(r), index.rerank(r, 0.3, 10), index.rerank(r, 0.3), r.interpolate(index(r), 0.3).cut(7))
    for a, b in zip(skewed, dense):
        pd.testing.assert_frame_equal(a._df, b._df)

