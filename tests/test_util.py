"""`fast_forward.util` — Indexer, sequential coalescing, evaluation frame: the reference's
tests (tests/test_indexer.py:21-81, tests/test_util.py:10-16, tests/test_index.py:351-374)
re-stated for the HBM-resident index, plus a differential test of the coalescing against the
unmodified reference function on random documents."""

import importlib.util
import os

import numpy as np
import pytest

V = np.tril(np.ones((5, 5), dtype=np.float32))
DOC = ["d0", "d0", "d1", "d2", "d3"]
REF_UTIL = os.path.join(os.path.dirname(__file__), "..", "baseline", "_ref", "fast_forward", "util", "__init__.py")


def test_ir_measures_frame_and_cos_dist():
    from fast_forward import Ranking
    from fast_forward.util import cos_dist, to_ir_measures

    r = Ranking.from_run({"q1": {"d0": 1.0, "d1": 2.0}, "q2": {"d0": 3.0}}, queries={"q1": "a", "q2": "b"})
    df = to_ir_measures(r)
    assert set(df.columns) == {"query_id", "doc_id", "score"}
    assert df["query_id"].equals(r._df["q_id"]) and df["doc_id"].equals(r._df["id"]) and df["score"].equals(r._df["score"])
    assert cos_dist(np.array([1.0, 0.0]), np.array([0.0, 2.0])) == 1.0
    assert abs(cos_dist(np.array([1.0, 1.0]), np.array([2.0, 2.0]))) < 1e-12


def test_coalescing_rule_on_one_document():
    from fast_forward.util import _coalesced, cos_dist

    merged = list(_coalesced(V[:2], 0.3, cos_dist))  # cos_dist(v1, v0) = 0.29 < 0.3: one group
    assert len(merged) == 1 and np.array_equal(merged[0], np.average(V[:2], axis=0))
    apart = list(_coalesced(V[:2], 0.2, cos_dist))
    assert len(apart) == 2 and np.array_equal(apart[0], V[0]) and np.array_equal(apart[1], V[1])
    assert list(_coalesced(V[:0], 0.3, cos_dist)) == []


@pytest.fixture(scope="module")
def api():
    import __graft_entry__ as g

    g.build()
    from fast_forward import _ffx

    if _ffx.device_count() < 1:
        pytest.skip("needs a CUDA device")
    import fast_forward.util as util
    from fast_forward.encoder import LambdaEncoder
    from fast_forward.index import InMemoryIndex, Mode
    from fast_forward.quantizer import NanoPQ

    class Api:
        pass

    a = Api()
    a.util, a.LambdaEncoder, a.InMemoryIndex, a.Mode, a.NanoPQ = util, LambdaEncoder, InMemoryIndex, Mode, NanoPQ
    a.ffx = _ffx
    return a


@pytest.mark.gpu
def test_indexer_from_dicts_and_from_index(api):
    target = api.InMemoryIndex()
    indexer = api.util.Indexer(target, api.LambdaEncoder(lambda _: np.zeros(16)), encoder_batch_size=2, batch_size=4)
    dicts = [{"text": "123", "doc_id": "d1", "psg_id": "d1_p1"}, {"text": "234", "doc_id": "d1", "psg_id": "d1_p2"},
             {"text": "456", "doc_id": "d1", "psg_id": "d1_p3"}, {"text": "567", "doc_id": "d2", "psg_id": "d2_p1"},
             {"text": "678", "doc_id": "d3", "psg_id": "d3_p1"}, {"text": "890", "doc_id": "d4"},
             {"text": "901", "psg_id": "d5_p1"}]
    indexer.from_dicts(dicts)
    assert len(target) == 7 and target.doc_ids == {"d1", "d2", "d3", "d4"}
    assert target.psg_ids == {"d1_p1", "d1_p2", "d1_p3", "d2_p1", "d3_p1", "d5_p1"}
    with pytest.raises(RuntimeError):
        api.util.Indexer(target, encoder=None).from_dicts(dicts)

    source = api.InMemoryIndex()
    source.add(np.arange(256, dtype=np.float32).reshape(16, 16), doc_ids=[f"d{i}" for i in range(16)])
    copy = api.InMemoryIndex()
    api.util.Indexer(copy, batch_size=5).from_index(source)
    assert copy.doc_ids == source.doc_ids and len(copy) == 16
    got, _ = copy._get_vectors(["d3"])
    assert np.array_equal(got[0], np.arange(48, 64, dtype=np.float32))


@pytest.mark.gpu
def test_indexer_fits_and_attaches_a_quantizer(api):
    rng = np.random.default_rng(0)
    for fit_batches in (1, 2):
        target = api.InMemoryIndex()
        indexer = api.util.Indexer(target, encoder=api.LambdaEncoder(lambda _: rng.normal(size=32).astype(np.float32)),
                                   quantizer=api.NanoPQ(4, 8), batch_size=16, quantizer_fit_batches=fit_batches)
        indexer.from_dicts([{"text": f"text_{i}", "doc_id": f"d{i}"} for i in range(64)])
        assert target.quantizer._trained and len(target) == 64
    fitted = api.NanoPQ(4, 8)
    fitted.fit(rng.normal(size=(64, 64)).astype(np.float32))
    with pytest.raises(ValueError):
        api.util.Indexer(api.InMemoryIndex(), quantizer=fitted)
    used = api.InMemoryIndex()
    used.add(np.zeros((8, 16), np.float32), doc_ids=[f"d{i}" for i in range(8)])
    with pytest.raises(ValueError):
        api.util.Indexer(used, quantizer=api.NanoPQ(4, 8))


@pytest.mark.gpu
def test_coalesced_index(api):
    source = api.InMemoryIndex(mode=api.Mode.MAXP)
    source.add(V, doc_ids=DOC)
    merged = api.InMemoryIndex(mode=api.Mode.MAXP)
    api.util.create_coalesced_index(source, merged, 0.3)  # d0's two vectors are averaged
    assert merged.doc_ids == source.doc_ids
    d0, _ = merged._get_vectors(["d0"])
    assert len(d0) == 1 and np.array_equal(d0[0], np.average(V[:2], axis=0))
    same = api.InMemoryIndex(mode=api.Mode.MAXP)
    api.util.create_coalesced_index(source, same, 0.2, batch_size=2)  # nothing merges
    for doc_id in source.doc_ids:
        a, _ = source._get_vectors([doc_id])
        b, _ = same._get_vectors([doc_id])
        assert np.array_equal(a, b)
    with pytest.raises(ValueError):
        api.util.create_coalesced_index(source, merged, 0.3)


@pytest.mark.gpu
def test_coalescing_equals_the_reference_function(api):
    """Random documents of 1..12 passages drawn around a few centres, so that groups really
    form; the reference function runs on stand-in index objects (plain dicts of numpy rows)."""
    if not os.path.exists(REF_UTIL):
        pytest.skip("reference package not installed")
    spec = importlib.util.spec_from_file_location("_reference_util", REF_UTIL,
                                                  submodule_search_locations=[os.path.dirname(REF_UTIL)])
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    rng = np.random.default_rng(3)
    centres = rng.standard_normal((6, 64)).astype(np.float32)
    vectors, doc_ids = [], []
    for d in range(300):
        for _ in range(int(rng.integers(1, 13))):
            vectors.append(centres[rng.integers(0, 2) + 2 * (d % 3)] + 0.35 * rng.standard_normal(64).astype(np.float32))
            doc_ids.append(f"doc{d}")
    vectors = np.stack(vectors)

    class Plain:  # what the reference function needs of an index
        def __init__(self):
            self.rows, self.ids = [], []

        doc_ids = property(lambda self: set(self.ids))

        def __len__(self):
            return len(self.rows)

        def add(self, v, doc_ids):
            self.rows.extend(v)
            self.ids.extend(doc_ids)

        def _get_vectors(self, ids):
            keep = [i for i, d in enumerate(self.ids) if d in ids]
            return np.stack([self.rows[i] for i in keep]), [self.ids[i] for i in keep]

    plain_source, plain_target = Plain(), Plain()
    plain_source.add(vectors, doc_ids)
    for delta in (0.25, 0.6):
        plain_target = Plain()
        ref.create_coalesced_index(plain_source, plain_target, delta)
        source, target = api.InMemoryIndex(mode=api.Mode.MAXP), api.InMemoryIndex(mode=api.Mode.MAXP)
        source.add(vectors, doc_ids=doc_ids)
        api.util.create_coalesced_index(source, target, delta, batch_size=257)
        assert len(target) == len(plain_target) < len(vectors)
        for d in range(0, 300, 7):
            got, _ = target._get_vectors([f"doc{d}"])
            want, _ = plain_target._get_vectors([f"doc{d}"])
            assert np.array_equal(got, want)


@pytest.mark.gpu
@pytest.mark.parametrize("dim", [768, 100, 64])
def test_device_coalescing_equals_the_host_loop(api, dim):
    """The default distance runs on the device (ffx_index_coalesce: one warp per document); any
    other callable — here the same function behind a lambda — takes the host loop.  Documents
    own scattered rows (extended by later adds); lane-major (768, 64) and tree-as-data (100)
    store layouts.  Same groups, bit-identical means."""
    rng = np.random.default_rng(dim)
    centres = rng.standard_normal((8, dim)).astype(np.float32)
    owner = np.repeat(np.arange(400), rng.integers(1, 15, 400))
    rng.shuffle(owner)
    vectors = (centres[(owner % 4) * 2 + rng.integers(0, 2, len(owner))] +
               0.3 * rng.standard_normal((len(owner), dim))).astype(np.float32)
    doc_ids = [f"doc{d}" for d in owner]
    source = api.InMemoryIndex(mode=api.Mode.MAXP)
    half = len(vectors) // 2
    source.add(vectors[:half], doc_ids=doc_ids[:half])
    source.add(vectors[half:], doc_ids=doc_ids[half:])
    for delta in (0.2, 0.5, 5.0):
        on_device, on_host = api.InMemoryIndex(mode=api.Mode.MAXP), api.InMemoryIndex(mode=api.Mode.MAXP)
        before = api.ffx.launch_count()
        api.util.create_coalesced_index(source, on_device, delta, batch_size=300)
        assert api.ffx.launch_count() > before
        api.util.create_coalesced_index(source, on_host, delta, distance_function=lambda a, b: api.util.cos_dist(a, b))
        assert len(on_device) == len(on_host) <= len(vectors) and on_device.doc_ids == on_host.doc_ids
        if delta == 5.0:  # nothing is farther than that: one mean per document
            assert len(on_device) == 400
        for d in range(0, 400, 9):
            got, _ = on_device._get_vectors([f"doc{d}"])
            want, _ = on_host._get_vectors([f"doc{d}"])
            assert np.array_equal(got, want), (delta, d)


@pytest.mark.gpu
def test_pyterrier_transformers_on_a_stand_in_module(api, monkeypatch):
    """`pyterrier` is not installed here: a stand-in with the two things the glue touches
    (`Transformer`, `model.add_ranks`) checks columns and values of FFScore -> FFInterpolate
    (reference: util/pyterrier.py:14-87)."""
    import sys
    import types

    import pandas as pd

    pt = types.ModuleType("pyterrier")
    pt.Transformer = type("Transformer", (), {})

    def add_ranks(df, single_query=False):
        out = df.copy()
        out["rank"] = out.groupby("qid")["score"].rank(ascending=False, method="first").astype(int) - 1
        return out

    pt.model = types.SimpleNamespace(add_ranks=add_ranks)
    monkeypatch.setitem(sys.modules, "pyterrier", pt)
    sys.modules.pop("fast_forward.util.pyterrier", None)
    from fast_forward.util.pyterrier import FFInterpolate, FFScore

    index = api.InMemoryIndex(api.LambdaEncoder(lambda _: np.ones(5, np.float32)), mode=api.Mode.MAXP)
    index.add(V, doc_ids=DOC)
    inp = pd.DataFrame({"qid": ["q1"] * 4 + ["q2"] * 4, "docno": ["d0", "d1", "d2", "d3"] * 2,
                        "score": [100.0, 2, 3, 200, 400, 5, 6, 800], "query": ["one"] * 4 + ["two"] * 4})
    scorer = FFScore(index)
    scored = scorer.transform(inp)
    assert list(scored.columns) == ["qid", "docno", "score", "query", "score_0", "rank"]
    by_pair = scored.set_index(["qid", "docno"])
    assert by_pair.loc[("q1", "d0"), "score"] == 2.0 and by_pair.loc[("q2", "d3"), "score"] == 5.0  # MAXP of ones
    assert by_pair.loc[("q1", "d3"), "score_0"] == 200.0 and len(scored) == 8
    mixed = FFInterpolate(0.5).transform(scored).set_index(["qid", "docno"])
    assert mixed.loc[("q1", "d0"), "score"] == 51.0 and mixed.loc[("q2", "d3"), "score"] == 402.5
    assert mixed.loc[("q2", "d3"), "rank"] == 0 and "score_0" not in mixed.columns
    assert repr(scorer) == f"FFScore({id(index)}, {id(index._query_encoder)})"
