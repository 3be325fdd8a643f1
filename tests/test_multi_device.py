"""Several GPUs behind the drop-in API, driven by ONE process (`devices=[...]`, SURVEY 8e / row g):
`shard="query"` = a replica of the rows on every device, each call's queries split over them;
`shard="doc"` = the documents spread over the devices, every device scoring its part of every
query, per-device lists merged.  Every frame must be the one a single device produces — which
the rest of the suite pins against the reference.  With fewer than two GPUs the same code runs
with both "devices" on GPU 0."""

import os

import numpy as np
import pandas as pd
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    import __graft_entry__ as g

    g.build()
    import fast_forward
    from fast_forward import _ffx
    from fast_forward.encoder import LambdaEncoder
    from fast_forward.index import InMemoryIndex, Mode, OnDiskIndex
    from fast_forward.quantizer import NanoOPQ

    class Api:
        pass

    a = Api()
    a.ff, a.Ranking, a.InMemoryIndex, a.OnDiskIndex, a.Mode, a.LambdaEncoder, a.NanoOPQ = \
        fast_forward, fast_forward.Ranking, InMemoryIndex, OnDiskIndex, Mode, LambdaEncoder, NanoOPQ
    n_gpu = _ffx.device_count()
    a.devices = list(range(min(n_gpu, 8))) if n_gpu >= 2 else [0, 0, 0]  # every GPU of the box
    if n_gpu < 2:
        os.environ["FFX_ALLOW_DUPLICATE_DEVICES"] = "1"
    return a


def corpus(rng, n_docs, dim, max_psg=6):
    cnt = rng.integers(1, max_psg + 1, n_docs)
    doc_of_row = np.repeat(np.arange(n_docs), cnt)
    rng.shuffle(doc_of_row)  # documents own scattered rows and are extended by later adds
    vec = rng.standard_normal((len(doc_of_row), dim)).astype(np.float32)
    doc_ids = [f"d{d}" for d in doc_of_row]
    psg_ids = [f"p{i}" for i in range(len(vec))]
    # some rows carry only one kind of id (reference tests/test_index.py:58-69)
    for i in range(0, len(vec), 17):
        psg_ids[i] = None
    for i in range(5, len(vec), 23):
        if psg_ids[i] is not None:
            doc_ids[i] = None
    return vec, doc_ids, psg_ids


def fill(index, vec, doc_ids, psg_ids, pieces=4):
    step = -(-len(vec) // pieces)
    for lo in range(0, len(vec), step):
        index.add(vec[lo:lo + step], doc_ids=doc_ids[lo:lo + step], psg_ids=psg_ids[lo:lo + step])
    return index


def first_stage(rng, api, ids, nq, per_query, queries):
    rows = []
    for q in range(nq):
        pick = rng.choice(len(ids), min(per_query, len(ids)), replace=False)
        rows += [(f"q{q}", ids[i], np.float32(rng.integers(0, 12) * 0.5)) for i in pick]  # coarse scores: many ties
    frame = pd.DataFrame(rows, columns=["q_id", "id", "score"])
    return api.Ranking(frame, queries=queries)


def same(a, b):
    pd.testing.assert_frame_equal(a._df, b._df)


@pytest.mark.parametrize("shard", ["query", "doc"])
@pytest.mark.parametrize("dim,nq", [(768, 320), (384, 7), (100, 40)])
def test_several_devices_give_the_frames_of_one(api, shard, dim, nq):
    rng = np.random.default_rng(dim + nq)
    vec, doc_ids, psg_ids = corpus(rng, 500, dim)
    qv = {f"query {i}": rng.standard_normal(dim).astype(np.float32) for i in range(nq)}
    queries = {f"q{i}": f"query {i}" for i in range(nq)}
    enc = api.LambdaEncoder(lambda t: qv[t])
    one = fill(api.InMemoryIndex(enc, init_size=64, alloc_size=64), vec, doc_ids, psg_ids)
    many = fill(api.InMemoryIndex(enc, init_size=64, alloc_size=64, devices=api.devices, shard=shard), vec, doc_ids, psg_ids)
    assert len(many) == len(one) and many.doc_ids == one.doc_ids and many.psg_ids == one.psg_ids and many.dim == dim
    docs = sorted(one.doc_ids)
    psgs = sorted(one.psg_ids)
    for mode in api.Mode:
        one.mode = many.mode = mode
        r = first_stage(rng, api, psgs if mode == api.Mode.PASSAGE else docs, nq, 150, queries)
        same(many(r), one(r))
        same(many(r, batch_size=3), one(r))
        same(r.interpolate(many(r), 0.2), r.interpolate(one(r), 0.2))
        same(many.rerank(r, 0.2, 10), one.rerank(r, 0.2, 10))
        same(many.rerank(r, 0.999, 10), one.rerank(r, 0.999, 10))  # ties straddling the cut
        same(many.rerank(r, 0.2), r.interpolate(one(r), 0.2))
        es = dict(early_stopping=5, early_stopping_alpha=0.3, early_stopping_depths=[10, 50, 150])
        same(many(r, **es), one(r, **es))
        some = (psgs if mode == api.Mode.PASSAGE else docs)[::37]
        got_v, got_ids = many._get_vectors(some)
        want_v, want_ids = one._get_vectors(some)
        assert got_ids == want_ids and (got_v == want_v).all()
    got = [(v.tobytes(), d, p) for v, d, p in many]
    want = [(v.tobytes(), d, p) for v, d, p in one]
    assert got == want
    with pytest.raises(IndexError, match="ID nope not found in the index."):
        many(api.Ranking(pd.DataFrame({"q_id": ["q0"], "id": ["nope"], "score": [1.0]}), queries=queries))
    if shard == "doc":
        loads = many._store.loads
        assert loads.sum() == len(vec) and (loads > 0).sum() == len(set(api.devices)) or len(set(api.devices)) == 1
        assert loads.max() - loads.min() <= 0.2 * len(vec)  # rows spread evenly


@pytest.mark.parametrize("shard", ["query", "doc"])
def test_quantized_and_on_disk_indexes_on_several_devices(api, shard, tmp_path):
    rng = np.random.default_rng(9)
    dim, nq = 64, 200
    vec, doc_ids, psg_ids = corpus(rng, 300, dim)
    quant = api.NanoOPQ(8, 32)
    quant.fit(vec)
    qv = {f"query {i}": rng.standard_normal(dim).astype(np.float32) for i in range(nq)}
    queries = {f"q{i}": f"query {i}" for i in range(nq)}
    enc = api.LambdaEncoder(lambda t: qv[t])
    one = fill(api.InMemoryIndex(enc, quantizer=quant), vec, doc_ids, psg_ids)
    many = fill(api.InMemoryIndex(enc, quantizer=quant, devices=api.devices, shard=shard), vec, doc_ids, psg_ids)
    r = first_stage(rng, api, sorted(one.doc_ids), nq, 100, queries)
    for mode in (api.Mode.MAXP, api.Mode.AVEP):
        one.mode = many.mode = mode
        same(many.rerank(r, 0.3, 20), one.rerank(r, 0.3, 20))
        same(many(r), one(r))

    path = tmp_path / "index.h5"
    ids8 = lambda xs: [None if x is None else x[:8] for x in xs]  # noqa: E731
    disk = fill(api.OnDiskIndex(path, enc, devices=api.devices, shard=shard), vec, ids8(doc_ids), ids8(psg_ids))
    plain = fill(api.InMemoryIndex(enc), vec, ids8(doc_ids), ids8(psg_ids))
    r = first_stage(rng, api, sorted(plain.doc_ids), nq, 100, queries)
    same(disk.rerank(r, 0.3, 20), plain.rerank(r, 0.3, 20))
    loaded = api.OnDiskIndex.load(path, enc, devices=api.devices, shard=shard)
    assert len(loaded) == len(plain) and loaded.doc_ids == plain.doc_ids
    same(loaded.rerank(r, 0.3, 20), plain.rerank(r, 0.3, 20))
    same(loaded(r), plain(r))
    same(loaded.to_memory().rerank(r, 0.3, 20), plain.rerank(r, 0.3, 20))
