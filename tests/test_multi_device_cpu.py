"""The host logic of the multi-device stores (`devices=[...]`, `shard="query" | "doc"`) without a
GPU: `_ffx.DeviceIndex` is replaced by tests/fake_device.py, which keeps rows in host memory and
scores with the oracle.  Checks what the host adds: placement of documents on devices, candidate
encoding `device * stride + local`, the split of queries over replicas, the merge of per-shard
lists — the frames must equal those of one (fake) device."""

import numpy as np
import pandas as pd
import pytest

from fake_device import FakeDeviceIndex


@pytest.fixture()
def api(monkeypatch):
    import fast_forward
    from fast_forward import _ffx
    from fast_forward.encoder import LambdaEncoder
    from fast_forward.index import InMemoryIndex, Mode

    monkeypatch.setattr(_ffx, "DeviceIndex", FakeDeviceIndex)
    monkeypatch.setattr(_ffx, "_DEVICE", False)  # no pinned allocations

    class Api:
        pass

    a = Api()
    a.Ranking, a.InMemoryIndex, a.Mode, a.LambdaEncoder = fast_forward.Ranking, InMemoryIndex, Mode, LambdaEncoder
    return a


def build(api, rng, devices, shard, dim=48, n_docs=120):
    cnt = rng.integers(1, 6, n_docs)
    doc_of_row = np.repeat(np.arange(n_docs), cnt)
    rng.shuffle(doc_of_row)
    vec = rng.standard_normal((len(doc_of_row), dim)).astype(np.float32)
    doc_ids = [f"d{d}" for d in doc_of_row]
    psg_ids = [None if i % 11 == 0 else f"p{i}" for i in range(len(vec))]
    for i in range(3, len(vec), 13):
        if psg_ids[i] is not None:
            doc_ids[i] = None
    return vec, doc_ids, psg_ids


def fill(index, vec, doc_ids, psg_ids, pieces=5):
    step = -(-len(vec) // pieces)
    for lo in range(0, len(vec), step):
        index.add(vec[lo:lo + step], doc_ids=doc_ids[lo:lo + step], psg_ids=psg_ids[lo:lo + step])
    return index


@pytest.mark.parametrize("shard,devices", [("query", [0, 1]), ("query", [0, 1, 2, 3]), ("doc", [0, 1]), ("doc", [0, 1, 2])])
def test_fake_devices_give_the_frames_of_one(api, shard, devices):
    rng = np.random.default_rng(len(devices) + (7 if shard == "doc" else 0))
    dim, nq = 48, 9
    vec, doc_ids, psg_ids = build(api, rng, devices, shard, dim)
    qv = {f"query {i}": rng.standard_normal(dim).astype(np.float32) for i in range(nq)}
    queries = {f"q{i}": f"query {i}" for i in range(nq)}
    enc = api.LambdaEncoder(lambda t: qv[t])
    one = fill(api.InMemoryIndex(enc, init_size=16, alloc_size=16), vec, doc_ids, psg_ids)
    many = fill(api.InMemoryIndex(enc, init_size=16, alloc_size=16, devices=devices, shard=shard), vec, doc_ids, psg_ids)
    assert len(many) == len(one) and many.doc_ids == one.doc_ids and many.psg_ids == one.psg_ids
    store = many._store
    if shard == "doc":
        assert store.loads.sum() == len(vec) and (store.loads > 0).all()
        assert store.loads.max() - store.loads.min() <= 0.35 * len(vec)
        # a document lives on one device
        enc_of_row = store._row_enc() // store.stride
        row_doc = store._row_doc()
        for d in np.unique(row_doc[row_doc >= 0]):
            assert len(set(enc_of_row[row_doc == d])) == 1
    else:
        assert all(len(r) == len(vec) for r in store.all_devices())
    docs, psgs = sorted(one.doc_ids), sorted(one.psg_ids)
    for mode in api.Mode:
        one.mode = many.mode = mode
        ids = psgs if mode == api.Mode.PASSAGE else docs
        rows = []
        for q in range(nq):
            pick = rng.choice(len(ids), int(rng.integers(1, 60)), replace=False)
            rows += [(f"q{q}", ids[i], np.float32(rng.integers(0, 8) * 0.5)) for i in pick]
        r = api.Ranking(pd.DataFrame(rows, columns=["q_id", "id", "score"]), queries=queries)
        for a, b in ((many(r), one(r)), (many(r, batch_size=2), one(r)), (many.rerank(r, 0.25, 7), one.rerank(r, 0.25, 7)),
                     (many.rerank(r, 1.0, 3), one.rerank(r, 1.0, 3)),
                     (r.interpolate(many(r), 0.5).cut(4), r.interpolate(one(r), 0.5).cut(4)),
                     (many(r, early_stopping=3, early_stopping_alpha=0.4, early_stopping_depths=[5, 20, 60]),
                      one(r, early_stopping=3, early_stopping_alpha=0.4, early_stopping_depths=[5, 20, 60]))):
            pd.testing.assert_frame_equal(a._df, b._df)
        got_v, got_ids = many._get_vectors(ids[::9])
        want_v, want_ids = one._get_vectors(ids[::9])
        assert got_ids == want_ids and (got_v == want_v).all()
    assert [(v.tobytes(), d, p) for v, d, p in many] == [(v.tobytes(), d, p) for v, d, p in one]
    with pytest.raises(IndexError, match="ID nope not found in the index."):
        many(api.Ranking(pd.DataFrame({"q_id": ["q0"], "id": ["nope"], "score": [1.0]}), queries=queries))
    if shard == "query":  # every replica took part in a many-query call
        assert all(d.calls > 0 for d in store.all_devices())


def test_store_factory_arguments():
    from fast_forward.index._store import DocShardedStore, ReplicatedStore, RowStore, make_store

    assert type(make_store(0, None, "query")) is RowStore and type(make_store(0, [1], "doc")) is RowStore
    assert type(make_store(0, [0, 1], "query")) is ReplicatedStore and type(make_store(0, [0, 1], "doc")) is DocShardedStore
    with pytest.raises(ValueError):
        make_store(0, [0, 1], "rows")
    with pytest.raises(ValueError):
        make_store(0, [1, 1], "query")


def test_skewed_lists_and_lazy_early_stopping_ids_on_the_fake_device(api, monkeypatch):
    """Host logic that does not depend on the GPU: the O(n) route for skewed list lengths gives the
    dense route's frames, and early stopping raises only for unknown ids it reaches."""
    from fast_forward.index.base import Index

    rng = np.random.default_rng(31)
    dim, n_docs, nq = 16, 400, 30
    vec = rng.standard_normal((n_docs, dim)).astype(np.float32)
    qv = {f"query {i}": rng.standard_normal(dim).astype(np.float32) for i in range(nq)}
    queries = {f"q{i}": f"query {i}" for i in range(nq)}
    index = api.InMemoryIndex(api.LambdaEncoder(lambda t: qv[t]))
    index.add(vec, doc_ids=[f"d{i}" for i in range(n_docs)])
    rows = [("q0", f"d{i}", np.float32(rng.integers(0, 9) * 0.5)) for i in range(n_docs)]
    for q in range(1, nq):
        rows += [(f"q{q}", f"d{i}", np.float32(rng.integers(0, 9) * 0.5)) for i in rng.choice(n_docs, 3, replace=False)]
    r = api.Ranking(pd.DataFrame(rows, columns=["q_id", "id", "score"]), queries=queries)
    monkeypatch.setattr(Index, "_lists_are_skewed", staticmethod(lambda *a: True))
    skewed = (index(r), index.rerank(r, 0.3, 5), index.rerank(r, 0.3))
    monkeypatch.setattr(Index, "_lists_are_skewed", staticmethod(lambda *a: False))
    dense = (index(r), index.rerank(r, 0.3, 5), index.rerank(r, 0.3))
    for a, b in zip(skewed, dense):
        pd.testing.assert_frame_equal(a._df, b._df)

    frame = pd.DataFrame({"q_id": ["q1"] * 40, "id": [f"d{i}" for i in range(40)], "score": np.linspace(90, 1, 40).astype(np.float32)})
    late = frame.copy()
    late.loc[33, "id"] = "unknown-late"
    kw = dict(early_stopping=4, early_stopping_alpha=1.0, early_stopping_depths=[8, 20, 40])
    pd.testing.assert_frame_equal(index(api.Ranking(late, queries=queries), **kw)._df,
                                  index(api.Ranking(frame, queries=queries), **kw)._df)
    with pytest.raises(IndexError, match="ID unknown-late not found in the index."):
        index(api.Ranking(late, queries=queries), early_stopping=4, early_stopping_alpha=0.0, early_stopping_depths=[8, 20, 40])


def test_on_disk_round_trip_and_to_memory_on_a_fake_device(api, tmp_path):
    """`OnDiskIndex` -> file -> `OnDiskIndex.load` (id columns adopted on their own thread beside the
    row stream) -> `to_memory()` (rows copied device to device, dictionaries cloned) with the fake
    device: same contents, same frames, and the copies grow apart."""
    from pathlib import Path

    from fast_forward.index import OnDiskIndex

    rng = np.random.default_rng(3)
    dim, nq = 16, 6
    vec, doc_ids, psg_ids = build(api, rng, [0], "query", dim, n_docs=60)
    qv = {f"query {i}": rng.standard_normal(dim).astype(np.float32) for i in range(nq)}
    queries = {f"q{i}": f"query {i}" for i in range(nq)}
    enc = api.LambdaEncoder(lambda t: qv[t])
    path = Path(tmp_path) / "index.h5"
    disk = fill(OnDiskIndex(path, enc, init_size=32, chunk_size=32), vec, doc_ids, psg_ids, pieces=3)
    loaded = OnDiskIndex.load(path, enc)
    assert len(loaded) == len(disk) == len(vec)
    assert loaded.doc_ids == disk.doc_ids and loaded.psg_ids == disk.psg_ids
    mem = loaded.to_memory()
    assert isinstance(mem, api.InMemoryIndex) and len(mem) == len(vec)
    assert mem._store.dev is not loaded._store.dev and mem.doc_ids == disk.doc_ids and mem.psg_ids == disk.psg_ids
    docs = sorted(d for d in disk.doc_ids)
    ids = [docs[i] for i in rng.integers(0, len(docs), nq * 20)]
    frame = pd.DataFrame({"q_id": np.repeat([f"q{i}" for i in range(nq)], 20), "id": ids,
                          "score": rng.uniform(0, 10, nq * 20).astype(np.float32)}).drop_duplicates(["q_id", "id"])
    r = api.Ranking(frame, queries=queries)
    for mode in (api.Mode.MAXP, api.Mode.AVEP, api.Mode.FIRSTP):
        disk.mode = loaded.mode = mem.mode = mode
        want = disk(r)
        assert loaded(r) == want and mem(r) == want
    mem.add(vec[:2], doc_ids=["brand-new", "brand-new"], psg_ids=["px1", "px2"])
    assert len(mem) == len(vec) + 2 and len(loaded) == len(vec) and "brand-new" not in loaded.doc_ids
    with pytest.raises(RuntimeError):
        mem.add(vec[:1], psg_ids=["px1"])  # the cloned dictionary knows its own additions
