"""`OnDiskIndex` (HDF5 layout of the reference's index/disk.py) — modelled on the reference's
TestOnDiskIndex (tests/test_index.py:444-658): load, to_memory, quantizer persistence, id
length limit.  Uses the real h5py when installed, else tests/fake_h5py.py (h5py is absent
from the build image and the GPU box, see SURVEY 8c "HDF5 path")."""

import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

try:
    import h5py  # noqa: F401
except ImportError:
    import fake_h5py

    sys.modules["h5py"] = fake_h5py

@pytest.fixture(params=["h5py", "native-writer"], autouse=True)
def writer(request, monkeypatch):
    """Every test runs twice: writing through h5py (the real package, or tests/fake_h5py.py
    persisting HDF5 bytes) and through the package's own writer (`fast_forward._h5_write`, what
    `OnDiskIndex` uses when h5py is not installed).  Reading is always the native reader."""
    if request.param == "native-writer":
        import fast_forward.index.disk as disk

        monkeypatch.setattr(disk, "_h5py", lambda: None)
    return request.param


DOC = ["d0", "d0", "d1", "d2", "d3"]
PSG = ["p0", "p1", "p2", "p3", "p4"]
V = np.tril(np.ones((5, 5), dtype=np.float32))
QUERIES = {"q1": "query 1", "q2": "query 2"}
DOC_RUN = {"q1": {"d0": 100, "d1": 2, "d2": 3, "d3": 200}, "q2": {"d0": 400, "d1": 5, "d2": 6, "d3": 800}}


@pytest.fixture(scope="module")
def api():
    import __graft_entry__ as g

    g.build()
    import fast_forward
    from fast_forward.encoder import LambdaEncoder
    from fast_forward.index import InMemoryIndex, Mode, OnDiskIndex
    from fast_forward.quantizer import NanoPQ

    class Api:
        pass

    a = Api()
    a.Ranking, a.OnDiskIndex, a.InMemoryIndex, a.Mode, a.NanoPQ = \
        fast_forward.Ranking, OnDiskIndex, InMemoryIndex, Mode, NanoPQ
    a.ones = LambdaEncoder(lambda _: np.ones(5))
    return a


def test_create_score_reload(api, tmp_path, golden_kat):
    path = tmp_path / "index.h5"
    index = api.OnDiskIndex(path, api.ones, init_size=2, chunk_size=2)
    with pytest.raises(ValueError):
        api.OnDiskIndex(path)  # exists, overwrite=False
    index.add(V[:3], doc_ids=DOC[:3], psg_ids=PSG[:3])
    index.add(V[3:], doc_ids=DOC[3:], psg_ids=PSG[3:])
    assert len(index) == 5 and index.dim == 5 and index.doc_ids == set(DOC)
    rank = api.Ranking.from_run(DOC_RUN, queries=QUERIES)
    want = golden_kat["full/MAXP"]
    assert index(rank)._df["score"].tolist() == want["score"]

    for mode in (api.Mode.MAXP, api.Mode.AVEP, api.Mode.FIRSTP):
        loaded = api.OnDiskIndex.load(path, api.ones, mode=mode)
        assert len(loaded) == 5 and loaded.doc_ids == set(DOC) and loaded.psg_ids == set(PSG)
        assert loaded._store.rows_for(["d0"], "MAXP")[0].tolist() == [0, 1]
        got = loaded(rank)
        assert got._df["score"].tolist() == golden_kat[f"full/{mode.name}"]["score"]
        assert got._df["id"].tolist() == golden_kat[f"full/{mode.name}"]["id"]
    loaded.mode = api.Mode.PASSAGE
    vecs, ids = loaded._get_vectors(PSG)
    assert np.array_equal(vecs, V) and ids == PSG

    mem = loaded.to_memory(batch_size=2)
    assert isinstance(mem, api.InMemoryIndex) and len(mem) == 5 and mem.doc_ids == set(DOC)
    mem.mode = api.Mode.MAXP
    assert mem(rank) == index(rank)
    # a deep copy (rows device to device, id dictionaries cloned): the two indexes grow apart
    assert mem._store.dev is not loaded._store.dev and mem.psg_ids == set(PSG)
    mem.add(V[:1] * 3, doc_ids=["d0"], psg_ids=["p-extra"])
    assert len(mem) == 6 and len(loaded) == 5 and "p-extra" not in loaded.psg_ids
    loaded.mode = api.Mode.MAXP
    assert loaded(rank) == index(rank)
    mem.mode = api.Mode.PASSAGE
    vecs, ids = mem._get_vectors(PSG + ["p-extra"])
    assert np.array_equal(vecs[:5], V) and np.array_equal(vecs[5], V[0] * 3) and ids == PSG + ["p-extra"]
    empty = api.OnDiskIndex(tmp_path / "empty.h5")
    assert len(api.OnDiskIndex.load(tmp_path / "empty.h5")) == 0 and empty.dim is None


def test_partial_ids_and_id_length(api, tmp_path):
    index = api.OnDiskIndex(tmp_path / "p.h5", max_id_length=3)
    index.add(V, doc_ids=[None, None] + DOC[2:], psg_ids=PSG[:-2] + [None, None])
    index.add(V[:2], doc_ids=DOC[:2])
    with pytest.raises(RuntimeError):
        index.add(V[:1], doc_ids=["long_id"])
    with pytest.raises(RuntimeError):
        index.add(V[:1], psg_ids=["p0"])
    assert len(index) == 7
    loaded = api.OnDiskIndex.load(tmp_path / "p.h5")
    assert loaded.doc_ids == set(DOC) and loaded.psg_ids == set(PSG[:3])
    assert loaded._store.rows_for(["d0"], "MAXP")[0].tolist() == [5, 6]
    assert loaded._store.rows_for(["p2"], "PASSAGE")[0].tolist() == [2]
    docs, psgs = zip(*[(d, p) for _, d, p in loaded])
    assert list(docs) == [None, None, "d1", "d2", "d3", "d0", "d0"]
    assert list(psgs) == ["p0", "p1", "p2", None, None, None, None]


def test_quantizer_is_persisted(api, tmp_path):
    rng = np.random.default_rng(0)
    pq = api.NanoPQ(2, 8)
    pq.fit(rng.normal(size=(32, 16)).astype(np.float32), iter=3)
    index = api.OnDiskIndex(tmp_path / "q.h5", quantizer=pq)
    x = rng.normal(size=(5, 16)).astype(np.float32)
    index.add(x, doc_ids=DOC)
    loaded = api.OnDiskIndex.load(tmp_path / "q.h5")
    assert loaded.quantizer == pq and loaded.dim == 16 and loaded._get_internal_dim() == 2
    loaded.mode = api.Mode.MAXP
    codes, _ = loaded._get_vectors(["d0", "d1", "d2", "d3"])
    assert np.array_equal(codes, pq.encode(x))


def test_load_falls_back_to_h5py_for_files_the_native_reader_refuses(api, tmp_path, monkeypatch):
    """The native HDF5 reader covers what h5py's defaults write; a file with anything else
    (libver="latest" chunk indexes, filters, dense link storage) makes it return
    FFX_ERR_UNSUPPORTED — `load` then reads the file through h5py when that is installed, and
    fails with the reader's message when it is not."""
    import fast_forward.index.disk as disk
    from fast_forward import _ffx, _h5

    path = tmp_path / "index.h5"
    rng = np.random.default_rng(0)
    vec = rng.standard_normal((300, 32)).astype(np.float32)
    doc_ids = [f"d{i // 3}" for i in range(300)]
    psg_ids = [None if i % 7 == 0 else f"p{i}" for i in range(300)]
    index = api.OnDiskIndex(path, api.ones, chunk_size=64, init_size=64)
    index.add(vec, doc_ids=doc_ids, psg_ids=psg_ids)
    want = api.OnDiskIndex.load(path)

    class Refusing:
        def __init__(self, *a, **kw):
            raise _ffx.FFXError(-5, "chunk index version 4 is not supported")

    class Wrapped:
        H5File = Refusing

    monkeypatch.setattr(disk, "_h5", Wrapped)
    if disk._h5py() is None:  # native-writer variant: no h5py, the reader's error surfaces
        with pytest.raises(_ffx.FFXError, match="not supported"):
            api.OnDiskIndex.load(path)
        return
    got = api.OnDiskIndex.load(path)
    monkeypatch.setattr(disk, "_h5", _h5)
    assert len(got) == len(want) == 300 and got.doc_ids == want.doc_ids and got.psg_ids == want.psg_ids
    ids = sorted(want.doc_ids)[::5]
    v1, i1 = got._get_vectors(ids)
    v2, i2 = want._get_vectors(ids)
    assert i1 == i2 and (v1 == v2).all()
    assert [(v.tobytes(), d, p) for v, d, p in got] == [(v.tobytes(), d, p) for v, d, p in want]


def test_wide_codes_are_refused_up_front(api, tmp_path):
    """Ks > 256 gives uint16 codes, which the device kernels do not score: `add` says so at once
    instead of storing float copies that fail later (the in-memory index does the same)."""
    quant = api.NanoPQ(2, 300)
    quant.fit(np.random.default_rng(1).standard_normal((400, 8)).astype(np.float32), iter=2)
    assert quant.dtype == np.uint16
    for index in (api.OnDiskIndex(tmp_path / "wide.h5", api.ones, quantizer=quant), api.InMemoryIndex(api.ones, quantizer=quant)):
        with pytest.raises((NotImplementedError, RuntimeError), match="Ks"):
            index.add(np.ones((3, 8), np.float32), doc_ids=["a", "b", "c"])
        assert len(index) == 0
