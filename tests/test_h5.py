"""The native HDF5 reader (csrc/ffx_h5.cpp, `fast_forward._h5.H5File`) — the h5py-free staging
source of `OnDiskIndex.load` (reference: index/disk.py:355-418; layout :83-85,138-165).

No libhdf5 exists in this image, so the files come from tests/h5_writer.py, which restates the
structures h5py's defaults produce (see its header: parity with libhdf5 itself is unpinned).
Runs without a GPU: the reader is host code."""

import os
import sys

import numpy as np
import pytest

import h5_writer as hw

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "fast-forward-indexes_b200"))


@pytest.fixture(scope="module")
def h5():
    import __graft_entry__ as g

    g.build()
    from fast_forward import _ffx, _h5

    _h5.FFXError = _ffx.FFXError
    return _h5


def index_tree(rng, n, dim, chunk_rows, id_chunk, width=8, capacity=None, dtype=np.float32, missing=None):
    """The objects OnDiskIndex writes: root attrs, vectors / doc_ids / psg_ids."""
    capacity = capacity or n
    vec = np.zeros((capacity, dim), dtype)
    vec[:n] = rng.standard_normal((n, dim)).astype(dtype) if np.dtype(dtype).kind == "f" else rng.integers(0, 200, (n, dim))
    docs = np.zeros(capacity, f"S{width}")
    psgs = np.zeros(capacity, f"S{width}")
    docs[:n] = [f"d{i // 3}".encode() for i in range(n)]
    psgs[:n] = [b"" if i % 7 == 3 else f"p{i}".encode() for i in range(n)]
    root = hw.Group()
    root.attrs = {"num_vectors": np.int64(n), "ff_version": "0.8.0"}
    root.children["vectors"] = hw.Dataset(vec, (chunk_rows, dim), (None, dim), missing_chunks=missing)
    root.children["doc_ids"] = hw.Dataset(docs, (id_chunk,), (None,))
    root.children["psg_ids"] = hw.Dataset(psgs, (id_chunk,), (None,))
    return root, vec, docs, psgs


@pytest.mark.parametrize("modern", [False, True])
@pytest.mark.parametrize("split", [False, True])
def test_index_layout_round_trip(h5, tmp_path, modern, split):
    rng = np.random.default_rng(1)
    root, vec, docs, psgs = index_tree(rng, 1000, 24, 64, 128, capacity=1024)
    path = tmp_path / "index.h5"
    hw.write_hdf5(root, path, modern=modern, split_headers=split)
    with h5.H5File(path) as fp:
        assert sorted(fp.keys()) == ["doc_ids", "psg_ids", "vectors"]
        assert "vectors" in fp and "quantizer" not in fp and "vectors/x" not in fp
        assert fp.kind("/") == 1 and fp.kind("vectors") == 2
        attrs = fp.attrs("/")
        assert attrs["num_vectors"] == 1000 and attrs["num_vectors"].dtype == np.int64
        assert attrs["ff_version"] == "0.8.0"
        meta = fp.info("vectors")
        assert meta["shape"] == (1024, 24) and meta["dtype"] == np.float32 and meta["chunk_rows"] == 64
        assert fp.info("doc_ids")["dtype"] == np.dtype("S8")
        assert (fp.read("vectors") == vec).all()
        assert (fp.read("vectors", 100, 333) == vec[100:333]).all()
        assert (fp.read("doc_ids") == docs).all() and (fp.read("psg_ids", 5, 900) == psgs[5:900]).all()
        at = 0
        for row, block in fp.spans("vectors", 0, 1000):  # one run per HDF5 chunk, in place
            assert row == at and len(block) <= 64 and not block.flags.writeable
            assert (block == vec[row:row + len(block)]).all()
            at += len(block)
        assert at == 1000
        assert [r for r, _ in fp.spans("vectors", 70, 200)] == [70, 128, 192]


@pytest.mark.parametrize("chunk_k", [1, 2, 32])
def test_deep_chunk_trees_and_unwritten_chunks(h5, tmp_path, chunk_k):
    """Many chunks -> internal B-tree levels (2K entries per node); chunks that were never
    written have no entry and read as zeros (HDF5's fill value)."""
    rng = np.random.default_rng(2)
    missing = {3, 17, 40}
    root, vec, _, _ = index_tree(rng, 5 * 97, 6, 5, 11, missing=missing)
    for c in missing:
        vec[5 * c:5 * c + 5] = 0
    path = tmp_path / "deep.h5"
    hw.write_hdf5(root, path, chunk_k=chunk_k)
    with h5.H5File(path) as fp:
        assert (fp.read("vectors") == vec).all()
        zero_runs = [r for r, b in fp.spans("vectors", 0, len(vec)) if b.flags.writeable]  # synthesised zeros
        assert zero_runs == [15, 85, 200]


def test_wide_groups_and_nested_paths(h5, tmp_path):
    """More links than one symbol-table node holds: several SNODs under a multi-level B-tree."""
    root = hw.Group()
    want = {}
    for i in range(150):
        want[f"item{i:03d}"] = np.arange(i + 1, dtype=np.int32)
        root.children[f"item{i:03d}"] = hw.Dataset(want[f"item{i:03d}"])
    deep = hw.Group()
    deep.children["leaf"] = hw.Dataset(np.float64(2.5).reshape(()))
    mid = hw.Group()
    mid.children["deep"] = deep
    root.children["mid"] = mid
    for modern in (False, True):
        path = tmp_path / f"wide{modern}.h5"
        hw.write_hdf5(root, path, modern=modern, group_leaf_k=2, group_node_k=2)
        with h5.H5File(path) as fp:
            assert sorted(fp.keys()) == sorted(list(want) + ["mid"])
            for name, arr in want.items():
                assert (fp.read(name) == arr).all()
            assert fp.read("/mid/deep/leaf") == 2.5 and fp.read("mid//deep/leaf/").shape == ()
            assert fp.kind("mid/deep") == 1 and fp.kind("mid/nope") == 0 and fp.keys("mid") == ["deep"]


@pytest.mark.parametrize("modern", [False, True])
def test_quantizer_group(h5, tmp_path, modern):
    """disk.py:123-136 — state in attributes (str, int, bool) and small contiguous datasets."""
    rng = np.random.default_rng(3)
    cw = rng.standard_normal((4, 16, 2)).astype(np.float32)
    rot = rng.standard_normal((8, 8)).astype(np.float32)
    root = hw.Group()
    root.attrs = {"num_vectors": np.int64(0), "ff_version": "0.8.0"}
    q = hw.Group()
    meta, attributes, data = hw.Group(), hw.Group(), hw.Group()
    meta.attrs = {"__module__": "fast_forward.quantizer.nanopq", "__name__": "NanoOPQ", "_trained": True}
    attributes.attrs = {"M": np.int64(4), "Ks": np.int64(16), "metric": "dot", "verbose": False,
                        "ratio": np.float32(0.25), "weights": np.arange(6, dtype=np.float64).reshape(2, 3),
                        "tags": np.array([b"ab", b"c"]), "unicode": "grüße ✓"}
    data.children["codewords"] = hw.Dataset(cw)
    data.children["R"] = hw.Dataset(rot)
    data.children["tiny"] = hw.Dataset(np.arange(5, dtype=np.uint8), compact=True)
    data.children["empty"] = hw.Dataset(np.zeros((0, 3), np.float32))
    q.children.update(meta=meta, attributes=attributes, data=data)
    root.children["quantizer"] = q
    path = tmp_path / "q.h5"
    hw.write_hdf5(root, path, modern=modern, split_headers=True)
    with h5.H5File(path) as fp:
        assert "quantizer" in fp and sorted(fp.keys("quantizer")) == ["attributes", "data", "meta"]
        m = fp.attrs("quantizer/meta")
        assert m == {"__module__": "fast_forward.quantizer.nanopq", "__name__": "NanoOPQ", "_trained": True}
        assert m["_trained"].dtype == np.bool_
        a = fp.attrs("quantizer/attributes")
        assert a["M"] == 4 and a["Ks"] == 16 and a["metric"] == "dot" and not a["verbose"]
        assert a["ratio"] == np.float32(0.25) and a["ratio"].dtype == np.float32
        assert (a["weights"] == np.arange(6.0).reshape(2, 3)).all() and a["unicode"] == "grüße ✓"
        assert list(a["tags"]) == ["ab", "c"]
        assert (fp.read("quantizer/data/codewords") == cw).all() and fp.read("quantizer/data/R").shape == (8, 8)
        assert (fp.read("quantizer/data/tiny") == np.arange(5)).all()
        assert fp.read("quantizer/data/empty").shape == (0, 3)
        with pytest.raises(h5.FFXError):
            fp.attr("quantizer/meta", "absent")
        with pytest.raises(h5.FFXError):
            fp.read("quantizer/data")  # a group


@pytest.mark.parametrize("dtype", [np.float16, np.float64, np.int64, np.uint8, np.int16])
def test_other_vector_types(h5, tmp_path, dtype):
    """The reference stores whatever dtype the first `add` had (disk.py:147)."""
    rng = np.random.default_rng(4)
    root, vec, _, _ = index_tree(rng, 100, 8, 16, 16, dtype=dtype)
    path = tmp_path / "t.h5"
    hw.write_hdf5(root, path)
    with h5.H5File(path) as fp:
        got = fp.read("vectors")
        assert got.dtype == np.dtype(dtype) and (got == vec).all()


def test_user_block_and_long_ids(h5, tmp_path):
    rng = np.random.default_rng(5)
    root, vec, docs, _ = index_tree(rng, 50, 4, 8, 8, width=21)
    path = tmp_path / "ub.h5"
    hw.write_hdf5(root, path, user_block=1024)
    with h5.H5File(path) as fp:
        assert (fp.read("vectors") == vec).all() and fp.info("doc_ids")["dtype"].itemsize == 21
        assert (fp.read("doc_ids") == docs).all()


def test_damaged_files_are_errors_not_faults(h5, tmp_path):
    rng = np.random.default_rng(6)
    root, _, _, _ = index_tree(rng, 300, 16, 32, 32)
    good = tmp_path / "good.h5"
    hw.write_hdf5(root, good, split_headers=True)
    raw = good.read_bytes()

    def opens(data, name):
        p = tmp_path / name
        p.write_bytes(data)
        with h5.H5File(p) as fp:
            fp.attrs("/")
            for k in fp.keys():
                fp.read(k)

    with pytest.raises(h5.FFXError, match="not an HDF5 file"):
        opens(b"\0" * 4096, "zeros.h5")
    with pytest.raises(h5.FFXError, match="not an HDF5 file"):
        opens(b"short", "short.h5")
    with pytest.raises(h5.FFXError):
        h5.H5File(tmp_path / "does-not-exist.h5")
    for cut in (100, 700, len(raw) // 3, len(raw) // 2, len(raw) - 40):
        with pytest.raises(h5.FFXError):
            opens(raw[:cut], f"cut{cut}.h5")
    # random corruption of the metadata must never crash the process; data bytes may change silently
    flips = np.random.default_rng(7)
    meta_from = max(i for i in range(len(raw) - 4) if raw[i:i + 4] in (b"TREE", b"HEAP", b"SNOD")) - 4096
    for trial in range(1500):
        data = bytearray(raw)
        # two thirds of the trials hit the structures (B-trees, heaps, headers sit behind the
        # chunk data in these files; the superblock in front), the rest anywhere
        where = (flips.integers(0, len(raw), 6) if trial % 3 == 0 else
                 np.concatenate([flips.integers(max(meta_from, 0), len(raw), 5), flips.integers(0, 96, 1)]))
        for pos in where:
            if trial % 2:
                data[pos] ^= 1 << int(flips.integers(0, 8))
            else:
                data[pos] = int(flips.integers(0, 256))
        try:
            opens(bytes(data), "flip.h5")
        except (h5.FFXError, ValueError, MemoryError):  # reported, or a datatype numpy cannot express
            pass


def test_unsupported_features_say_so(h5, tmp_path):
    rng = np.random.default_rng(8)
    root, _, _, _ = index_tree(rng, 64, 4, 16, 16)
    path = tmp_path / "f.h5"
    hw.write_hdf5(root, path)
    raw = bytearray(path.read_bytes())
    # a filter pipeline message in place of the modification-time message of `vectors`
    # (message type 0x12 -> 0x0B with one filter): compressed datasets are refused, not misread
    hits = [i for i in range(0, len(raw) - 16, 8) if raw[i:i + 4] == b"\x12\x00\x08\x00" and raw[i + 8] == 1]
    assert hits
    for i in hits:
        raw[i:i + 2] = b"\x0b\x00"
        raw[i + 8:i + 10] = b"\x01\x01"
    bad = tmp_path / "filtered.h5"
    bad.write_bytes(raw)
    with h5.H5File(bad) as fp:
        with pytest.raises(h5.FFXError, match="filter pipeline"):
            fp.read("vectors")


def test_native_writer_grows_reopens_and_reads_back(h5, tmp_path):
    """`fast_forward._h5_write.IndexFile` (OnDiskIndex's writer when h5py is missing) against the
    native reader: many small adds (two B-tree levels with 4-row chunks), capacity growth in whole
    chunks, ids with gaps, re-opening between adds, the quantizer group with every attribute type
    `Quantizer.serialize` produces, replacing the quantizer, and refusing foreign files."""
    from fast_forward import _h5_write as hw2

    rng = np.random.default_rng(9)
    path = tmp_path / "native.h5"
    hw2.IndexFile.create(path, "0.8.0").close()
    with h5.H5File(path) as fp:
        assert fp.attrs("/") == {"num_vectors": 0, "ff_version": "0.8.0", "ffx_writer": 1} and fp.keys() == []
    dim, chunk = 6, 4
    vec = rng.standard_normal((1000, dim)).astype(np.float32)
    docs = [None if i % 5 == 2 else f"d{i // 3}" for i in range(1000)]
    psgs = [None if i % 7 == 3 else f"p{i}" for i in range(1000)]
    have = 0
    for step in (1, 3, 4, 9, 200, 83, 700):
        with hw2.IndexFile.open_existing(path) as out:
            if not out.datasets:
                out.create_datasets(dim, np.float32, 8, chunk, 8)
            assert out.num_vectors == have
            if have + step > out.capacity:
                out.resize(-(-(have + step) // chunk) * chunk)
            for name, ids in (("doc_ids", docs), ("psg_ids", psgs)):
                part = ids[have:have + step]
                out.write_at(name, [have + i for i, v in enumerate(part) if v is not None], [v for v in part if v is not None])
            out.write_rows("vectors", have, vec[have:have + step])
            out.set_num_vectors(have + step)
        have += step
        with h5.H5File(path) as fp:
            assert fp.attr("/", "num_vectors") == have and fp.info("vectors")["chunk_rows"] == chunk
            assert (fp.read("vectors", 0, have) == vec[:have]).all()
            got_docs = [b.decode() or None for b in fp.read("doc_ids", 0, have)]
            got_psgs = [b.decode() or None for b in fp.read("psg_ids", 0, have)]
            assert got_docs == docs[:have] and got_psgs == psgs[:have]
            assert fp.info("vectors")["shape"][0] >= have and fp.info("vectors")["shape"][0] % chunk == 0
    cw = rng.standard_normal((2, 8, 3)).astype(np.float32)
    for trained in (False, True):
        with hw2.IndexFile.open_existing(path) as out:
            out.set_quantizer({"__module__": "fast_forward.quantizer.nanopq", "__name__": "NanoOPQ", "_trained": trained},
                              {"M": 2, "Ks": 8, "Ds": None if not trained else 3, "metric": "dot", "verbose": False,
                               "note": "grüße"}, {"codewords": cw, "R": np.eye(6, dtype=np.float32)} if trained else {})
        with h5.H5File(path) as fp:
            assert sorted(fp.keys()) == ["doc_ids", "psg_ids", "quantizer", "vectors"]
            meta, attrs = fp.attrs("quantizer/meta"), fp.attrs("quantizer/attributes")
            assert meta == {"__module__": "fast_forward.quantizer.nanopq", "__name__": "NanoOPQ", "_trained": trained}
            assert meta["_trained"].dtype == np.bool_ and attrs["M"] == 2 and attrs["note"] == "grüße"
            assert ("Ds" in attrs) == trained and attrs["metric"] == "dot" and not attrs["verbose"]
            assert sorted(fp.keys("quantizer/data")) == (["R", "codewords"] if trained else [])
            if trained:
                assert (fp.read("quantizer/data/codewords") == cw).all()
            assert (fp.read("vectors", 0, have) == vec[:have]).all()  # untouched by the new subtree
    foreign = tmp_path / "foreign.h5"
    root, _, _, _ = index_tree(rng, 10, 4, 4, 4)
    hw.write_hdf5(root, foreign)
    with pytest.raises(ValueError, match="needs h5py"):
        hw2.IndexFile.open_existing(foreign)


def test_byte_format_does_not_drift(tmp_path):
    """The writers are deterministic; their output is pinned by checksum so that a change of the
    byte format is a deliberate act (update the digests together with DESIGN.md section 3a)."""
    import hashlib

    from fast_forward import _h5_write as hw2

    rng = np.random.default_rng(123)
    root = hw.Group()
    root.attrs = {"num_vectors": np.int64(10), "ff_version": "0.8.0"}
    root.children["vectors"] = hw.Dataset(rng.standard_normal((12, 4)).astype(np.float32), (4, 4), (None, 4))
    root.children["doc_ids"] = hw.Dataset(np.array([f"d{i}".encode() for i in range(12)], "S8"), (8,), (None,))
    extra = hw.Group()
    extra.attrs = {"flag": True, "name": "x"}
    root.children["quantizer"] = extra
    digests = []
    for modern in (False, True):
        hw.write_hdf5(root, tmp_path / f"t{modern}.h5", modern=modern, split_headers=True)
        digests.append(hashlib.sha256((tmp_path / f"t{modern}.h5").read_bytes()).hexdigest())
    with hw2.IndexFile.create(tmp_path / "n.h5", "0.8.0") as out:
        out.create_datasets(4, np.float32, 4, 4, 8)
        out.resize(12)
        out.write_at("doc_ids", [0, 1, 5], ["a", "b", "c"])
        out.write_rows("vectors", 0, rng.standard_normal((10, 4)).astype(np.float32))
        out.set_num_vectors(10)
        out.set_quantizer({"__name__": "NanoPQ", "_trained": True}, {"M": 2, "metric": "dot"},
                          {"codewords": np.arange(8, dtype=np.float32)})
    digests.append(hashlib.sha256((tmp_path / "n.h5").read_bytes()).hexdigest())
    assert digests == ["5a0e3b75ddca12f515358bbd44d631b95d982164b3061976ecf257f797287a52",
                       "f3951d9f2d9e634b2810bc950dc606033dbe94a65fe2af974e3daace8dbfd47b",
                       "be113f9c0e76a6df47d251baee462973926a46671acf4780a64cf4860c5e1ee3"]


def test_large_reads_are_split_over_threads(h5, tmp_path):
    """ffx_h5_read_rows copies reads of 32 MB and more with several threads (page faults of a
    mapped file are per 4 KB): a 48 MB dataset with a partial last chunk and an unwritten one."""
    rng = np.random.default_rng(10)
    rows, dim, chunk = 12_345, 1024, 1000
    vec = rng.integers(0, 2**31, (rows, dim), dtype=np.int64).astype(np.float32)
    vec[3000:4000] = 0
    root = hw.Group()
    root.children["vectors"] = hw.Dataset(vec, (chunk, dim), (None, dim), missing_chunks={3})
    root.children["flat"] = hw.Dataset(vec[:9000])  # contiguous layout, 36 MB
    path = tmp_path / "big.h5"
    hw.write_hdf5(root, path)
    with h5.H5File(path) as fp:
        assert (fp.read("vectors") == vec).all()
        assert (fp.read("vectors", 17, 12_001) == vec[17:12_001]).all()
        assert (fp.read("flat") == vec[:9000]).all()
