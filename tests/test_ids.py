"""Host-side id coding (`ffx_dict_*`, `ffx_csr_build`; no GPU needed) against a plain Python
model of the reference's two dictionaries (index/memory.py:84-95, index/util.py:29-41)."""

import numpy as np
import pandas as pd
import pyarrow as pa
import pytest


@pytest.fixture(scope="module")
def ids():
    import __graft_entry__ as g

    g.build()
    from fast_forward import _ids

    return _ids


def random_ids(rng, n, pool):
    alphabet = np.array(list("abcXYZ019-_é漢"))  # multi-byte UTF-8 too
    words = ["".join(rng.choice(alphabet, rng.integers(0, 12))) for _ in range(pool)]  # may contain ""
    return [words[i] for i in rng.integers(0, pool, n)]


def test_document_ordinals_follow_first_appearance(ids):
    rng = np.random.default_rng(0)
    d = ids.IdDict()
    model: dict[str, int] = {}
    for step in range(6):
        batch = random_ids(rng, 500, 300)
        batch = [None if rng.random() < 0.1 else b for b in batch]
        got = d.insert_ordinal(batch)
        want = [(-1 if b is None else model.setdefault(b, len(model))) for b in batch]
        assert got.tolist() == want
    assert d.keys() == list(model)
    keys, values = d.export()
    assert values.tolist() == list(range(len(model))) and keys.to_pylist() == list(model)


def test_passage_ids_are_unique_and_all_or_nothing(ids):
    d = ids.IdDict()
    assert d.insert_unique(["p0", None, "p2"], 10) == -1
    assert len(d) == 2
    assert d.insert_unique(["p7", "p0"], 13) == 1 and len(d) == 2  # known key: nothing inserted
    assert d.insert_unique(["a", "b", "a"], 13) == 2 and len(d) == 2  # repeat inside the batch
    assert d.insert_unique(["a", "b"], 13, dry_run=True) == -1 and len(d) == 2
    assert d.insert_unique(["a", "b"], 13) == -1
    codes, missing = d.lookup(["b", "p2", "p0", "a"])
    assert codes.tolist() == [14, 12, 10, 13] and missing == -1
    # growth across a rehash keeps everything reachable
    many = [f"q{i}" for i in range(5000)]
    assert d.insert_unique(many, 100) == -1
    codes, missing = d.lookup(many + ["p0"])
    assert codes.tolist() == list(range(100, 5100)) + [10] and missing == -1


@pytest.mark.parametrize("threads", [1, 0])
def test_lookup_over_every_column_flavour(ids, threads):
    rng = np.random.default_rng(1)
    keys = sorted(set(random_ids(rng, 4000, 3000)))
    d = ids.IdDict()
    d.insert_ordinal(keys)
    model = {k: i for i, k in enumerate(keys)}
    probe = [keys[i] for i in rng.integers(0, len(keys), 50_000)]
    probe[123] = "definitely-missing"
    probe[77] = None
    want = [model.get(p, -1) if p is not None else -1 for p in probe]
    first_missing = 77
    flavours = {
        "list": probe,
        "object array": np.array(probe, dtype=object),
        "pandas str column": pd.Series(probe, dtype="str"),
        "pandas object column": pd.Series(probe, dtype=object),
        "arrow string (32-bit offsets)": pa.array(probe, type=pa.string()),
        "arrow chunked": pa.chunked_array([pa.array(probe[:20_000], type=pa.large_string()),
                                           pa.array(probe[20_000:], type=pa.large_string())]),
    }
    for name, col in flavours.items():
        codes, missing = d.lookup(col, threads=threads)
        assert codes.tolist() == want, name
        assert missing == first_missing, name
    # a slice of an Arrow array: non-zero offset into offsets AND validity bitmap
    sliced = pa.array(probe, type=pa.large_string()).slice(70, 1000)
    codes, missing = d.lookup(sliced, threads=threads)
    assert codes.tolist() == want[70:1070] and missing == 7
    series_slice = pd.Series(probe, dtype="str").iloc[1000:3000]
    assert d.lookup(series_slice, threads=threads)[0].tolist() == want[1000:3000]


def test_csr_matches_a_stable_sort(ids):
    rng = np.random.default_rng(2)
    n_docs = 400
    row_doc = rng.integers(-1, n_docs, 5000)
    row_doc[row_doc == 17] = 16  # a document without rows in the middle
    off, rows = ids.csr_from_ordinals(row_doc, n_docs)
    has = np.flatnonzero(row_doc >= 0)
    order = has[np.argsort(row_doc[has], kind="stable")]
    assert rows.tolist() == order.tolist()
    assert off.tolist() == np.concatenate([[0], np.cumsum(np.bincount(row_doc[has], minlength=n_docs))]).tolist()
    off, rows = ids.csr_from_ordinals(np.zeros(0, np.int64), 0)
    assert off.tolist() == [0] and len(rows) == 0


def test_errors_report_through_ffx_last_error(ids):
    from fast_forward import _ffx

    with pytest.raises(_ffx.FFXError):
        ids.csr_from_ordinals(np.array([0, 5]), 3)  # ordinal out of range


def test_fixed_width_id_columns_become_arrow_strings():
    """The HDF5 id columns (`S{max_id_length}`, b"" = no id; index/disk.py:152-165,414-417) are
    turned into Arrow string buffers by ffx_fixed_width_to_arrow, without one Python object per id."""
    import __graft_entry__ as g

    g.build()
    from fast_forward.index.disk import _text_ids

    rng = np.random.default_rng(0)
    for width in (1, 3, 8, 21):
        words = ["".join(chr(97 + c) for c in rng.integers(0, 26, rng.integers(0, width + 1))) for _ in range(3000)]
        raw = np.array([w.encode() for w in words], dtype=f"S{width}")
        got = _text_ids(raw)
        assert got.to_pylist() == [w or None for w in words] and got.null_count == words.count("")
    assert _text_ids(np.array([b"abcdefgh", b"", b"a\0b", b"\0x"], dtype="S8")).to_pylist() == ["abcdefgh", None, "a\0b", "\0x"]
    assert len(_text_ids(np.zeros(0, "S8"))) == 0
    assert _text_ids(np.array(["grüße".encode()], dtype="S8")).to_pylist() == ["grüße"]
