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


def test_host_sort_and_match_routines_at_multithreaded_sizes():
    """ffx_ranking_order / ffx_order_u64 split their radix passes over the host cores from ~130 k
    rows up, ffx_first_repeat / ffx_match_keys prefetch ahead: checked against numpy at sizes
    where those paths are active, with heavy ties (stability) and special floats."""
    import ctypes as C

    import __graft_entry__ as g

    g.build()
    from fast_forward import _ffx

    lib = _ffx.lib()
    rng = np.random.default_rng(4)

    def ptr(a):
        return C.c_void_p(a.ctypes.data)

    for n in (1, 7, 70_000, 600_001):
        q_rank = rng.integers(0, 300, n).astype(np.int32)
        score = rng.choice(np.array([0.0, -0.0, 1.5, -2.25, 3e-7, 1e30, -1e30, np.inf, -np.inf], np.float32), n) \
            if n % 2 else rng.standard_normal(n).astype(np.float32)
        order = np.empty(n, np.int64)
        _ffx.check(lib.ffx_ranking_order(ptr(q_rank), ptr(score), n, ptr(order), 0))
        want = np.lexsort((np.arange(n), -(score + np.float32(0.0)).astype(np.float64), q_rank))
        assert (order == want).all()
        one = np.empty(n, np.int64)
        _ffx.check(lib.ffx_ranking_order(ptr(q_rank), ptr(score), n, ptr(one), 1))  # single-threaded: same order
        assert (one == order).all()

        keys = rng.integers(0, 2**63, n, dtype=np.uint64) if n % 2 else rng.integers(0, 50, n).astype(np.uint64) << np.uint64(40)
        _ffx.check(lib.ffx_order_u64(ptr(keys), n, ptr(order), 0))
        assert (order == np.argsort(keys, kind="stable")).all()

        have = rng.permutation(3 * n)[:n].astype(np.int64) - n  # distinct, some negative
        want_keys = np.concatenate([have[rng.integers(0, n, n // 2 + 1)], rng.integers(5 * n, 6 * n, n // 3)]).astype(np.int64)
        pos = np.empty(len(want_keys), np.int64)
        _ffx.check(lib.ffx_match_keys(ptr(have), n, ptr(want_keys), len(want_keys), ptr(pos)))
        lookup = {int(k): i for i, k in enumerate(have)}
        assert pos.tolist() == [lookup.get(int(k), -1) for k in want_keys]

        first = C.c_int64(0)
        _ffx.check(lib.ffx_first_repeat(ptr(have), n, C.byref(first)))
        assert first.value == -1
        if n > 2:
            again = have.copy()
            a, b = sorted(rng.choice(n, 2, replace=False).tolist())
            again[b] = again[a]
            _ffx.check(lib.ffx_first_repeat(ptr(again), n, C.byref(first)))
            assert first.value == b
    # more distinct queries than per-thread histograms are kept for: the LSD radix route
    n = 400_000
    q_rank = rng.integers(0, 20_000_000, n).astype(np.int32)
    q_rank[0:n - 1:3] = q_rank[1:n:3][: len(q_rank[0:n - 1:3])]  # some queries with several rows
    score = rng.choice(np.array([0.5, 1.5, -2.25], np.float32), n)
    order = np.empty(n, np.int64)
    _ffx.check(lib.ffx_ranking_order(ptr(q_rank), ptr(score), n, ptr(order), 0))
    assert (order == np.lexsort((np.arange(n), -score.astype(np.float64), q_rank))).all()
    bad = q_rank.copy()
    bad[17] = -1
    assert lib.ffx_ranking_order(ptr(bad), ptr(score), n, ptr(order), 0) != 0  # negative rank: refused

    nan_scores = np.array([1.0, np.nan, 2.0, np.nan, -1.0], np.float32)
    order = np.empty(5, np.int64)
    _ffx.check(lib.ffx_ranking_order(ptr(np.zeros(5, np.int32)), ptr(nan_scores), 5, ptr(order), 0))
    assert order.tolist() == [2, 0, 4, 1, 3]  # NaN after every number, in incoming order


@pytest.mark.parametrize("n, distinct", [(0, 1), (7, 3), (300_000, 41_000)])
def test_factorize_gives_equal_codes_to_equal_strings(ids, n, distinct):
    """`ffx_factorize` (all cores, hash-partitioned) against pandas.factorize: the numbering differs,
    the partition of the rows into equal strings must not; the exported keys are the distinct strings."""
    rng = np.random.default_rng(n)
    pool = np.array([f"doc-{i * 7919 % 100_003}" + "x" * (i % 5) for i in range(distinct)] + [""], dtype=object)
    col = pool[rng.integers(0, len(pool), n)]
    for values in (col, pd.array(col, dtype="string[pyarrow]"), pa.chunked_array([pa.array(col[: n // 2], pa.string()), pa.array(col[n // 2 :], pa.string())])):
        codes, keys = ids.factorize(values)
        assert codes.dtype == np.int32 and len(codes) == n
        want_codes, want_keys = pd.factorize(col)
        assert len(keys) == len(want_keys)
        assert sorted(keys.to_pylist()) == sorted(want_keys.tolist())
        if n:
            assert codes.min() >= 0 and codes.max() < len(keys)
            assert np.array_equal(np.asarray(keys.to_pylist(), dtype=object)[codes], col)
    with pytest.raises(ValueError):
        ids.factorize(pa.array(["a", None]))
