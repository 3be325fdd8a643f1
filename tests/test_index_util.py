"""Host-side id utilities of `fast_forward.index.util` (no GPU): `get_indices` and `ChunkIndexer`
(reference: src/fast_forward/index/util.py)."""

import numpy as np
import pytest

import __graft_entry__ as g

g.build()


def test_chunk_indexer_matches_the_reference_class():
    """`index.util.ChunkIndexer` (a host utility of the reference's chunked indexes, kept for code
    written against it): same vectors and owner ids as the unmodified reference class where
    baseline/_ref is installed, and as a flat gather otherwise."""
    import importlib.util
    import os
    import sys

    from fast_forward.index import Mode
    from fast_forward.index.util import ChunkIndexer

    rng = np.random.default_rng(12)
    sizes = [7, 3, 3, 3, 2]  # the first chunk longer, the last one partly filled
    chunks = [rng.standard_normal((s, 4)).astype(np.float32) for s in sizes]
    flat = np.concatenate(chunks)
    n = len(flat)
    doc_of = rng.integers(0, 6, n)
    docs = {f"d{d}": np.flatnonzero(doc_of == d).tolist() for d in range(6) if (doc_of == d).any()}
    psgs = {f"p{i}": i for i in range(n)}
    ours = ChunkIndexer(chunks, docs, psgs)
    assert [ours._get_chunk_indices(i) for i in (0, 6, 7, 9, 10, 17)] == [(0, 0), (0, 6), (1, 0), (1, 2), (2, 0), (4, 1)]
    path = os.path.join(os.path.dirname(__file__), "..", "baseline", "_ref", "fast_forward", "index", "util.py")
    ref = None
    if os.path.exists(path):
        spec = importlib.util.spec_from_file_location("_reference_index_util", path)
        module = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(module)  # imports Mode from fast_forward.index: ours, same members
        ref = module.ChunkIndexer(chunks, docs, psgs)
    for mode in (Mode.MAXP, Mode.FIRSTP, Mode.AVEP, Mode.PASSAGE):
        pool = list(psgs) if mode is Mode.PASSAGE else list(docs)
        for _ in range(5):
            ids = [pool[i] for i in rng.integers(0, len(pool), rng.integers(0, 8))]
            vec, owners = ours(ids, mode)
            if ref is not None:
                want_vec, want_owners = ref(ids, mode)
                assert owners == want_owners and np.array_equal(vec, want_vec)
            elif ids:
                assert sorted(owners) == sorted(o for i in ids for o in [i] * (1 if mode in (Mode.PASSAGE, Mode.FIRSTP) else len(docs[i])))
                assert len(vec) == len(owners)
    with pytest.raises(IndexError):
        ours(["nope"], Mode.MAXP)
    one = ChunkIndexer([flat], docs, psgs)
    vec, owners = one(["d0", "d1"] if "d0" in docs and "d1" in docs else list(docs)[:2], Mode.MAXP)
    assert len(vec) == len(owners)
