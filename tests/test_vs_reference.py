"""Live cross-check on the GPU box: the UNMODIFIED reference package (baseline/_ref, installed
offline by __graft_entry__.build(); it travels with the repository snapshot) is run in its own
interpreter on a seeded random corpus, and every frame it produces — `index(ranking)`,
`interpolate`, `cut`, `index(ranking, early_stopping=...)`, all four modes — is compared with
the drop-in API's, ids in order and scores bit for bit.  Larger and differently seeded than
the committed goldens; skipped where baseline/_ref does not exist (a fresh clone)."""

import json
import os
import subprocess
import sys

import numpy as np
import pandas as pd
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
REF = os.path.join(ROOT, "baseline", "_ref", "fast_forward")


@pytest.mark.skipif(not os.path.isdir(REF), reason="baseline/_ref (the reference package) is not installed here")
@pytest.mark.parametrize("seed,dim,variant", [(1, 768, "plain"), (2, 384, "plain"), (3, 100, "plain"),
                                              (4, 768, "scattered+batched"), (5, 768, "float64"),
                                              (6, 100, "scattered+batched"), (7, 768, "coded"),
                                              (8, 100, "coded+scattered+batched"), (9, 384, "coded+float64")])
def test_frames_equal_the_reference(tmp_path, monkeypatch, seed, dim, variant):
    import __graft_entry__ as g

    g.build()
    import fast_forward
    import fast_forward.ranking as rk
    from fast_forward.encoder import TableEncoder
    from fast_forward.index import InMemoryIndex, Mode

    if "coded" in variant:  # rankings on integer-coded columns, as at the large sizes (ranking._CODED_FROM)
        monkeypatch.setattr(rk, "_CODED_FROM", 0)
    rng = np.random.default_rng(seed)
    n_docs, nq, C = 3000, 40, 600
    cnt = rng.integers(1, 9, n_docs)
    vectors = rng.standard_normal((int(cnt.sum()), dim)).astype(np.float32)
    doc_ids = np.repeat([f"d{i}" for i in range(n_docs)], cnt)
    split = len(vectors)
    extra = {}
    if "scattered" in variant:  # the last quarter of the rows extends random earlier documents
        split = len(vectors) - len(vectors) // 4
        doc_ids[split:] = [f"d{int(x)}" for x in rng.integers(0, n_docs, len(vectors) - split)]
        extra = {"split": split, "batch_size": 7}
    score_dtype = np.float64 if "float64" in variant else np.float32
    if "float64" in variant:
        extra["score_dtype"] = "float64"
    qvecs = rng.standard_normal((nq, dim)).astype(np.float32)
    cols = {}
    for key, pool in (("doc", sorted(set(doc_ids.tolist()))), ("psg", [f"p{i}" for i in range(len(vectors))])):
        picks = np.concatenate([rng.choice(len(pool), C, replace=False) for _ in range(nq)])
        cols[f"{key}_q_id"] = np.repeat([f"q{i}" for i in range(nq)], C)
        cols[f"{key}_id"] = np.array(pool)[picks]
        # coarse, steeply falling first-stage scores: ties, and early stopping really stops
        depth = np.tile(np.arange(C), nq)
        cols[f"{key}_score"] = (np.round(150.0 * 0.97 ** depth * rng.uniform(0.5, 1.5, nq).repeat(C), 0)).astype(np.float32)
    bundle = tmp_path / "in.npz"
    np.savez(bundle, vectors=vectors, doc_ids=doc_ids, qvecs=qvecs, alpha=0.3, cutoff=25, es_cutoff=10,
             es_alpha=0.5, es_depths=np.array([20, 50, 150, 600]), **cols, **extra)
    out_file = tmp_path / "out.json"
    env = {k: v for k, v in os.environ.items() if k != "PYTHONPATH"}
    run = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ref_run.py"), str(bundle), str(out_file)],
                         capture_output=True, text=True, env=env, timeout=900)
    assert run.returncode == 0, run.stderr[-3000:]
    want = json.load(open(out_file))

    queries = {f"q{i}": f"text {i}" for i in range(nq)}
    index = InMemoryIndex(TableEncoder({f"text {i}": qvecs[i] for i in range(nq)}), init_size=64, alloc_size=500)
    psg_ids = [f"p{i}" for i in range(len(vectors))]
    index.add(vectors[:split], doc_ids=doc_ids[:split].tolist(), psg_ids=psg_ids[:split])
    if split < len(vectors):
        index.add(vectors[split:], doc_ids=doc_ids[split:].tolist(), psg_ids=psg_ids[split:])
    batch_size = extra.get("batch_size")

    def same(r, w, what):
        df = r._df
        assert df["q_id"].tolist() == w["q_id"] and df["id"].tolist() == w["id"], what
        assert str(df["score"].dtype) == w["dtype"], what
        assert df["score"].to_numpy().astype(np.float64).view(np.uint64).tolist() == w["score_bits"], what

    for mode in (Mode.MAXP, Mode.AVEP, Mode.FIRSTP, Mode.PASSAGE):
        key = "psg" if mode == Mode.PASSAGE else "doc"
        first = fast_forward.Ranking(pd.DataFrame({"q_id": cols[f"{key}_q_id"], "id": cols[f"{key}_id"],
                                                   "score": cols[f"{key}_score"]}), queries=queries, dtype=score_dtype)
        w = want[mode.name]
        same(first, w["first"], (mode.name, "first stage"))
        index.mode = mode
        scored = index(first, batch_size=batch_size)
        same(scored, w["ff"], (mode.name, "ff"))
        inter = first.interpolate(scored, 0.3)
        same(inter, w["interpolated"], (mode.name, "interpolated"))
        same(inter.cut(25), w["cut"], (mode.name, "cut"))
        same(index.rerank(first, 0.3, cutoff=25), w["cut"], (mode.name, "fused rerank"))
        es = index(first, early_stopping=10, early_stopping_alpha=0.5, early_stopping_depths=(20, 50, 150, 600))
        same(es, w["early_stopping"], (mode.name, "early stopping"))
        assert 0 < len(es._df) < len(first._df)  # some queries stopped early, none scored nothing
