"""TEST INFRASTRUCTURE — generate tests/golden/* by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference; see oracle/ref_shim.py):

    python oracle/gen_golden.py

Outputs (committed):
  tests/golden/ref_kat.json       the reference's own known-answer fixtures
                                  (tests/test_index.py:19-46,135-200,273-349;
                                  tests/test_ranking.py:116-121,157-188) re-run through the
                                  reference, every resulting frame recorded row by row
  tests/golden/ref_random.npz     seeded random fp32 cases: inputs + the reference's
  tests/golden/ref_random.json    Index.__call__ / interpolate / cut outputs per Mode
  tests/golden/ref_es.npz/.json   seeded early-stopping cases (index/base.py:316-387): the
                                  reference's `Index.__call__(early_stopping=...)` frames

Scores are recorded as raw float32 bit patterns (uint32) so the comparison is bit-exact.
"""

import json
import os
import sys

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

ff = ref_shim.load_reference()
from fast_forward import Ranking  # noqa: E402
from fast_forward.encoder import LambdaEncoder  # noqa: E402
from fast_forward.encoder.base import Encoder  # noqa: E402
from fast_forward.index import InMemoryIndex, Mode  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")


def frame(r: Ranking):
    df = r._df
    return {
        "q_id": df["q_id"].tolist(),
        "id": df["id"].tolist(),
        "score_bits": df["score"].to_numpy().astype(np.float32).view(np.uint32).tolist(),
        "score": [float(x) for x in df["score"].to_numpy()],
    }


class TableEncoder(Encoder):
    """Precomputed query vectors keyed by query text (query encoding is outside the path)."""

    def __init__(self, table):
        self.table = table

    def _encode(self, texts):
        return np.stack([self.table[t] for t in texts])


# ------------------------------------------------------------------ known-answer fixtures
def kat():
    out = {}
    Q = {"q1": "query 1", "q2": "query 2"}
    DOC = ["d0", "d0", "d1", "d2", "d3"]
    PSG = ["p0", "p1", "p2", "p3", "p4"]
    V = np.tril(np.ones((5, 5), dtype=np.int64))
    doc_rank = Ranking.from_run(
        {"q1": {"d0": 100, "d1": 2, "d2": 3, "d3": 200}, "q2": {"d0": 400, "d1": 5, "d2": 6, "d3": 800}},
        queries=Q)
    psg_rank = Ranking.from_run(
        {"q1": {"p0": 100, "p1": 2, "p2": 3, "p3": 4, "p4": 5},
         "q2": {"p0": 500, "p1": 6, "p2": 7, "p3": 8, "p4": 9}}, queries=Q)
    enc = LambdaEncoder(lambda _: np.array([1, 1, 1, 1, 1]))

    full = InMemoryIndex(enc)
    full.add(V, doc_ids=DOC, psg_ids=PSG)
    partial = InMemoryIndex(enc)  # tests/test_index.py:58-69
    partial.add(V, doc_ids=[None, None] + DOC[2:], psg_ids=PSG[:-2] + [None, None])
    partial.add(V[:2], doc_ids=DOC[:2])
    partial.add(V[-2:], psg_ids=PSG[-2:])

    for name, idx in (("full", full), ("partial", partial)):
        for mode in (Mode.MAXP, Mode.FIRSTP, Mode.AVEP):
            idx.mode = mode
            out[f"{name}/{mode.name}"] = frame(idx(doc_rank))
        idx.mode = Mode.PASSAGE
        out[f"{name}/PASSAGE"] = frame(idx(psg_rank))
    out["input/doc_ranking"] = frame(doc_rank)
    out["input/psg_ranking"] = frame(psg_rank)

    # batch_size (tests/test_index.py:335-349); batch_size=5 divides the query count and
    # crashes the reference on the empty trailing batch only when < num_queries — 5 >= 5 is fine
    r5 = Ranking.from_run(
        {"q1": {"d0": 2, "d1": 3, "d2": 4, "d3": 10}, "q2": {"d0": 5, "d1": 4, "d2": 3, "d3": 12},
         "q3": {"d0": 8, "d1": 5, "d2": 2, "d3": 1}, "q4": {"d0": 11, "d1": 6, "d2": 1, "d3": 2},
         "q5": {"d0": 14, "d1": 7, "d2": 0, "d3": 3}},
        queries={f"q{n}": f"query {n}" for n in range(1, 6)})
    full.mode = Mode.MAXP
    out["batch/input"] = frame(r5)
    out["batch/none"] = frame(full(r5))
    out["batch/2"] = frame(full(r5, batch_size=2))

    # early stopping (tests/test_index.py:273-333)
    es = InMemoryIndex(LambdaEncoder(lambda q: np.array([10, 10])), mode=Mode.PASSAGE)
    es.add(np.stack([[1, 0], [1, 1]] * 10), psg_ids=[f"p{i}" for i in range(20)])
    r = Ranking(pd.DataFrame([{"q_id": q, "query": q, "id": f"p{i}", "score": i}
                              for i in range(20) for q in ("q1", "q2")]))
    out["es/input"] = frame(r)
    out["es/5_0.5_2-5-10-20"] = frame(es(r, early_stopping=5, early_stopping_alpha=0.5,
                                         early_stopping_depths=(2, 5, 10, 20)))
    out["es/3_0.2_4-8-20"] = frame(es(r, early_stopping=3, early_stopping_alpha=0.2,
                                      early_stopping_depths=(4, 8, 20)))

    # ranking algebra (tests/test_ranking.py)
    RUN = {"q1": {"d0": 1, "d1": 2, "d2": 300}, "q2": {"d0": 4, "d1": 5, "d2": 600, "d3": 7}}
    rk = Ranking.from_run(RUN)
    rkq = Ranking.from_run(RUN, queries=Q)
    out["rank/base"] = frame(rk)
    out["rank/cut2"] = frame(rk.cut(2))
    df = rkq._df.copy()
    df["score"] = list(map(np.float32, range(len(rk._df))))
    out["rank/interp_other"] = frame(Ranking(df))
    out["rank/interp0.5"] = frame(rk.interpolate(Ranking(df), 0.5))
    r4 = Ranking.from_run({"q1": {"d1": 1, "d2": 1}, "q2": {"d0": 1}})
    r5b = Ranking.from_run({"q1": {"d0": 1, "d1": 1}, "q3": {"d0": 1}})
    out["rank/r4"] = frame(r4)
    out["rank/r5"] = frame(r5b)
    out["rank/r4_interp_r5_0.5"] = frame(r4.interpolate(r5b, 0.5))
    out["rank/r4_plus_r5"] = frame(r4 + r5b)
    out["rank/normalize"] = frame(Ranking.from_run(
        {"q1": {"d0": 1, "d1": 2, "d2": 3}, "q2": {"d0": 4, "d1": 5, "d2": 6}}).normalize())
    out["rank/rr1"] = frame(rk.rr_scores(k=1))
    # q_id ordering is DESC on strings (q2 > q10 > q1) and ties keep frame order
    ties = pd.DataFrame({"q_id": ["q1", "q10", "q2", "q1", "q10", "q2", "q1"],
                         "id": ["a", "b", "c", "d", "e", "f", "g"],
                         "score": [1.0, 2.0, 3.0, 1.0, 2.0, 4.0, 1.0]})
    out["rank/ties_input"] = {"q_id": ties["q_id"].tolist(), "id": ties["id"].tolist(),
                              "score": ties["score"].tolist()}
    out["rank/ties_sorted"] = frame(Ranking(ties))
    return out


# ------------------------------------------------------------------ seeded random cases
def random_case(seed, dim, n_docs, max_psg, n_q, n_cand, alpha, cutoff, arrays, meta):
    rng = np.random.default_rng(seed)
    counts = rng.integers(1, max_psg + 1, size=n_docs)
    n_rows = int(counts.sum())
    vectors = rng.standard_normal((n_rows, dim)).astype(np.float32)
    doc_ids = [f"d{d}" for d, c in enumerate(counts) for _ in range(c)]
    psg_ids = [f"p{i}" for i in range(n_rows)]
    qvecs = rng.standard_normal((n_q, dim)).astype(np.float32)
    queries = {f"q{i}": f"text {i}" for i in range(n_q)}
    enc = TableEncoder({f"text {i}": qvecs[i] for i in range(n_q)})
    index = InMemoryIndex(enc, init_size=n_rows)
    # two add calls, the second one extends documents non-contiguously for the tail docs
    split = n_rows - n_rows // 4
    index.add(vectors[:split], doc_ids=doc_ids[:split], psg_ids=psg_ids[:split])
    tail_docs = [f"d{int(x)}" for x in rng.integers(0, n_docs, size=n_rows - split)]
    index.add(vectors[split:], doc_ids=tail_docs, psg_ids=psg_ids[split:])
    doc_ids = doc_ids[:split] + tail_docs

    key = f"s{seed}"
    arrays[f"{key}/vectors"] = vectors
    arrays[f"{key}/qvecs"] = qvecs
    case = {"dim": dim, "doc_ids": doc_ids, "psg_ids": psg_ids, "alpha": alpha, "cutoff": cutoff,
            "queries": queries, "split": split, "modes": {}}
    for mode in (Mode.MAXP, Mode.AVEP, Mode.FIRSTP, Mode.PASSAGE):
        pool = psg_ids if mode == Mode.PASSAGE else sorted(set(doc_ids))
        run = {}
        for qi in range(n_q):
            cands = rng.choice(len(pool), size=min(n_cand, len(pool)), replace=False)
            # coarse lexical scores so the first stage has ties too
            run[f"q{qi}"] = {pool[c]: float(np.float32(rng.integers(0, 40) / 2.0)) for c in cands}
        first = Ranking.from_run(run, queries=queries)
        index.mode = mode
        ff_out = index(first)
        inter = first.interpolate(ff_out, alpha)
        case["modes"][mode.name] = {
            "first_stage": frame(first),
            "ff": frame(ff_out),
            "interpolated": frame(inter),
            "cut": frame(inter.cut(cutoff)),
        }
    meta[key] = case


# ------------------------------------------------------------------ seeded early-stopping cases
def es_case(seed, dim, n_docs, max_psg, n_q, n_cand, arrays, meta):
    """`Index.__call__(early_stopping=...)` (index/base.py:316-387) on seeded random data:
    first-stage scores fall off steeply with depth, so some queries stop at the first depth,
    some later, some never.  Recorded: the reference's output frame per (mode, setting)."""
    rng = np.random.default_rng(seed)
    counts = rng.integers(1, max_psg + 1, size=n_docs)
    n_rows = int(counts.sum())
    vectors = rng.standard_normal((n_rows, dim)).astype(np.float32)
    doc_ids = [f"d{d}" for d, c in enumerate(counts) for _ in range(c)]
    psg_ids = [f"p{i}" for i in range(n_rows)]
    qvecs = rng.standard_normal((n_q, dim)).astype(np.float32)
    queries = {f"q{i}": f"text {i}" for i in range(n_q)}
    index = InMemoryIndex(TableEncoder({f"text {i}": qvecs[i] for i in range(n_q)}), init_size=n_rows)
    index.add(vectors, doc_ids=doc_ids, psg_ids=psg_ids)
    key = f"es{seed}"
    arrays[f"{key}/vectors"] = vectors
    arrays[f"{key}/qvecs"] = qvecs
    case = {"dim": dim, "doc_ids": doc_ids, "psg_ids": psg_ids, "queries": queries, "modes": {}}
    scale = float(np.sqrt(dim))
    for mode in (Mode.MAXP, Mode.AVEP, Mode.FIRSTP, Mode.PASSAGE):
        pool = psg_ids if mode == Mode.PASSAGE else sorted(set(doc_ids))
        run = {}
        for qi in range(n_q):
            cands = rng.choice(len(pool), size=min(n_cand, len(pool)), replace=False)
            # geometric fall-off, different steepness per query, on the scale of the dot products
            steep = rng.uniform(0.55, 0.98)
            lex = (6.0 * scale * steep ** np.arange(len(cands))).astype(np.float32)
            run[f"q{qi}"] = {pool[c]: float(x) for c, x in zip(cands, lex)}
        first = Ranking.from_run(run, queries=queries)
        index.mode = mode
        entry = {"first_stage": frame(first), "settings": []}
        for cutoff, alpha, depths in ((3, 0.5, (5, 10, 20, 40)), (5, 0.2, (8, 16, 64)), (2, 0.8, (1, 4, 4, 30)),
                                      (10, 0.5, (10, 12, 1000))):
            res = index(first, early_stopping=cutoff, early_stopping_alpha=alpha, early_stopping_depths=depths)
            entry["settings"].append({"cutoff": cutoff, "alpha": alpha, "depths": list(depths), "out": frame(res)})
            per_q = res._df.groupby("q_id").size().to_dict()
            print(key, mode.name, cutoff, alpha, depths, "rows scored per query:", per_q)
        case["modes"][mode.name] = entry
    meta[key] = case


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    with open(os.path.join(GOLDEN, "ref_kat.json"), "w") as f:
        json.dump(kat(), f, indent=0)
    arrays, meta = {}, {}
    random_case(11, 768, 48, 7, 5, 30, 0.1, 10, arrays, meta)
    random_case(12, 384, 40, 5, 4, 25, 0.3, 5, arrays, meta)
    random_case(13, 100, 30, 9, 3, 20, 0.5, 7, arrays, meta)  # non-uniform dim: generic kernel
    random_case(14, 1024, 20, 4, 3, 15, 0.05, 4, arrays, meta)
    np.savez(os.path.join(GOLDEN, "ref_random.npz"), **arrays)
    with open(os.path.join(GOLDEN, "ref_random.json"), "w") as f:
        json.dump(meta, f, indent=0)
    arrays, meta = {}, {}
    es_case(21, 768, 60, 5, 6, 45, arrays, meta)
    es_case(22, 100, 40, 4, 4, 30, arrays, meta)  # non-uniform dim: host-walked depths
    np.savez(os.path.join(GOLDEN, "ref_es.npz"), **arrays)
    with open(os.path.join(GOLDEN, "ref_es.json"), "w") as f:
        json.dump(meta, f, indent=0)
    for fn in sorted(os.listdir(GOLDEN)):
        print(fn, os.path.getsize(os.path.join(GOLDEN, fn)))


if __name__ == "__main__":
    main()
