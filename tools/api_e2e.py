"""Wall-clock of the drop-in Python API (strings in, Ranking out) at a scale where the host
shell matters: InMemoryIndex.add with string ids, then for a first-stage ranking of nq x C
pairs `index.rerank(ranking, alpha, cutoff)` and the reference's three-call idiom
`ranking.interpolate(index(ranking), alpha).cut(cutoff)`.  Prints one JSON line."""
import json
import os
import sys
import time

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fast-forward-indexes_b200"))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from fast_forward import Ranking, _ffx  # noqa: E402
from fast_forward.encoder import TableEncoder  # noqa: E402
from fast_forward.index import InMemoryIndex, Mode  # noqa: E402

n_docs = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
C = int(sys.argv[3]) if len(sys.argv) > 3 else 5000
D, alpha, cutoff = 768, 0.1, 1000
rng = np.random.default_rng(0)
cnt = bench.doc_lengths(n_docs, 6.25, seed=0)
n_rows = int(cnt.sum())
doc_ids = np.repeat(np.char.add("D", np.arange(n_docs).astype(str)), cnt).tolist()
psg_ids = np.char.add("P", np.arange(n_rows).astype(str)).tolist()
qv = rng.standard_normal((nq, D), dtype=np.float32)
index = InMemoryIndex(TableEncoder({f"text {i}": qv[i] for i in range(nq)}), mode=Mode.MAXP, init_size=n_rows)
t_add = time.perf_counter()
step = 1 << 18
for lo in range(0, n_rows, step):
    hi = min(n_rows, lo + step)
    index.add(rng.standard_normal((hi - lo, D), dtype=np.float32), doc_ids=doc_ids[lo:hi], psg_ids=psg_ids[lo:hi])
t_add = time.perf_counter() - t_add

bucket = n_docs // C
docs = np.argsort(rng.random((nq, C)), axis=1) * bucket + rng.integers(0, bucket, (nq, C))
frame = pd.DataFrame({"q_id": np.repeat(np.char.add("q", np.arange(nq).astype(str)), C),
                      "id": np.char.add("D", docs.ravel().astype(str)),
                      "score": rng.uniform(0, 20, nq * C).astype(np.float32)})
t_rank = time.perf_counter()
ranking = Ranking(frame, queries={f"q{i}": f"text {i}" for i in range(nq)})
t_rank = time.perf_counter() - t_rank

index.rerank(ranking, alpha, cutoff)  # warm: CSR upload, scratch buffers
times = {}
for name, fn in (("rerank", lambda: index.rerank(ranking, alpha, cutoff)),
                 ("call_interpolate_cut", lambda: ranking.interpolate(index(ranking), alpha).cut(cutoff)),
                 ("call_only", lambda: index(ranking))):
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        out = fn()
        best = min(best, time.perf_counter() - t0)
    times[name] = best
fused = index.rerank(ranking, alpha, cutoff)
three = ranking.interpolate(index(ranking), alpha).cut(cutoff)
if not fused == three:
    a = fused._df.sort_values(["q_id", "id"]).reset_index(drop=True)
    b = three._df.sort_values(["q_id", "id"]).reset_index(drop=True)
    m = a.merge(b, on=["q_id", "id"], how="outer", suffixes=("_f", "_t"), indicator=True)
    bad = m[(m["_merge"] != "both") | (m["score_f"] != m["score_t"])]
    print(len(a), len(b), m["_merge"].value_counts().to_dict(), len(bad), file=sys.stderr)
    print(bad.head(12), file=sys.stderr)
print(json.dumps({"docs": n_docs, "passages": n_rows, "pairs": nq * C, "cutoff": cutoff,
                  "add_s": round(t_add, 2), "ranking_ctor_s": round(t_rank, 2),
                  "seconds": {k: round(v, 3) for k, v in times.items()},
                  "pairs_per_s": {k: round(nq * C / v) for k, v in times.items()},
                  "fused_equals_three_call": bool(fused == three),
                  "kernel_launches": _ffx.launch_count()}))
if os.environ.get("FFX_PROFILE"):
    import cProfile
    import pstats

    pr = cProfile.Profile()
    pr.enable()
    index.rerank(ranking, alpha, cutoff)
    pr.disable()
    pstats.Stats(pr, stream=sys.stderr).sort_stats("cumulative").print_stats(28)
