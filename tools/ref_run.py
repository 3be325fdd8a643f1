"""TEST INFRASTRUCTURE — run the UNMODIFIED reference package (baseline/_ref) on an input bundle.

    python tools/ref_run.py <in.npz> <out.json>

in.npz: vectors [n, D] f32, doc_ids [n] str, qvecs [nq, D] f32, q_id / id / score columns of a
first-stage ranking, alpha, cutoff, es_cutoff, es_alpha, es_depths.  Runs, per Mode,
`index(ranking)`, `ranking.interpolate(out, alpha)`, `.cut(cutoff)` and
`index(ranking, early_stopping=...)` and writes every resulting frame (ids + float32 bit
patterns) to out.json.  Own interpreter: the reference shares its import name with the drop-in."""
import json
import os
import sys

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (only for the import helper of the reference)

ff = bench._import_reference()
from fast_forward.encoder import LambdaEncoder  # noqa: E402
from fast_forward.index import InMemoryIndex, Mode  # noqa: E402

data = np.load(sys.argv[1], allow_pickle=False)
qvecs = data["qvecs"]
queries = {f"q{i}": f"text {i}" for i in range(len(qvecs))}
table = {f"text {i}": qvecs[i] for i in range(len(qvecs))}
index = InMemoryIndex(LambdaEncoder(lambda q: table[q]), init_size=len(data["vectors"]))
split = int(data["split"]) if "split" in data else len(data["vectors"])
psg_ids = [f"p{i}" for i in range(len(data["vectors"]))]
# two add() calls: rows after `split` extend earlier documents (their rows are not contiguous)
index.add(data["vectors"][:split], doc_ids=data["doc_ids"][:split].tolist(), psg_ids=psg_ids[:split])
if split < len(psg_ids):
    index.add(data["vectors"][split:], doc_ids=data["doc_ids"][split:].tolist(), psg_ids=psg_ids[split:])
score_dtype = np.dtype(str(data["score_dtype"])) if "score_dtype" in data else np.dtype(np.float32)
batch_size = int(data["batch_size"]) if "batch_size" in data and int(data["batch_size"]) > 0 else None


def frame(r):
    df = r._df
    return {"q_id": df["q_id"].tolist(), "id": df["id"].tolist(), "dtype": str(df["score"].dtype),
            "score_bits": df["score"].to_numpy().astype(np.float64).view(np.uint64).tolist()}


out = {}
alpha, cutoff = float(data["alpha"]), int(data["cutoff"])
for mode in (Mode.MAXP, Mode.AVEP, Mode.FIRSTP, Mode.PASSAGE):
    key = "psg" if mode == Mode.PASSAGE else "doc"
    first = ff.Ranking(pd.DataFrame({"q_id": data[f"{key}_q_id"], "id": data[f"{key}_id"], "score": data[f"{key}_score"]}),
                       queries=queries, dtype=score_dtype)
    index.mode = mode
    scored = index(first, batch_size=batch_size)
    inter = first.interpolate(scored, alpha)
    es = index(first, early_stopping=int(data["es_cutoff"]), early_stopping_alpha=float(data["es_alpha"]),
               early_stopping_depths=tuple(int(d) for d in data["es_depths"]))
    out[mode.name] = {"first": frame(first), "ff": frame(scored), "interpolated": frame(inter),
                      "cut": frame(inter.cut(cutoff)), "early_stopping": frame(es)}
with open(sys.argv[2], "w") as f:
    json.dump(out, f)
