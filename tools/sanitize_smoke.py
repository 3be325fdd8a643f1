"""Small invocations of every kernel family, for compute-sanitizer (memcheck / racecheck):
fp32 scoring (TMA-staged + register-staged, fused and tiled), early stopping, the three ADC
kernels (fused and tiled), interpolate + top-k, shard merge.  Results are checked against the
oracle by __graft_entry__.smoke() and by the asserts below."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "fast-forward-indexes_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import __graft_entry__ as g  # noqa: E402

g.smoke()
import ff_oracle as fo  # noqa: E402
from fast_forward import _ffx  # noqa: E402

rng = np.random.default_rng(0)
n_docs, dim = 400, 768
cnt = rng.integers(1, 40, n_docs)
off = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
vec = rng.standard_normal((off[-1], dim)).astype(np.float32)
idx = _ffx.DeviceIndex(dim, capacity=len(vec))
idx.stage(0, vec)
idx.set_docs(off)
for nq in (5, 330):  # tiled + separate top-k, fused
    qv = rng.standard_normal((nq, dim)).astype(np.float32)
    cnts = rng.integers(0, 120, nq)
    q_off = np.concatenate([[0], np.cumsum(cnts)]).astype(np.int64)
    cand = np.concatenate([rng.choice(n_docs, c, replace=False) for c in cnts] + [np.zeros(0, np.int64)]).astype(np.int32)
    lex = rng.uniform(0, 20, len(cand)).astype(np.float32)
    for kern in (1, 2):
        _ffx.set_option("kernel", kern)
        for mode in (1, 2, 3, 4):
            c = (cand % len(vec)) if mode == 1 else cand
            out = idx.rerank_host(mode, qv, q_off, c, lex, 0.1, 50, want_ff=True, want_int=True)
            ts, tp = fo.topk_per_query(q_off, out["int"], 50)
            assert (out["topk_pos"] == tp).all()
    _ffx.set_option("kernel", 0)
    if nq == 5:
        lex_sorted = np.concatenate([np.sort(lex[q_off[q]:q_off[q + 1]])[::-1] for q in range(nq)] + [np.zeros(0, np.float32)])
        es = idx.rerank_early_stop_host(2, qv, q_off, cand, lex_sorted, 0.5, 5, (5, 20, 60, 200))
        assert (es["scored"] <= cnts).all()
idx.close()

for M, Ks, Ds in ((96, 256, 8), (64, 32, 4), (8, 16, 4)):
    n_rows = int(off[-1])
    codes = rng.integers(0, Ks, (n_rows, M)).astype(np.uint8)
    pq = _ffx.DeviceIndex(M, capacity=n_rows, row_kind=_ffx.ROWS_PQ_U8)
    pq.stage(0, codes)
    pq.set_docs(off)
    D = M * Ds
    pq.set_pq(rng.standard_normal((M, Ks, Ds)).astype(np.float32), np.linalg.qr(rng.standard_normal((D, D)))[0].astype(np.float32))
    for nq in (4, 160):
        qv = rng.standard_normal((nq, D)).astype(np.float32)
        cnts = rng.integers(0, 150, nq)
        q_off = np.concatenate([[0], np.cumsum(cnts)]).astype(np.int64)
        cand = np.concatenate([rng.choice(n_docs, c, replace=False) for c in cnts] + [np.zeros(0, np.int64)]).astype(np.int32)
        lex = rng.uniform(0, 20, len(cand)).astype(np.float32)
        ref = None
        for adc in (1, 2, 3):
            _ffx.set_option("adc", adc)
            out = pq.rerank_host(4, qv, q_off, cand, lex, 0.1, 64, want_ff=True, want_int=True)
            ts, tp = fo.topk_per_query(q_off, out["int"], 64)
            assert (out["topk_pos"] == tp).all()
            ref = out["ff"] if ref is None else ref
            assert np.allclose(out["ff"], ref, rtol=1e-4, atol=1e-3)
        _ffx.set_option("adc", 0)
    pq.close()
print("sanitize_smoke ok")
