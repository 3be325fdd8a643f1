"""Debug: which top-k configurations disagree with the oracle ordering of the kernel's own scores."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fast-forward-indexes_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import ff_oracle as fo
from fast_forward import _ffx as ffx
rng = np.random.default_rng(0)
def check(tag, idx, mode, qv, q_off, cand, lex, alpha, k):
    out = idx.rerank_host(mode, qv, q_off, cand, lex, alpha, k, want_ff=True, want_int=True)
    ts, tp = fo.topk_per_query(q_off, out["int"], k)
    badq = np.flatnonzero((out["topk_pos"] != tp).any(axis=1))
    out2 = idx.rerank_host(mode, qv, q_off, cand, lex, alpha, k, want_ff=False, want_int=False)
    badq2 = np.flatnonzero((out2["topk_pos"] != tp).any(axis=1))
    print(tag, "k", k, "alpha", alpha, "bad queries (with outputs)", len(badq), "(topk only)", len(badq2), "of", len(q_off) - 1, flush=True)
    if len(badq2):
        q = badq2[0]; n = q_off[q+1]-q_off[q]
        got = out2["topk_pos"][q]; want = tp[q]
        d = np.flatnonzero(got != want)
        print("   first bad query", q, "n", n, "first diff at rank", d[0], "got", got[d[0]:d[0]+6], "want", want[d[0]:d[0]+6], "sorted-set-equal", set(got[:n]) == set(want[:n]))
def pairs(nq, pool, lo, hi):
    sizes = rng.integers(lo, hi + 1, nq); q_off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    cand = np.concatenate([rng.choice(pool, s, replace=False) for s in sizes]).astype(np.int32)
    return q_off, cand
# fp32
D = 384; n_docs = 13000
cnt = rng.integers(1, 3, n_docs); off = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
vec = rng.standard_normal((off[-1], D)).astype(np.float32)
idx = ffx.DeviceIndex(D, capacity=len(vec)); idx.stage(0, vec); idx.set_docs(off)
for kern in (1, 2):
    ffx.set_option("kernel", kern)
    for lo, hi in ((2048, 2400), (4700, 5000)):
        nq = 300; qv = rng.standard_normal((nq, D)).astype(np.float32); q_off, cand = pairs(nq, n_docs, lo, hi)
        lex = rng.standard_normal(len(cand)).astype(np.float32)
        check(f"fp32 kernel={kern} lists {lo}-{hi}", idx, fo.MODE_MAXP, qv, q_off, cand, lex, 0.3, int(np.diff(q_off).max()))
ffx.set_option("kernel", 0)
idx.close()
# ADC
for M, Ks, Ds, mode in ((64, 16, 2, fo.MODE_MAXP), (96, 256, 8, fo.MODE_AVEP), (96, 256, 8, fo.MODE_MAXP)):
    n_docs = 6000; cnt = rng.integers(1, 4, n_docs); off = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
    idx = ffx.DeviceIndex(M, capacity=int(off[-1]), row_kind=ffx.ROWS_PQ_U8)
    idx.stage(0, rng.integers(0, Ks, (int(off[-1]), M)).astype(np.uint8)); idx.set_docs(off)
    idx.set_pq(rng.standard_normal((M, Ks, Ds)).astype(np.float32), None)
    for nq in (150, 400):
        qv = rng.standard_normal((nq, M * Ds)).astype(np.float32); q_off, cand = pairs(nq, n_docs, 4200, 5000)
        lex = rng.standard_normal(len(cand)).astype(np.float32)
        check(f"adc M={M} Ks={Ks} mode={mode} nq={nq}", idx, mode, qv, q_off, cand, lex, 0.3, int(np.diff(q_off).max()))
    idx.close()
