"""BASELINE.json's full C3 size (3.2 M docs / ~20 M passages, 61 GB of fp32 rows in HBM,
5193 queries x 5000 candidates) checked through size-independent properties, plus oracle
spot checks on rows read back from the far end of the store (64-bit addressing)."""

import ctypes
import os
import sys

import numpy as np
import pytest

import ff_oracle as fo
from conftest import ROOT

pytestmark = pytest.mark.gpu

DIM, CANDS = 768, 5000


@pytest.fixture(scope="module")
def big():
    import torch

    import __graft_entry__ as g

    g.build()
    from fast_forward import _ffx

    free, _ = torch.cuda.mem_get_info()
    scale = 1.0 if free > 90e9 else 0.1  # a smaller GPU still runs the properties
    sys.path.insert(0, ROOT)
    import bench

    n_docs = int(3_200_000 * scale)
    cnt = bench.doc_lengths(n_docs, 6.25, seed=0)
    off = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
    n_rows = int(off[-1])
    dev = torch.device("cuda", 0)
    idx = _ffx.DeviceIndex(DIM, capacity=n_rows)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234)
    for r0 in range(0, n_rows, 1 << 20):
        nr = min(1 << 20, n_rows - r0)
        t = torch.randn((nr, DIM), device=dev, generator=gen)
        torch.cuda.synchronize()
        idx.stage_device(r0, nr, t.data_ptr())
        del t
    idx.set_docs(off)
    torch.cuda.empty_cache()
    yield {"idx": idx, "off": off, "cnt": cnt, "n_docs": n_docs, "n_rows": n_rows, "ffx": _ffx, "torch": torch}
    idx.close()


def bits(x):
    return np.ascontiguousarray(x, np.float32).view(np.uint32)


def test_modes_are_consistent_with_passage_scores(big, oracle_c):
    """MAXP == max, AVEP == Kahan mean (N2), FIRSTP == first of the PASSAGE scores of the
    document's rows, bit for bit; and a handful of pairs against the plain-C oracle on rows
    read back from both ends of the store."""
    idx, off, cnt = big["idx"], big["off"], big["cnt"]
    rng = np.random.default_rng(0)
    nq, C = 300, 400
    qv = rng.standard_normal((nq, DIM)).astype(np.float32)
    docs = np.concatenate([rng.choice(big["n_docs"], C, replace=False) for _ in range(nq)])
    docs[:4] = [0, big["n_docs"] - 1, big["n_docs"] - 2, 1]  # both ends of the 61 GB store
    docs = docs.astype(np.int32)
    q_off = np.arange(nq + 1, dtype=np.int64) * C
    rows = np.concatenate([np.arange(off[d], off[d + 1]) for d in docs]).astype(np.int32)
    per_pair = cnt[docs]
    seg = np.concatenate([[0], np.cumsum(per_pair)])
    row_q_off = seg[np.arange(nq + 1) * C]
    psg = idx.rerank_host(fo.MODE_PASSAGE, qv, row_q_off, rows)["ff"]
    maxp = idx.rerank_host(fo.MODE_MAXP, qv, q_off, docs)["ff"]
    avep = idx.rerank_host(fo.MODE_AVEP, qv, q_off, docs)["ff"]
    first = idx.rerank_host(fo.MODE_FIRSTP, qv, q_off, docs)["ff"]
    assert (bits(maxp) == bits(np.maximum.reduceat(psg, seg[:-1]))).all()
    assert (bits(first) == bits(psg[seg[:-1]])).all()
    assert (bits(avep) == bits(fo.kahan_mean_f32(psg, seg))).all()
    assert (first <= maxp).all() and (avep <= maxp).all()

    # oracle on read-back rows
    some = np.array([0, 1, big["n_rows"] - 1, big["n_rows"] - 2, big["n_rows"] // 2], np.int64)
    vec = idx.read_rows(some)
    out = np.empty(len(some), np.float32)
    P = ctypes.c_void_p
    u_off = np.arange(len(some) + 1, dtype=np.int64)
    u_rows = np.arange(len(some), dtype=np.int64)
    pq = np.zeros(len(some), np.int64)
    oracle_c.ffo_score_pairs(P(vec.ctypes.data), ctypes.c_int64(DIM), P(u_off.ctypes.data), P(u_rows.ctypes.data),
                             P(pq.ctypes.data), P(u_rows.ctypes.data), ctypes.c_int64(len(some)),
                             P(qv.ctypes.data), ctypes.c_int(fo.MODE_PASSAGE), P(out.ctypes.data))
    got = idx.rerank_host(fo.MODE_PASSAGE, qv[:1], [0, len(some)], some.astype(np.int32))["ff"]
    assert (bits(got) == bits(out)).all()


def test_full_workload_ranked_lists(big):
    """The bench's full pass (5193 x 5000, MAXP, alpha 0.1, k = 5000): every list is sorted,
    is a permutation of the query's block, equals interpolate(lex, ff) gathered at its
    positions, and a second pass reproduces it bit for bit."""
    torch, idx = big["torch"], big["idx"]
    dev = torch.device("cuda", 0)
    nq = 5193
    gen = torch.Generator(device=dev)
    gen.manual_seed(5)
    qv = torch.randn((nq, DIM), device=dev, generator=gen)
    bucket = big["n_docs"] // CANDS
    perm = torch.rand((nq, CANDS), device=dev, generator=gen).argsort(dim=1)
    cand = (perm * bucket + torch.randint(0, bucket, (nq, CANDS), device=dev, generator=gen)).to(torch.int32)
    cand = cand.contiguous().view(-1)
    del perm
    lex = (torch.rand((nq * CANDS,), device=dev, generator=gen) * 20).contiguous()
    q_off = (torch.arange(nq + 1, device=dev, dtype=torch.int64) * CANDS).contiguous()
    outs = []
    for _ in range(2):
        ff = torch.zeros(nq * CANDS, device=dev)
        it = torch.zeros(nq * CANDS, device=dev)
        ts = torch.empty((nq, CANDS), device=dev)
        tp = torch.empty((nq, CANDS), device=dev, dtype=torch.int32)
        torch.cuda.synchronize()
        idx.rerank_device(fo.MODE_MAXP, qv.data_ptr(), nq, q_off.data_ptr(), cand.data_ptr(), lex.data_ptr(), 0.1,
                          CANDS, CANDS, ff.data_ptr(), it.data_ptr(), ts.data_ptr(), tp.data_ptr())
        idx.sync()
        outs.append((ff, it, ts, tp))
    ff, it, ts, tp = outs[0]
    for a, b in zip(outs[0], outs[1]):
        assert torch.equal(a, b)
    assert bool((ts[:, 1:] <= ts[:, :-1]).all())
    assert bool((tp.long().sort(dim=1).values == torch.arange(CANDS, device=dev)).all())
    assert torch.equal(it.view(nq, CANDS).gather(1, tp.long()), ts)
    a32, b32 = np.float32(0.1), np.float32(1 - 0.1)
    want_it = (lex * float(a32)) + (ff * float(b32))  # same two roundings in fp32 on the device
    assert torch.equal(want_it, it)
    # ties by position: equal neighbours must be in ascending position order
    eq = ts[:, 1:] == ts[:, :-1]
    assert bool((tp[:, 1:][eq] > tp[:, :-1][eq]).all())
