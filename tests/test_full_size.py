"""BASELINE.json's full sizes on the GPU, against the ORACLE (not against the kernel itself):

* C3 (3.2 M docs / ~20 M passages, 61 GB of fp32 rows in HBM, 5193 queries x 5000 candidates,
  MAXP / AVEP / FIRSTP through the fused kernel the bench times): for whole queries spread over the job
  the ~31 k rows of their 5000 candidates are read back from the store and scored by the
  plain-C oracle; semantic scores, interpolated scores and the complete ranked list are
  compared bit for bit.  Size-independent properties cover the rest of the pass.
* C2 (PASSAGE over 8.8 M passages, 6980 queries x 1000 candidates), the same way.
* C5, one doc-id-range shard of eight (12 500 queries x 5000 candidates, k = 1000).
* C4 (OPQ M = 96, Ks = 256, Ds = 8 over 20 M codes, AVEP, 5193 x 5000) against
  decode-then-dot in float64 (rtol 1e-5 + atol 1e-5 |q||d|: the ADC sum reassociates).
"""

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


# ------------------------------------------------------------------------------------------
# whole queries against the oracle
# ------------------------------------------------------------------------------------------
P = ctypes.c_void_p


def oracle_query(oracle_c, idx, off, mode, qvec, cand_units, lex, alpha, k, base=0, n_units=None):
    """One whole query through the plain-C oracle on rows READ BACK from the device store:
    semantic scores (index/base.py:279-314), interpolation (ranking.py:319) and the ranked list
    (ranking.py:115-117,285-291).  `cand_units`: local document ordinals (row numbers in PASSAGE
    mode), -1 = pair owned by another shard (not scored, not ranked)."""
    mine = cand_units >= 0
    units = cand_units[mine].astype(np.int64)
    if mode == fo.MODE_PASSAGE:
        rows = units
        u_off = np.arange(len(units) + 1, dtype=np.int64)
    elif mode == fo.MODE_FIRSTP:
        rows = off[units]
        u_off = np.arange(len(units) + 1, dtype=np.int64)
    else:
        cnt = off[units + 1] - off[units]
        u_off = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
        rows = (np.repeat(off[units] - u_off[:-1], cnt) + np.arange(u_off[-1])).astype(np.int64)
    vec = idx.read_rows(rows)  # ~31 k rows x 3 KB for a C3 query
    u_rows = np.arange(len(rows), dtype=np.int64)
    pair_q = np.zeros(len(units), np.int64)
    pair_u = np.arange(len(units), dtype=np.int64)
    ff_mine = np.empty(len(units), np.float32)
    qv = np.ascontiguousarray(qvec[None, :], np.float32)
    o_mode = fo.MODE_PASSAGE if mode in (fo.MODE_PASSAGE, fo.MODE_FIRSTP) else mode
    oracle_c.ffo_score_pairs(P(vec.ctypes.data), ctypes.c_int64(vec.shape[1]), P(u_off.ctypes.data),
                             P(u_rows.ctypes.data), P(pair_q.ctypes.data), P(pair_u.ctypes.data),
                             ctypes.c_int64(len(units)), P(qv.ctypes.data), ctypes.c_int(o_mode), P(ff_mine.ctypes.data))
    it_mine = fo.interpolate_f32(lex[mine], ff_mine, alpha)
    # rank only this shard's pairs; positions stay positions in the full candidate block
    order = np.lexsort((np.flatnonzero(mine), -it_mine.astype(np.float64)))  # score desc, ties by position
    pos = np.flatnonzero(mine)[order][:k]
    top_s = np.full(k, -np.inf, np.float32)
    top_p = np.full(k, -1, np.int32)
    top_s[:len(pos)] = it_mine[order][:k]
    top_p[:len(pos)] = pos
    return mine, ff_mine, it_mine, top_s, top_p


def draw_job(torch, dev, nq, cands, pool, seed):
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    qv = torch.randn((nq, DIM), device=dev, generator=gen)
    bucket = pool // cands
    perm = torch.rand((nq, cands), device=dev, generator=gen).argsort(dim=1)
    cand = (perm * bucket + torch.randint(0, bucket, (nq, cands), device=dev, generator=gen)).to(torch.int32)
    cand = cand.contiguous().view(-1)
    del perm
    # coarse first-stage scores: many exact ties inside a query, so the tie rule matters
    lex = (torch.randint(0, 400, (nq * cands,), device=dev, generator=gen).float() * 0.05).contiguous()
    q_off = (torch.arange(nq + 1, device=dev, dtype=torch.int64) * cands).contiguous()
    return qv, cand, lex, q_off


def device_pass(big, mode, qv, cand, lex, q_off, nq, cands, k, alpha, zero_outputs=False):
    torch, idx = big["torch"], big["idx"]
    dev = qv.device
    ff = torch.zeros(nq * cands, device=dev)
    it = torch.zeros(nq * cands, device=dev)
    ts = torch.empty((nq, k), device=dev)
    tp = torch.empty((nq, k), device=dev, dtype=torch.int32)
    torch.cuda.synchronize()
    idx.rerank_device(mode, qv.data_ptr(), nq, q_off.data_ptr(), cand.data_ptr(), lex.data_ptr(), alpha,
                      k, cands, ff.data_ptr(), it.data_ptr(), ts.data_ptr(), tp.data_ptr())
    idx.sync()
    return ff, it, ts, tp


WHOLE_QUERIES = 5  # per mode; first, last and three in between


def pick(nq):
    return sorted({0, nq - 1, nq // 2, nq // 3, (2 * nq) // 3})[:WHOLE_QUERIES]


@pytest.mark.parametrize("mode", ["MAXP", "AVEP", "FIRSTP"])
def test_c3_whole_queries_against_the_oracle(big, oracle_c, mode):
    """The bench's pass (5193 x 5000 over the full store, fused ffx_score_tma_kernel<2,12,true,32>)
    and, for five whole queries, every semantic score, every interpolated score and the full
    5000-entry ranked list against the plain-C oracle on read-back rows: bit for bit."""
    torch, ffx = big["torch"], big["ffx"]
    dev = torch.device("cuda", 0)
    nq, alpha = 5193, 0.1
    m = getattr(fo, "MODE_" + mode)
    qv, cand, lex, q_off = draw_job(torch, dev, nq, CANDS, big["n_docs"], 11)
    ff, it, ts, tp = device_pass(big, m, qv, cand, lex, q_off, nq, CANDS, CANDS, alpha)
    assert "ffx_score_tma_kernel<2, 12, true, 32>" in ffx.last_kernel()
    for q in pick(nq):
        sl = slice(q * CANDS, (q + 1) * CANDS)
        c_host = cand[sl].cpu().numpy().astype(np.int64)
        l_host = lex[sl].cpu().numpy()
        _, o_ff, o_it, o_ts, o_tp = oracle_query(oracle_c, big["idx"], big["off"], m, qv[q].cpu().numpy(), c_host,
                                                 l_host, alpha, CANDS)
        assert (bits(ff[sl].cpu().numpy()) == bits(o_ff)).all(), (mode, q)
        assert (bits(it[sl].cpu().numpy()) == bits(o_it)).all(), (mode, q)
        assert (tp[q].cpu().numpy() == o_tp).all(), (mode, q)
        assert (bits(ts[q].cpu().numpy()) == bits(o_ts)).all(), (mode, q)


def test_c2_whole_queries_against_the_oracle(big, oracle_c):
    """BASELINE configs[1]: PASSAGE over 8.8 M passages (the first 8.8 M rows of the store),
    6980 queries x 1000 candidates, k = 1000; five whole queries against the oracle."""
    torch = big["torch"]
    dev = torch.device("cuda", 0)
    nq, cands, alpha = 6980, 1000, 0.1
    pool = min(8_800_000, big["n_rows"])
    qv, cand, lex, q_off = draw_job(torch, dev, nq, cands, pool, 12)
    ff, it, ts, tp = device_pass(big, fo.MODE_PASSAGE, qv, cand, lex, q_off, nq, cands, cands, alpha)
    for q in pick(nq):
        sl = slice(q * cands, (q + 1) * cands)
        _, o_ff, o_it, o_ts, o_tp = oracle_query(oracle_c, big["idx"], big["off"], fo.MODE_PASSAGE,
                                                 qv[q].cpu().numpy(), cand[sl].cpu().numpy().astype(np.int64),
                                                 lex[sl].cpu().numpy(), alpha, cands)
        assert (bits(ff[sl].cpu().numpy()) == bits(o_ff)).all(), q
        assert (bits(it[sl].cpu().numpy()) == bits(o_it)).all(), q
        assert (tp[q].cpu().numpy() == o_tp).all() and (bits(ts[q].cpu().numpy()) == bits(o_ts)).all(), q


def test_c5_shard_whole_queries_against_the_oracle(big, oracle_c):
    """BASELINE configs[4], one rank's view: the store is doc-id-range shard 3 of 8 of an 8x larger
    corpus; 12 500 queries x 5000 candidates (global ordinals over the whole corpus), k = 1000.
    The shard scores only its own pairs and ranks them at their positions in the full block;
    five whole queries against the oracle (foreign pairs: outputs untouched, no list slot)."""
    torch, idx = big["torch"], big["idx"]
    dev = torch.device("cuda", 0)
    nq, k, alpha, shards, mine_shard = 12_500, 1000, 0.1, 8, 3
    n_local = big["n_docs"]
    qv, cand, lex, q_off = draw_job(torch, dev, nq, CANDS, n_local * shards, 13)
    idx.set_shard(mine_shard * n_local, shards * n_local, mine_shard * big["n_rows"], shards * big["n_rows"])
    try:
        ff, it, ts, tp = device_pass(big, fo.MODE_MAXP, qv, cand, lex, q_off, nq, CANDS, k, alpha)
    finally:
        idx.set_shard(0, 0, 0, 0)
    for q in pick(nq):
        sl = slice(q * CANDS, (q + 1) * CANDS)
        c_glob = cand[sl].cpu().numpy().astype(np.int64)
        local = np.where((c_glob >= mine_shard * n_local) & (c_glob < (mine_shard + 1) * n_local),
                         c_glob - mine_shard * n_local, -1)
        mine, o_ff, o_it, o_ts, o_tp = oracle_query(oracle_c, idx, big["off"], fo.MODE_MAXP, qv[q].cpu().numpy(),
                                                    local, lex[sl].cpu().numpy(), alpha, k)
        assert 400 < mine.sum() < 900  # ~1/8 of the block
        g_ff, g_it = ff[sl].cpu().numpy(), it[sl].cpu().numpy()
        assert (bits(g_ff[mine]) == bits(o_ff)).all() and (bits(g_it[mine]) == bits(o_it)).all(), q
        assert (g_ff[~mine] == 0).all() and (g_it[~mine] == 0).all()  # other shards' pairs: untouched
        assert (tp[q].cpu().numpy() == o_tp).all() and (bits(ts[q].cpu().numpy()) == bits(o_ts)).all(), q


def test_c4_shape_whole_queries_against_decode_then_dot(big):
    """BASELINE configs[3]: OPQ M = 96, Ks = 256, Ds = 8 over ~20 M codes with the C3 document
    structure, AVEP, 5193 queries x 5000 candidates through the fused ADC kernel.  Five whole
    queries against decode-then-dot in float64 (quantizer/nanopq.py:43-44,111-112 then
    index/base.py:296-314): |ADC - reference| <= 1e-5 |reference| + 1e-5 |q| |d| per pair (the ADC
    sum reassociates the D products into M table entries); interpolation bit-exact on the
    kernel's own scores; the ranked list = the stable descending order of those."""
    torch, ffx = big["torch"], big["ffx"]
    dev = torch.device("cuda", 0)
    M, Ks, Ds = 96, 256, 8
    n_rows, off = big["n_rows"], big["off"]
    pq = ffx.DeviceIndex(M, capacity=n_rows, row_kind=ffx.ROWS_PQ_U8)
    gen = torch.Generator(device=dev)
    gen.manual_seed(77)
    for r0 in range(0, n_rows, 1 << 22):
        nr = min(1 << 22, n_rows - r0)
        t = torch.randint(0, Ks, (nr, M), device=dev, dtype=torch.uint8, generator=gen)
        torch.cuda.synchronize()
        pq.stage_device(r0, nr, t.data_ptr())
        del t
    pq.set_docs(off)
    rng = np.random.default_rng(5)
    cw = rng.standard_normal((M, Ks, Ds)).astype(np.float32)
    R = np.linalg.qr(rng.standard_normal((DIM, DIM)))[0].astype(np.float32)
    pq.set_pq(cw, R)
    nq, alpha = 5193, 0.1
    qv, cand, lex, q_off = draw_job(torch, dev, nq, CANDS, big["n_docs"], 14)
    ff = torch.zeros(nq * CANDS, device=dev)
    it = torch.zeros(nq * CANDS, device=dev)
    ts = torch.empty((nq, CANDS), device=dev)
    tp = torch.empty((nq, CANDS), device=dev, dtype=torch.int32)
    torch.cuda.synchronize()
    pq.rerank_device(fo.MODE_AVEP, qv.data_ptr(), nq, q_off.data_ptr(), cand.data_ptr(), lex.data_ptr(), alpha,
                     CANDS, CANDS, ff.data_ptr(), it.data_ptr(), ts.data_ptr(), tp.data_ptr())
    pq.sync()
    assert "ffx_adc_xor_kernel<3, true>" in ffx.last_kernel()
    try:
        for q in pick(nq):
            sl = slice(q * CANDS, (q + 1) * CANDS)
            units = cand[sl].cpu().numpy().astype(np.int64)
            cnt = off[units + 1] - off[units]
            u_off = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
            rows = (np.repeat(off[units] - u_off[:-1], cnt) + np.arange(u_off[-1])).astype(np.int64)
            codes = pq.read_rows(rows)
            dec = fo.pq_decode(codes, cw).astype(np.float64) @ R.astype(np.float64).T  # nanopq OPQ.decode
            q64 = qv[q].cpu().numpy().astype(np.float64)
            line = dec @ q64
            want = np.add.reduceat(line, u_off[:-1]) / cnt
            scale = np.maximum.reduceat(np.linalg.norm(dec, axis=1), u_off[:-1]) * np.linalg.norm(q64)
            got = ff[sl].cpu().numpy()
            assert (np.abs(got - want) <= 1e-5 * np.abs(want) + 1e-5 * scale).all(), q
            l_host = lex[sl].cpu().numpy()
            o_it = fo.interpolate_f32(l_host, got, alpha)
            assert (bits(it[sl].cpu().numpy()) == bits(o_it)).all(), q
            o_ts, o_tp = fo.topk_per_query(np.array([0, CANDS]), o_it, CANDS)
            assert (tp[q].cpu().numpy() == o_tp[0]).all() and (bits(ts[q].cpu().numpy()) == bits(o_ts[0])).all(), q
    finally:
        pq.close()
