"""Pin the oracle (numpy restatement + plain-C restatement) against the golden vectors
written by the UNMODIFIED reference (oracle/gen_golden.py) — bit-exact."""

import ctypes

import numpy as np
import pytest

import ff_oracle as fo

MODES = {"PASSAGE": fo.MODE_PASSAGE, "MAXP": fo.MODE_MAXP, "FIRSTP": fo.MODE_FIRSTP,
         "AVEP": fo.MODE_AVEP}


def bits(x):
    return np.asarray(x, dtype=np.float32).view(np.uint32)


def build_kat_indexes():
    DOC = ["d0", "d0", "d1", "d2", "d3"]
    PSG = ["p0", "p1", "p2", "p3", "p4"]
    V = np.tril(np.ones((5, 5), dtype=np.int64))
    full = fo.OracleIndex()
    full.add(V, doc_ids=DOC, psg_ids=PSG)
    partial = fo.OracleIndex()  # reference tests/test_index.py:58-69
    partial.add(V, doc_ids=[None, None] + DOC[2:], psg_ids=PSG[:-2] + [None, None])
    partial.add(V[:2], doc_ids=DOC[:2])
    partial.add(V[-2:], psg_ids=PSG[-2:])
    return {"full": full, "partial": partial}


@pytest.mark.parametrize("which", ["full", "partial"])
@pytest.mark.parametrize("mode", ["MAXP", "FIRSTP", "AVEP", "PASSAGE"])
def test_known_answers(golden_kat, which, mode):
    """reference tests/test_index.py:135-200 (test_maxp/firstp/avep/passage)."""
    index = build_kat_indexes()[which]
    inp = golden_kat["input/psg_ranking" if mode == "PASSAGE" else "input/doc_ranking"]
    want = golden_kat[f"{which}/{mode}"]
    qv = {q: np.ones(5, dtype=np.int64) for q in set(inp["q_id"])}
    ff = fo.call_index(index, inp["q_id"], inp["id"], qv, MODES[mode])
    order = fo.sort_ranking(inp["q_id"], inp["id"], ff.astype(np.float32))
    assert [inp["q_id"][i] for i in order] == want["q_id"]
    assert [inp["id"][i] for i in order] == want["id"]
    assert (bits(ff[order]) == np.array(want["score_bits"], dtype=np.uint32)).all()
    # the hand-written expectations of the reference tests themselves
    d0 = {"MAXP": 2.0, "FIRSTP": 1.0, "AVEP": 1.5}
    if mode in d0:
        got = {(q, i): s for q, i, s in zip(want["q_id"], want["id"], want["score"])}
        assert got[("q1", "d0")] == d0[mode] and got[("q2", "d3")] == 5.0


def test_missing_id_raises():
    """reference tests/test_index.py:266-271, index/util.py:38-39."""
    index = build_kat_indexes()["full"]
    with pytest.raises(IndexError):
        fo.call_index(index, ["q1", "q1"], ["d0", "dx"], {"q1": np.ones(5)}, fo.MODE_MAXP)
    with pytest.raises(IndexError):
        index.rows_for("d0", fo.MODE_PASSAGE)


def test_ranking_algebra(golden_kat):
    """reference tests/test_ranking.py:116-121 (cut), :157-188 (interpolate)."""
    base = golden_kat["rank/base"]
    q, i, s = fo.cut(base["q_id"], base["id"], np.array(base["score"], np.float32), 2)
    assert (q, i) == (golden_kat["rank/cut2"]["q_id"], golden_kat["rank/cut2"]["id"])
    other = golden_kat["rank/interp_other"]
    a = {(x, y): z for x, y, z in zip(base["q_id"], base["id"], base["score"])}
    b = {(x, y): z for x, y, z in zip(other["q_id"], other["id"], other["score"])}
    q, i, s = fo.interpolate_rankings(a, b, 0.5)
    want = golden_kat["rank/interp0.5"]
    assert (q, i) == (want["q_id"], want["id"])
    assert (bits(s) == np.array(want["score_bits"], np.uint32)).all()
    r4, r5 = golden_kat["rank/r4"], golden_kat["rank/r5"]
    a = {(x, y): z for x, y, z in zip(r4["q_id"], r4["id"], r4["score"])}
    b = {(x, y): z for x, y, z in zip(r5["q_id"], r5["id"], r5["score"])}
    q, i, s = fo.interpolate_rankings(a, b, 0.5)
    want = golden_kat["rank/r4_interp_r5_0.5"]
    assert (q, i, s.tolist()) == (want["q_id"], want["id"], want["score"])


def test_sort_semantics(golden_kat):
    """ranking.py:115-117: q_id DESC as strings, score DESC, ties keep frame order."""
    inp, want = golden_kat["rank/ties_input"], golden_kat["rank/ties_sorted"]
    order = fo.sort_ranking(inp["q_id"], inp["id"], np.array(inp["score"], np.float32))
    assert [inp["q_id"][k] for k in order] == want["q_id"]
    assert [inp["id"][k] for k in order] == want["id"]


def _random_index(case, arrays, key):
    idx = fo.OracleIndex()
    v = arrays[f"{key}/vectors"]
    sp = case["split"]
    idx.add(v[:sp], doc_ids=case["doc_ids"][:sp], psg_ids=case["psg_ids"][:sp])
    idx.add(v[sp:], doc_ids=case["doc_ids"][sp:], psg_ids=case["psg_ids"][sp:])
    return idx


@pytest.mark.parametrize("key", ["s11", "s12", "s13", "s14"])
@pytest.mark.parametrize("mode", ["MAXP", "AVEP", "FIRSTP", "PASSAGE"])
def test_random_cases_bit_exact(golden_random, oracle_c, key, mode):
    meta, arrays = golden_random
    case = meta[key]
    g = case["modes"][mode]
    index = _random_index(case, arrays, key)
    qvecs = arrays[f"{key}/qvecs"]
    qv = {q: qvecs[int(q[1:])] for q in case["queries"]}
    fs = g["first_stage"]
    ff = fo.call_index(index, fs["q_id"], fs["id"], qv, MODES[mode])
    assert ff.dtype == np.float32
    order = fo.sort_ranking(fs["q_id"], fs["id"], ff)
    assert [fs["q_id"][k] for k in order] == g["ff"]["q_id"]
    assert [fs["id"][k] for k in order] == g["ff"]["id"]
    assert (bits(ff[order]) == np.array(g["ff"]["score_bits"], np.uint32)).all()

    # interpolate + cut on top (ranking.py:293-326, :279-291)
    a = {(x, y): np.float32(z) for x, y, z in zip(fs["q_id"], fs["id"], fs["score"])}
    b = {(x, y): z for x, y, z in zip(fs["q_id"], fs["id"], ff)}
    q, i, s = fo.interpolate_rankings(a, b, case["alpha"])
    assert (q, i) == (g["interpolated"]["q_id"], g["interpolated"]["id"])
    assert (bits(s) == np.array(g["interpolated"]["score_bits"], np.uint32)).all()
    q, i, s = fo.cut(q, i, s, case["cutoff"])
    assert (q, i) == (g["cut"]["q_id"], g["cut"]["id"])
    assert (bits(s) == np.array(g["cut"]["score_bits"], np.uint32)).all()

    # plain-C restatement on the same integer-coded problem
    uniq = list(dict.fromkeys(fs["id"]))
    uo, ur = index.csr(uniq, MODES[mode])
    qn = {x: k for k, x in enumerate(dict.fromkeys(fs["q_id"]))}
    pq = np.array([qn[x] for x in fs["q_id"]], np.int64)
    pu = np.array([uniq.index(x) for x in fs["id"]], np.int64)
    qmat = np.ascontiguousarray(np.stack([qv[x] for x in qn]), np.float32)
    vec = np.ascontiguousarray(index.vectors, np.float32)
    out = np.empty(len(pq), np.float32)
    P = ctypes.c_void_p
    oracle_c.ffo_score_pairs(P(vec.ctypes.data), ctypes.c_int64(vec.shape[1]), P(uo.ctypes.data),
                             P(ur.ctypes.data), P(pq.ctypes.data), P(pu.ctypes.data),
                             ctypes.c_int64(len(pq)), P(qmat.ctypes.data), ctypes.c_int(MODES[mode]),
                             P(out.ctypes.data))
    assert (bits(out) == bits(ff)).all()


def test_c_interpolate_and_topk(oracle_c):
    rng = np.random.default_rng(5)
    n = 4000
    s = (rng.integers(0, 50, n) / 4).astype(np.float32)
    f = rng.standard_normal(n).astype(np.float32) * 20
    P = ctypes.c_void_p
    for alpha in (0.1, 0.5, 0.0, 1.0, 0.3333):
        out = np.empty(n, np.float32)
        oracle_c.ffo_interpolate(P(s.ctypes.data), P(f.ctypes.data), ctypes.c_int64(n),
                                 ctypes.c_double(alpha), P(out.ctypes.data))
        assert (bits(out) == bits(fo.interpolate_f32(s, f, alpha))).all()
    q_off = np.array([0, 1000, 1000, 1007, 4000], np.int64)
    sc = np.round(f).astype(np.float32)  # many ties
    ws, wp = fo.topk_per_query(q_off, sc, 16)
    gs = np.empty((4, 16), np.float32)
    gp = np.empty((4, 16), np.int32)
    oracle_c.ffo_topk(P(q_off.ctypes.data), ctypes.c_int64(4), P(sc.ctypes.data), ctypes.c_int64(16),
                      P(gs.ctypes.data), P(gp.ctypes.data))
    assert (gp == wp).all() and (bits(gs) == bits(ws)).all()


def test_kahan_mean_matches_pandas():
    """N2: the restated fp32 Kahan mean equals pandas' groupby mean bit-for-bit."""
    import pandas as pd

    rng = np.random.default_rng(9)
    cnt = rng.integers(1, 40, 200)
    off = np.concatenate([[0], np.cumsum(cnt)])
    v = (rng.standard_normal(off[-1]) * 10 ** rng.uniform(-2, 3, off[-1])).astype(np.float32)
    lab = np.repeat(np.arange(200), cnt)
    want = pd.DataFrame({"g": lab, "v": v}).groupby("g")["v"].mean().to_numpy()
    assert want.dtype == np.float32
    assert (bits(fo.kahan_mean_f32(v, off)) == bits(want)).all()


def test_pq_port_consistency():
    """nanopq restatement: decode(encode(x)) picks the nearest codeword per subspace and
    ADC over decode equals the dot with the decoded vector (algebraic identity)."""
    import nanopq_port as npq

    rng = np.random.default_rng(3)
    x = rng.standard_normal((600, 32)).astype(np.float32)
    pq = npq.PQ(M=4, Ks=16, metric="dot", verbose=False).fit(x, iter=5)
    codes = pq.encode(x[:50])
    assert codes.dtype == np.uint8 and codes.shape == (50, 4)
    dec = pq.decode(codes)
    assert (dec == fo.pq_decode(codes, pq.codewords)).all()
    for m in range(4):
        sub = x[:50, m * 8:(m + 1) * 8]
        d = ((sub[:, None, :] - pq.codewords[m][None]) ** 2).sum(-1)
        assert (d.argmin(1) == codes[:, m]).all()
    opq = npq.OPQ(M=4, Ks=16, metric="dot", verbose=False).fit(x, pq_iter=5, rotation_iter=3)
    assert np.allclose(opq.R @ opq.R.T, np.eye(32), atol=1e-4)
    c2 = opq.encode(x[:20])
    q = rng.standard_normal(32).astype(np.float32)
    lut = np.einsum("mkd,md->mk", opq.codewords, (q @ opq.R).reshape(4, 8))
    adc = lut[np.arange(4)[None, :], c2].sum(1)
    assert np.allclose(adc, opq.decode(c2) @ q, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("key", ["es21", "es22"])
def test_early_stopping_golden(golden_es, key):
    """`Index.__call__(early_stopping=...)` (index/base.py:316-387): rows scored per query and
    their scores, against the frames the unmodified reference produced."""
    meta, arrays = golden_es
    case = meta[key]
    vec, qvecs = arrays[f"{key}/vectors"], arrays[f"{key}/qvecs"]
    index = fo.OracleIndex()
    index.add(vec, doc_ids=case["doc_ids"], psg_ids=case["psg_ids"])
    qv = {f"q{i}": qvecs[i] for i in range(len(qvecs))}
    for mode_name, entry in case["modes"].items():
        fs = entry["first_stage"]
        ff = fo.call_index(index, fs["q_id"], fs["id"], qv, MODES[mode_name]).astype(np.float32)
        # the first-stage frame is sorted (q_id DESC, score DESC): blocks are in rank order
        q_ids = np.array(fs["q_id"])
        starts = np.flatnonzero(np.concatenate([[True], q_ids[1:] != q_ids[:-1]]))
        q_off = np.concatenate([starts, [len(q_ids)]])
        lex = np.array(fs["score"], np.float32)
        for st in entry["settings"]:
            done = fo.early_stopping_depth(q_off, lex, ff, st["alpha"], st["cutoff"], st["depths"])
            got = {}
            for qi, b in enumerate(q_off[:-1]):
                for r in range(b, b + done[qi]):
                    got[(fs["q_id"][r], fs["id"][r])] = int(bits(ff[r:r + 1])[0])
            want = {(q, i): b for q, i, b in zip(st["out"]["q_id"], st["out"]["id"], st["out"]["score_bits"])}
            assert got == want, (mode_name, st["cutoff"], st["alpha"], st["depths"])
