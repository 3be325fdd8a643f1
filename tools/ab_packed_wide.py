"""A/B: D = 384 / 512 through ffx_score_packed_kernel (the default: two rows per warp step) against the
TMA-staged whole-warp kernel (option kernel = 2) — outputs must be identical bit for bit."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fast-forward-indexes_b200")); sys.path.insert(0, ROOT)
from fast_forward import _ffx

rng = np.random.default_rng(5)
for dim in (384, 512):
    n_docs = 3000
    cnt = rng.integers(1, 12, n_docs)
    off = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
    vec = rng.standard_normal((off[-1], dim)).astype(np.float32)
    idx = _ffx.DeviceIndex(dim, capacity=len(vec))
    idx.stage(0, vec)
    idx.set_docs(off)
    nq, C = 400, 300
    qv = rng.standard_normal((nq, dim)).astype(np.float32)
    q_off = (np.arange(nq + 1) * C).astype(np.int64)
    lex = rng.uniform(0, 20, nq * C).astype(np.float32)
    for mode in (1, 2, 3, 4):
        pool = len(vec) if mode == 1 else n_docs
        cand = np.concatenate([rng.choice(pool, C, replace=False) for _ in range(nq)]).astype(np.int32)
        outs, names = [], []
        for kernel in (2, 0):
            _ffx.set_option("kernel", kernel)
            for sub in (nq, 7):
                o = idx.rerank_host(mode, qv[:sub], q_off[:sub + 1], cand[:sub * C], lex[:sub * C], 0.3, 50, want_ff=True, want_int=True)
                outs.append(o); names.append(_ffx.last_kernel())
        _ffx.set_option("kernel", 0)
        for a, b in ((0, 2), (1, 3)):
            for key in ("ff", "int", "topk_score"):
                assert (outs[a][key].view(np.uint32) == outs[b][key].view(np.uint32)).all(), (dim, mode, key)
            assert (outs[a]["topk_pos"] == outs[b]["topk_pos"]).all()
        print(dim, mode, "identical;", names[0].split("(")[0], "|", names[2].split("(")[0])
    idx.close()
