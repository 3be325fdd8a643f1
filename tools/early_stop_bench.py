"""Early stopping on the device (ffx_rerank_early_stop, index/base.py:316-387) at C3 shape:
time and HBM bytes of the depth walk next to scoring every candidate with the same kernel."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fast-forward-indexes_b200")); sys.path.insert(0, ROOT)
import bench
from fast_forward import _ffx
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
dev = torch.device("cuda", 0)
n_docs, nq, C, D = int(3_200_000 * scale), 5193, 5000, 768
cnt = bench.doc_lengths(n_docs, 6.25, seed=0)
off = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
n_rows = int(off[-1])
idx = _ffx.DeviceIndex(D, capacity=n_rows)
gen = torch.Generator(device=dev); gen.manual_seed(1234)
for r0 in range(0, n_rows, 1 << 20):
    nr = min(1 << 20, n_rows - r0)
    t = torch.randn((nr, D), device=dev, generator=gen); torch.cuda.synchronize()
    idx.stage_device(r0, nr, t.data_ptr()); del t
idx.set_docs(off)
qv = torch.randn((nq, D), device=dev, generator=gen)
bucket = n_docs // C
cand = (torch.rand((nq, C), device=dev, generator=gen).argsort(dim=1) * bucket +
        torch.randint(0, bucket, (nq, C), device=dev, generator=gen)).to(torch.int32).view(-1).contiguous()
# first-stage scores in rank order, steepness differs per query so that queries stop at different depths
steep = 0.995 + 0.0049 * torch.rand((nq, 1), device=dev, generator=gen)
lex = (300.0 * steep ** torch.arange(C, device=dev).view(1, -1)).float().view(-1).contiguous()
q_off = (torch.arange(nq + 1, device=dev, dtype=torch.int64) * C).contiguous()
d_cnt = torch.from_numpy(cnt).to(dev)
rows_of_pair = d_cnt[cand.long()].view(nq, C)
alpha, cutoff = 0.5, 10
out = {}
for name, depths in (("all_candidates", (C,)), ("early_stopping", (100, 250, 500, 1000, 2500, C))):
    ff = torch.zeros(nq * C, device=dev); scored = torch.zeros(nq, device=dev, dtype=torch.int32)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    st = torch.cuda.current_stream()
    for rep in range(4):
        if rep == 1:
            ev[0].record(st)
        idx.rerank_early_stop_device(2, qv.data_ptr(), nq, q_off.data_ptr(), cand.data_ptr(), lex.data_ptr(), alpha,
                                     cutoff, depths, C, ff.data_ptr(), 0, scored.data_ptr(), st.cuda_stream)
    ev[1].record(st); torch.cuda.synchronize(); idx.sync()
    ms = ev[0].elapsed_time(ev[1]) / 3
    mask = torch.arange(C, device=dev).view(1, -1) < scored.view(-1, 1)
    rows = int((rows_of_pair * mask).sum())
    out[name] = {"ms": round(ms, 2), "pairs_scored": int(scored.sum()), "fraction_scored": round(float(scored.sum()) / (nq * C), 4),
                 "row_bytes_GB": round(rows * D * 4 / 1e9, 1), "GB_per_s": round(rows * D * 4 / 1e9 / (ms * 1e-3), 0),
                 "depth_histogram": {int(k): int(v) for k, v in zip(*np.unique(scored.cpu().numpy(), return_counts=True))}}
print(json.dumps({"docs": n_docs, "queries": nq, "candidates": C, "alpha": alpha, "cutoff": cutoff, **out}))
