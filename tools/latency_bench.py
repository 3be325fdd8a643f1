"""Serving latency: a handful of queries x 5000 candidates (MAXP, k=100) over the full C3 index,
device-resident inputs, CUDA events, median of 50 calls (each on different random candidates so
nothing is L2 resident)."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fast-forward-indexes_b200")); sys.path.insert(0, ROOT)
import bench
from fast_forward import _ffx
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
dev = torch.device("cuda", 0)
n_docs, C, D, k = int(3_200_000 * scale), 5000, 768, 100
cnt = bench.doc_lengths(n_docs, 6.25, seed=0)
off = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
n_rows = int(off[-1])
idx = _ffx.DeviceIndex(D, capacity=n_rows)
gen = torch.Generator(device=dev); gen.manual_seed(1)
for r0 in range(0, n_rows, 1 << 20):
    nr = min(1 << 20, n_rows - r0)
    t = torch.randn((nr, D), device=dev, generator=gen); torch.cuda.synchronize()
    idx.stage_device(r0, nr, t.data_ptr()); del t
idx.set_docs(off)
d_cnt = torch.from_numpy(cnt).to(dev)
out = {}
reps = int(os.environ.get('LAT_REPS', '50'))
for nq in [int(x) for x in os.environ.get('LAT_NQ', '1,2,4,8,16,32,64,128,256').split(',')]:
    qv = torch.randn((nq, D), device=dev, generator=gen)
    bucket = n_docs // C
    cands = [(torch.rand((nq, C), device=dev, generator=gen).argsort(dim=1) * bucket +
              torch.randint(0, bucket, (nq, C), device=dev, generator=gen)).to(torch.int32).view(-1).contiguous()
             for _ in range(reps + 5)]
    lex = (torch.rand((nq * C,), device=dev, generator=gen) * 20).contiguous()
    q_off = (torch.arange(nq + 1, device=dev, dtype=torch.int64) * C).contiguous()
    ts = torch.empty((nq, k), device=dev); tp = torch.empty((nq, k), device=dev, dtype=torch.int32)
    st = torch.cuda.current_stream()
    times = []
    for i, cand in enumerate(cands):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        idx.rerank_device(2, qv.data_ptr(), nq, q_off.data_ptr(), cand.data_ptr(), lex.data_ptr(), 0.1, k, C,
                          0, 0, ts.data_ptr(), tp.data_ptr(), st.cuda_stream)
        e1.record(st); torch.cuda.synchronize()
        if i >= 5:
            times.append(e0.elapsed_time(e1) * 1e3)
    rows = float(d_cnt[cands[-1].long()].sum())
    med = float(np.median(times))
    out[nq] = {"median_us": round(med, 1), "p95_us": round(float(np.percentile(times, 95)), 1),
               "GB_per_s": round(rows * D * 4 / 1e9 / (med * 1e-6), 0)}
print(json.dumps({"docs": n_docs, "candidates": C, "k": k, "latency_by_queries": out}))
