"""Throughput of the fused kernel at other vector dimensions (MAXP, device-resident inputs)."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fast-forward-indexes_b200")); sys.path.insert(0, ROOT)
import bench
from fast_forward import _ffx
dev = torch.device("cuda", 0)
for name in ("kernel", "tma_stages", "tma_warps", "batch"):  # A/B switches: FFX_OPT_kernel=3 ...
    if os.environ.get("FFX_OPT_" + name):
        _ffx.set_option(name, int(os.environ["FFX_OPT_" + name]))
out = {}
for D in [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "384,768,1024,2048,3072,4096").split(",")]:
    n_docs = int(24e9 / (D * 4 * 6.25))  # ~24 GB of rows
    nq, C, k = 1500, 2000, 2000
    cnt = bench.doc_lengths(n_docs, 6.25, seed=0)
    off = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
    n_rows = int(off[-1])
    idx = _ffx.DeviceIndex(D, capacity=n_rows)
    gen = torch.Generator(device=dev); gen.manual_seed(D)
    step = max(1, (1 << 30) // (D * 4))
    for r0 in range(0, n_rows, step):
        nr = min(step, n_rows - r0)
        t = torch.randn((nr, D), device=dev, generator=gen); torch.cuda.synchronize()
        idx.stage_device(r0, nr, t.data_ptr()); del t
    idx.set_docs(off)
    qv = torch.randn((nq, D), device=dev, generator=gen)
    bucket = n_docs // C
    cand = (torch.rand((nq, C), device=dev, generator=gen).argsort(dim=1) * bucket +
            torch.randint(0, bucket, (nq, C), device=dev, generator=gen)).to(torch.int32).view(-1).contiguous()
    lex = (torch.rand((nq * C,), device=dev, generator=gen) * 20).contiguous()
    q_off = (torch.arange(nq + 1, device=dev, dtype=torch.int64) * C).contiguous()
    ts = torch.empty((nq, k), device=dev); tp = torch.empty((nq, k), device=dev, dtype=torch.int32)
    rows = float(torch.from_numpy(cnt).to(dev)[cand.long()].sum())
    st = torch.cuda.current_stream()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for rep in range(6):
        if rep == 2:
            ev[0].record(st)
        idx.rerank_device(2, qv.data_ptr(), nq, q_off.data_ptr(), cand.data_ptr(), lex.data_ptr(), 0.1, k, C,
                          0, 0, ts.data_ptr(), tp.data_ptr(), st.cuda_stream)
    ev[1].record(st); torch.cuda.synchronize(); idx.sync()
    ms = ev[0].elapsed_time(ev[1]) / 4
    out[D] = {"ms": round(ms, 2), "pairs_per_s": round(nq * C / (ms * 1e-3)), "GB_per_s": round(rows * D * 4 / 1e9 / (ms * 1e-3))}
    idx.close(); del qv, cand, lex, ts, tp
    torch.cuda.empty_cache()
print(json.dumps(out))
