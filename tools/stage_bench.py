"""Host -> HBM staging throughput of ffx_index_stage_rows (pinned double buffer + permute)."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fast-forward-indexes_b200"))
from fast_forward import _ffx
gb = float(sys.argv[1]) if len(sys.argv) > 1 else 8.0
D = 768
n = int(gb * 1e9 / (D * 4))
x = np.empty((n, D), np.float32)
x[:] = np.arange(D, dtype=np.float32)  # touch every page
idx = _ffx.DeviceIndex(D, capacity=n)
idx.stage(0, x[:100000])
out = {}
for name, chunk in (("one_call", n), ("chunks_of_65536_rows", 65536)):
    t = time.perf_counter()
    for r0 in range(0, n, chunk):
        idx.stage(r0, x[r0:r0 + chunk])
    dt = time.perf_counter() - t
    out[name] = {"seconds": round(dt, 3), "GB_per_s": round(n * D * 4 / 1e9 / dt, 2)}
codes = np.zeros((n * 8, 96), np.uint8)
pq = _ffx.DeviceIndex(96, capacity=len(codes), row_kind=_ffx.ROWS_PQ_U8)
t = time.perf_counter(); pq.stage(0, codes); dt = time.perf_counter() - t
out["pq_codes"] = {"seconds": round(dt, 3), "GB_per_s": round(codes.nbytes / 1e9 / dt, 2)}
assert (idx.read_rows([0, n - 1]) == x[[0, n - 1]]).all()
print(json.dumps({"GB": gb, **out}))
