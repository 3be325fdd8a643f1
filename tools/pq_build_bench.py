"""Encode / k-means throughput of the GPU PQ build (M=96, Ks=256, D=768) next to scipy on the host."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fast-forward-indexes_b200"))
from fast_forward import _ffx
from scipy.cluster.vq import vq, kmeans2
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
M, Ks, Ds = 96, 256, 8
rng = np.random.default_rng(0)
x = rng.standard_normal((n, M * Ds), dtype=np.float32)
cw = rng.standard_normal((M, Ks, Ds), dtype=np.float32)
_ffx.pq_encode(x[:1000], cw)
t = time.perf_counter(); codes = _ffx.pq_encode(x, cw); t_enc = time.perf_counter() - t
ns = min(n, 20000)
t = time.perf_counter()
host = np.stack([vq(x[:ns, m * Ds:(m + 1) * Ds], cw[m])[0] for m in range(M)], axis=1)
t_host = (time.perf_counter() - t) * n / ns
agree = float((host == codes[:ns]).mean())
nk = min(n, 200_000)
init = np.stack([x[rng.choice(nk, Ks, replace=False), m * Ds:(m + 1) * Ds] for m in range(M)])
t = time.perf_counter(); _ffx.pq_kmeans(x[:nk], init, 10); t_km = time.perf_counter() - t
t = time.perf_counter()
for m in range(4):
    kmeans2(x[:nk, m * Ds:(m + 1) * Ds], init[m].copy(), iter=10, minit="matrix")
t_km_host = (time.perf_counter() - t) * M / 4
print(json.dumps({"n": n, "encode_gpu_s": round(t_enc, 3), "encode_scipy_s_extrapolated": round(t_host, 1),
                  "codes_agree": agree, "kmeans_n": nk, "kmeans10_gpu_s": round(t_km, 3),
                  "kmeans10_scipy_s_extrapolated": round(t_km_host, 1)}))
