"""Cold-start of an on-disk index: `OnDiskIndex.load` of an HDF5 file (page cache -> pinned
buffers -> HBM, ids coded by the C++ dictionaries) through the native reader (csrc/ffx_h5.cpp).

The file is produced by tests/h5_writer.py in the layout the reference writes
(index/disk.py:138-165: `vectors` chunked (65536, dim), fixed-width id columns)."""
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fast-forward-indexes_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import h5_writer as hw  # noqa: E402
from fast_forward import _h5  # noqa: E402
from fast_forward.index import Mode, OnDiskIndex  # noqa: E402

gb = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
D, CHUNK = 768, 65536
n = int(gb * 1e9 / (D * 4))
rng = np.random.default_rng(0)
vec = np.empty((n, D), np.float32)
vec[:] = rng.standard_normal((1, D), dtype=np.float32)
vec[:, 0] = np.arange(n, dtype=np.float32)
docs = np.char.add("d", (np.arange(n) // 6).astype(str)).astype("S8")
psgs = np.char.add("p", np.arange(n).astype(str)).astype("S8")
root = hw.Group()
root.attrs = {"num_vectors": np.int64(n), "ff_version": "0.8.0"}
root.children["vectors"] = hw.Dataset(vec, (CHUNK, D), (None, D))
root.children["doc_ids"] = hw.Dataset(docs, (8192,), (None,))
root.children["psg_ids"] = hw.Dataset(psgs, (8192,), (None,))
tmp = tempfile.mkdtemp(dir=os.environ.get("FFX_TMP", "/dev/shm" if os.path.isdir("/dev/shm") else None))
path = os.path.join(tmp, "index.h5")
t = time.perf_counter()
hw.write_hdf5(root, path)
write_s = time.perf_counter() - t
size = os.path.getsize(path)

out = {"GB": round(size / 1e9, 2), "rows": n, "write_s": round(write_s, 2)}
for trial in ("first", "second"):
    t = time.perf_counter()
    from pathlib import Path

    index = OnDiskIndex.load(Path(path), mode=Mode.MAXP)
    index._store.dev.sync()
    dt = time.perf_counter() - t
    out[f"load_{trial}"] = {"seconds": round(dt, 3), "GB_per_s": round(size / 1e9 / dt, 2)}
    got = index._store.read(np.array([0, n // 2, n - 1]))
    assert (got == vec[[0, n // 2, n - 1]]).all()
    assert len(index) == n and len(index.doc_ids) == (n + 5) // 6
    if trial == "second":  # OnDiskIndex.to_memory(): a device-to-device copy of the rows + two dictionary copies
        t = time.perf_counter()
        mem = index.to_memory()
        mem._store.dev.sync()
        dt = time.perf_counter() - t
        out["to_memory"] = {"seconds": round(dt, 3), "GB_per_s": round(size / 1e9 / dt, 2)}
        assert (mem._store.read(np.array([0, n // 2, n - 1])) == vec[[0, n // 2, n - 1]]).all() and len(mem) == n
        del mem
    del index

# the pieces: rows only, ids only
from fast_forward import _ffx  # noqa: E402
from fast_forward.index.disk import _text_ids  # noqa: E402
from fast_forward.index._store import RowStore  # noqa: E402

with _h5.H5File(path) as fp:
    dev = _ffx.DeviceIndex(D, capacity=n)
    t = time.perf_counter()
    for row, block in fp.spans("vectors", 0, n):
        dev.stage(row, block)
    dev.sync()
    dt = time.perf_counter() - t
    out["rows_only"] = {"seconds": round(dt, 3), "GB_per_s": round(n * D * 4 / 1e9 / dt, 2)}
    t = time.perf_counter()
    store = RowStore(0)
    store.count = n
    store.adopt_id_columns(_text_ids(fp.read("doc_ids", 0, n)), _text_ids(fp.read("psg_ids", 0, n)))
    dt = time.perf_counter() - t
    out["ids_only"] = {"seconds": round(dt, 3), "M_rows_per_s": round(n / 1e6 / dt, 2)}
os.remove(path)
os.rmdir(tmp)
print(json.dumps(out))
