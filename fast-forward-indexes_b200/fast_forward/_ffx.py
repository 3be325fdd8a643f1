"""ctypes binding of libffx.so (the C ABI in include/ffx.h) + a thin numpy-level wrapper.

There is no CPU implementation behind this module: if the shared library has not been built
(`python -c "import __graft_entry__ as g; g.build()"`) or no CUDA device is visible, the
calls raise — nothing silently falls back to numpy.
"""

from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FFX_LIB") or os.path.join(_HERE, "lib", "libffx.so")  # FFX_LIB: an alternative build (A/B measurements)

ROWS_F32, ROWS_PQ_U8 = 0, 1
MODE_PASSAGE, MODE_MAXP, MODE_FIRSTP, MODE_AVEP = 1, 2, 3, 4

# every symbol include/ffx.h declares: name -> (restype, argtypes)
_P, _I, _L, _D = C.c_void_p, C.c_int, C.c_int64, C.c_double
SYMBOLS = {
    "ffx_abi_version": (_I, []),
    "ffx_last_error": (C.c_char_p, []),
    "ffx_device_count": (_I, []),
    "ffx_set_option": (_I, [C.c_char_p, _I]),
    "ffx_host_alloc": (_I, [C.POINTER(_P), _L]),
    "ffx_host_free": (_I, [_P]),
    "ffx_index_create": (_I, [_I, _I, _L, _L, C.POINTER(_P)]),
    "ffx_index_destroy": (_I, [_P]),
    "ffx_index_reserve": (_I, [_P, _L]),
    "ffx_index_copy_rows": (_I, [_P, _P]),
    "ffx_index_stage_rows": (_I, [_P, _L, _L, _P, _I]),
    "ffx_index_read_rows": (_I, [_P, _P, _L, _P]),
    "ffx_index_num_rows": (_L, [_P]),
    "ffx_index_capacity": (_L, [_P]),
    "ffx_index_dim": (_L, [_P]),
    "ffx_index_has_fast_path": (_I, [_P]),
    "ffx_index_set_docs": (_I, [_P, _L, _P, _P]),
    "ffx_index_set_shard": (_I, [_P, _L, _L, _L, _L]),
    "ffx_index_set_pq": (_I, [_P, _I, _I, _I, _P, _P]),
    "ffx_index_set_topk_scatter": (_I, [_P, _I, _I, _L, _P, _P, _P]),
    "ffx_rerank": (_I, [_P, _I, _P, _L, _P, _P, _P, _D, _I, _L, _P, _P, _P, _P, _P]),
    "ffx_rerank_host": (_I, [_P, _I, _P, _L, _P, _P, _P, _D, _I, _P, _P, _P, _P]),
    "ffx_rerank_early_stop": (_I, [_P, _I, _P, _L, _P, _P, _P, _D, _I, _P, _I, _L, _P, _P, _P, _P]),
    "ffx_rerank_early_stop_host": (_I, [_P, _I, _P, _L, _P, _P, _P, _D, _I, _P, _I, _P, _P, _P]),
    "ffx_index_sync": (_I, [_P, _P]),
    "ffx_index_coalesce": (_I, [_P, _L, _L, _P, _D, _P, _P]),
    "ffx_interpolate_topk": (_I, [_P, _P, _P, _L, _P, _D, _I, _L, _P, _P, _P, _P]),
    "ffx_interpolate_topk_host": (_I, [_P, _P, _P, _L, _P, _D, _I, _P, _P, _P]),
    "ffx_merge_topk": (_I, [_I, _P, _P, _I, _L, _I, _P, _P, _P]),
    "ffx_merge_topk_host": (_I, [_P, _P, _P, _I, _L, _I, _P, _P]),
    "ffx_launch_count": (_L, []),
    "ffx_last_kernel": (C.c_char_p, []),
    "ffx_fixed_width_to_arrow": (_I, [_P, _L, _I, _P, _P, _P, C.POINTER(_L)]),
    "ffx_first_repeat": (_I, [_P, _L, C.POINTER(_L)]),
    "ffx_ranking_order": (_I, [_P, _P, _L, _P, _I]),
    "ffx_order_u64": (_I, [_P, _L, _P, _I]),
    "ffx_match_keys": (_I, [_P, _L, _P, _L, _P]),
    "ffx_run_open": (_I, [C.c_char_p, _I, C.POINTER(_P), _P]),
    "ffx_run_read": (_I, [_P, _P, _P, _P, _P, _P, _P]),
    "ffx_run_close": (None, [_P]),
    "ffx_run_write": (_I, [C.c_char_p, _L, _P, _P, _P, _P, _P, _P, _P, C.c_char_p, _I]),
    "ffx_tie_runs": (_I, [_P, _P, _L, _L, _P, _P, C.POINTER(_L), _I]),
    "ffx_lut_gather": (_I, [_P, _L, _P, _L, _P, C.POINTER(_L), _I]),
    "ffx_topk_gather": (_I, [_P, _P, _L, _L, _L, _P, _P, _P, _P, _P, C.POINTER(_L), _P, _I]),
    "ffx_h5_open": (_I, [C.c_char_p, C.POINTER(_P)]),
    "ffx_h5_close": (None, [_P]),
    "ffx_h5_kind": (_I, [_P, C.c_char_p, C.POINTER(_I)]),
    "ffx_h5_list": (_I, [_P, C.c_char_p, _P, _L, C.POINTER(_L)]),
    "ffx_h5_attr_names": (_I, [_P, C.c_char_p, _P, _L, C.POINTER(_L)]),
    "ffx_h5_dataset_info": (_I, [_P, C.c_char_p, _P]),
    "ffx_h5_read_rows": (_I, [_P, C.c_char_p, _L, _L, _P]),
    "ffx_h5_row_span": (_I, [_P, C.c_char_p, _L, C.POINTER(_P), C.POINTER(_L)]),
    "ffx_h5_attr_read": (_I, [_P, C.c_char_p, C.c_char_p, _P, _P, _L, C.POINTER(_L)]),
    "ffx_pq_encode": (_I, [_I, _P, _L, _I, _I, _I, _P, _P]),
    "ffx_pq_kmeans": (_I, [_I, _P, _L, _I, _I, _I, _P, _I]),
    "ffx_sgemm": (_I, [_I, _I, _L, _L, _L, _P, _P, _P]),
    "ffx_dict_create": (_I, [C.POINTER(_P)]),
    "ffx_dict_destroy": (_I, [_P]),
    "ffx_dict_size": (_L, [_P]),
    "ffx_dict_key_bytes": (_L, [_P]),
    "ffx_dict_insert_ordinal": (_I, [_P, _P, _P, _P, _L, _L, _P]),
    "ffx_dict_insert_unique": (_I, [_P, _P, _P, _P, _L, _L, _L, _I, C.POINTER(_L)]),
    "ffx_dict_lookup": (_I, [_P, _P, _P, _P, _L, _L, _P, C.POINTER(_L), _I]),
    "ffx_dict_export": (_I, [_P, _P, _P, _P]),
    "ffx_dict_clone": (_I, [_P, C.POINTER(_P)]),
    "ffx_csr_build": (_I, [_P, _L, _L, _P, _P]),
    "ffx_factorize": (_I, [_P, _P, _L, _P, C.POINTER(_P), C.POINTER(_L), C.POINTER(_L), _I]),
    "ffx_factor_export": (_I, [_P, _P, _P]),
    "ffx_factor_free": (None, [_P]),
}

_lib = None


class FFXError(RuntimeError):
    """A libffx call failed (status code + message of ffx_last_error)."""

    def __init__(self, code: int, message: str):
        super().__init__(f"libffx error {code}: {message}")
        self.code = code


def lib():
    """Load libffx.so once.  Raises (never falls back) when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not built. Build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` from the repository root; fast_forward has no CPU scoring path."
            )
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        if handle.ffx_abi_version() != 1:
            raise ImportError("libffx.so ABI version mismatch; rebuild it")
        _lib = handle
    return _lib


def check(code: int) -> None:
    if code != 0:
        raise FFXError(code, lib().ffx_last_error().decode("utf-8", "replace"))


def device_count() -> int:
    return lib().ffx_device_count()


def set_option(name: str, value: int) -> None:
    """ffx_set_option: kernel / tma_stages / batch tuning knobs (0 = automatic)."""
    check(lib().ffx_set_option(name.encode(), int(value)))


def launch_count() -> int:
    return lib().ffx_launch_count()


def last_kernel() -> str:
    """Demangled symbol of the scoring kernel launched last ("" before the first launch)."""
    return lib().ffx_last_kernel().decode("utf-8", "replace")


def _ptr(a):
    if a is None:
        return None
    return C.c_void_p(a.ctypes.data)


def _arr(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


class PinnedBuffer:
    """Page-locked host memory from ffx_host_alloc, exposed as a numpy array."""

    def __init__(self, shape, dtype):
        self.shape = tuple(int(s) for s in np.atleast_1d(shape))
        self.dtype = np.dtype(dtype)
        nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        self._p = C.c_void_p()
        check(lib().ffx_host_alloc(C.byref(self._p), nbytes))
        if nbytes:
            buf = (C.c_char * nbytes).from_address(self._p.value)
            self.array = np.frombuffer(buf, dtype=self.dtype).reshape(self.shape)
        else:
            self.array = np.empty(self.shape, self.dtype)

    def free(self):
        if self._p is not None and self._p.value:
            self.array = None
            lib().ffx_host_free(self._p)
        self._p = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class _PinnedPool:
    """Caching allocator for page-locked host memory.  Pinning pages costs far more than using
    them (cudaHostAlloc of 200 MB: tens of milliseconds), so the large per-call arrays of the
    Python shell — candidate codes, scores, ranked lists — come out of SLABS pinned four blocks
    at a time and are recycled through per-size free lists: an array from `pinned_empty` returns
    its block here when the last view of it is garbage-collected.  Slabs are kept until exit."""

    MIN_BYTES = 1 << 20
    MAX_SLAB = 1 << 30

    def __init__(self):
        self.free: dict[int, list[int]] = {}   # size class -> addresses of free blocks
        self.slabs: list[tuple[int, int, int]] = []  # (base address, bytes, bump offset) of live slabs
        self.pinned_bytes = 0

    @staticmethod
    def size_class(nbytes: int) -> int:
        c = 1 << 20
        while c < nbytes:
            c += max(c >> 2, 1 << 20)  # geometric steps of 25 %
        return c

    def take(self, nbytes: int) -> tuple[int, int]:
        c = self.size_class(nbytes)
        stack = self.free.get(c)
        if stack:
            return c, stack.pop()
        for i, (base, size, used) in enumerate(self.slabs):
            if size - used >= c:
                self.slabs[i] = (base, size, used + c)
                return c, base + used
        slab = c if c >= self.MAX_SLAB else min(self.MAX_SLAB, 4 * c)
        p = C.c_void_p()
        check(lib().ffx_host_alloc(C.byref(p), slab))
        self.pinned_bytes += slab
        self.slabs.append((p.value, slab, c))
        return c, p.value

    def give(self, c: int, address: int) -> None:
        self.free.setdefault(c, []).append(address)


_POOL = _PinnedPool()


def pinned_empty(shape, dtype) -> np.ndarray:
    """np.empty in page-locked memory (recycled through a pool) for arrays that cross PCIe; small
    arrays, and any array when no CUDA device is present, are ordinary numpy arrays."""
    import weakref

    shape = tuple(int(x) for x in np.atleast_1d(shape))
    dtype = np.dtype(dtype)
    count = int(np.prod(shape))
    nbytes = count * dtype.itemsize
    if nbytes < _PinnedPool.MIN_BYTES or not _has_device():
        return np.empty(shape, dtype)
    c, p = _POOL.take(nbytes)
    raw = (C.c_char * nbytes).from_address(p)
    root = np.frombuffer(raw, dtype=dtype, count=count)
    weakref.finalize(root, _POOL.give, c, p)
    return root.reshape(shape)


_DEVICE = None


def _has_device() -> bool:
    global _DEVICE
    if _DEVICE is None:
        _DEVICE = device_count() > 0
    return _DEVICE


class DeviceIndex:
    """An HBM-resident row store + doc->rows map + optional PQ codebooks (one ffx_index)."""

    def __init__(self, dim: int, capacity: int = 0, row_kind: int = ROWS_F32, device: int = 0):
        self._h = C.c_void_p()
        self.row_kind = row_kind
        self.dim = int(dim)
        self.device = device
        check(lib().ffx_index_create(device, row_kind, self.dim, int(capacity), C.byref(self._h)))

    # -- lifetime -----------------------------------------------------------------------
    def close(self):
        if self._h is not None and self._h.value:
            lib().ffx_index_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        if self._h is None:
            raise RuntimeError("index closed")
        return self._h

    # -- storage ------------------------------------------------------------------------
    def __len__(self):
        return int(lib().ffx_index_num_rows(self.handle))

    @property
    def capacity(self):
        return int(lib().ffx_index_capacity(self.handle))

    @property
    def has_fast_path(self):
        return bool(lib().ffx_index_has_fast_path(self.handle))

    def reserve(self, capacity: int):
        check(lib().ffx_index_reserve(self.handle, int(capacity)))

    def copy_rows_from(self, other: "DeviceIndex"):
        """ffx_index_copy_rows: all rows of `other` into this (empty) index, device to device."""
        check(lib().ffx_index_copy_rows(self.handle, other.handle))

    def stage(self, row0: int, rows: np.ndarray):
        dt = np.float32 if self.row_kind == ROWS_F32 else np.uint8
        rows = _arr(rows, dt)
        if rows.ndim != 2 or rows.shape[1] != self.dim:
            raise ValueError(f"expected rows of shape [n, {self.dim}], got {rows.shape}")
        check(lib().ffx_index_stage_rows(self.handle, int(row0), rows.shape[0], _ptr(rows), 0))

    def stage_device(self, row0: int, nrows: int, device_ptr: int):
        check(lib().ffx_index_stage_rows(self.handle, int(row0), int(nrows), C.c_void_p(device_ptr), 1))

    def read_rows(self, rows) -> np.ndarray:
        rows = _arr(rows, np.int64)
        dt = np.float32 if self.row_kind == ROWS_F32 else np.uint8
        out = np.empty((len(rows), self.dim), dt)
        check(lib().ffx_index_read_rows(self.handle, _ptr(rows), len(rows), _ptr(out)))
        return out

    def set_docs(self, doc_off, doc_rows=None):
        doc_off = _arr(doc_off, np.int64)
        doc_rows = None if doc_rows is None else _arr(doc_rows, np.int64)
        check(lib().ffx_index_set_docs(self.handle, len(doc_off) - 1, _ptr(doc_off), _ptr(doc_rows)))

    def set_shard(self, doc_base=0, global_docs=0, row_base=0, global_rows=0):
        check(lib().ffx_index_set_shard(self.handle, int(doc_base), int(global_docs), int(row_base),
                                        int(global_rows)))

    def set_topk_scatter(self, world=0, rank=0, stride=0, bounds=None, peer_score=None, peer_pos=None):
        """ffx_index_set_topk_scatter: peer receive buffers (lists of device pointers, one per
        owner rank) the fused kernel writes its top-k lists into; world=0 removes the plan."""
        if not world:
            check(lib().ffx_index_set_topk_scatter(self.handle, 0, 0, 0, None, None, None))
            return
        b = _arr(bounds, np.int64)
        ps = (C.c_void_p * world)(*[int(x) for x in peer_score])
        pp = (C.c_void_p * world)(*[int(x) for x in peer_pos])
        check(lib().ffx_index_set_topk_scatter(self.handle, int(world), int(rank), int(stride), _ptr(b), ps, pp))

    def set_pq(self, codewords, R=None):
        codewords = _arr(codewords, np.float32)
        M, Ks, Ds = codewords.shape
        R = None if R is None else _arr(R, np.float32)
        check(lib().ffx_index_set_pq(self.handle, M, Ks, Ds, _ptr(codewords), _ptr(R)))

    # -- the hot path -------------------------------------------------------------------
    def rerank_host(self, mode, qvecs, q_off, cand, lex=None, alpha=0.0, k=0, want_ff=True,
                    want_int=False, out=None):
        """ffx_rerank_host on numpy (or pinned) buffers.  Returns a dict with the requested
        outputs: ff [n], int [n], topk_score / topk_pos [nq, k]."""
        qvecs = _arr(qvecs, np.float32)
        q_off = _arr(q_off, np.int64)
        cand = _arr(cand, np.int32)
        lex = None if lex is None else _arr(lex, np.float32)
        nq, n = len(q_off) - 1, len(cand)
        out = {} if out is None else out
        ff = out.get("ff") if want_ff else None
        if want_ff and ff is None:
            ff = out["ff"] = np.empty(n, np.float32)
        it = out.get("int") if want_int else None
        if want_int and it is None:
            it = out["int"] = np.empty(n, np.float32)
        ts = tp = None
        if k > 0:
            ts = out.get("topk_score")
            tp = out.get("topk_pos")
            if ts is None:
                ts = out["topk_score"] = np.empty((nq, k), np.float32)
            if tp is None:
                tp = out["topk_pos"] = np.empty((nq, k), np.int32)
        check(lib().ffx_rerank_host(self.handle, int(mode), _ptr(qvecs), nq, _ptr(q_off), _ptr(cand),
                                    _ptr(lex), float(alpha), int(k), _ptr(ff), _ptr(it), _ptr(ts),
                                    _ptr(tp)))
        return out

    def rerank_early_stop_host(self, mode, qvecs, q_off, cand, lex, alpha, cutoff, depths, want_int=False):
        """ffx_rerank_early_stop_host: depth-interval scoring with the early-stopping criterion
        evaluated on the device.  Returns ff [n] (0 where not scored), scored [nq] (rows scored
        per query, a prefix of each block) and optionally int [n]."""
        qvecs = _arr(qvecs, np.float32)
        q_off = _arr(q_off, np.int64)
        cand = _arr(cand, np.int32)
        lex = _arr(lex, np.float32)
        depths = _arr(list(depths), np.int32)
        nq, n = len(q_off) - 1, len(cand)
        out = {"ff": np.zeros(n, np.float32), "scored": np.zeros(nq, np.int32)}
        if want_int:
            out["int"] = np.zeros(n, np.float32)
        check(lib().ffx_rerank_early_stop_host(self.handle, int(mode), _ptr(qvecs), nq, _ptr(q_off), _ptr(cand),
                                               _ptr(lex), float(alpha), int(cutoff), _ptr(depths), len(depths),
                                               _ptr(out["ff"]), _ptr(out.get("int")), _ptr(out["scored"])))
        return out

    def rerank_early_stop_device(self, mode, qvecs_ptr, nq, q_off_ptr, cand_ptr, lex_ptr, alpha, cutoff, depths,
                                 max_cand, out_ff_ptr, out_int_ptr, scored_ptr, stream=0):
        """ffx_rerank_early_stop on raw device pointers (`depths` is a host sequence); asynchronous."""
        vp = lambda x: C.c_void_p(x) if x else None  # noqa: E731
        depths = _arr(list(depths), np.int32)
        check(lib().ffx_rerank_early_stop(self.handle, int(mode), vp(qvecs_ptr), int(nq), vp(q_off_ptr), vp(cand_ptr),
                                          vp(lex_ptr), float(alpha), int(cutoff), _ptr(depths), len(depths),
                                          int(max_cand), vp(out_ff_ptr), vp(out_int_ptr), vp(scored_ptr), vp(stream)))

    def sync(self, stream=0):
        """ffx_index_sync: wait for `stream`, raise if a kernel saw an out-of-range candidate."""
        check(lib().ffx_index_sync(self.handle, C.c_void_p(stream) if stream else None))

    def interpolate_topk_host(self, lex, ff, q_off, alpha, k, want_int=True):
        """ffx_interpolate_topk_host: interpolation + per-query ordering of existing scores."""
        ff = _arr(ff, np.float32)
        lex = None if lex is None else _arr(lex, np.float32)
        q_off = _arr(q_off, np.int64)
        nq = len(q_off) - 1
        out = {}
        it = out["int"] = np.empty(len(ff), np.float32) if want_int else None
        ts = tp = None
        if k > 0:
            ts = out["topk_score"] = np.empty((nq, k), np.float32)
            tp = out["topk_pos"] = np.empty((nq, k), np.int32)
        check(lib().ffx_interpolate_topk_host(self.handle, _ptr(lex), _ptr(ff), nq, _ptr(q_off),
                                              float(alpha), int(k), _ptr(it), _ptr(ts), _ptr(tp)))
        return out

    def coalesce(self, doc0: int, doc_off, delta: float):
        """ffx_index_coalesce over documents [doc0, doc0 + len(doc_off) - 1): (vectors [rows, dim] with
        every document's group means at the start of its row range, groups per document)."""
        doc_off = _arr(doc_off, np.int64)
        n_docs = len(doc_off) - 1
        out = np.empty((int(doc_off[-1]), self.dim), np.float32)
        groups = np.empty(n_docs, np.int32)
        check(lib().ffx_index_coalesce(self.handle, int(doc0), n_docs, _ptr(doc_off), float(delta), _ptr(out), _ptr(groups)))
        return out, groups

    def merge_topk_host(self, shard_scores, shard_pos):
        """ffx_merge_topk_host: [S, nq, k] per-shard lists (positions in the full candidate blocks)
        -> the merged [nq, k] lists."""
        shard_scores = _arr(shard_scores, np.float32)
        shard_pos = _arr(shard_pos, np.int32)
        S, nq, k = shard_scores.shape
        out_s, out_p = np.empty((nq, k), np.float32), np.empty((nq, k), np.int32)
        check(lib().ffx_merge_topk_host(self.handle, _ptr(shard_scores), _ptr(shard_pos), S, nq, k, _ptr(out_s),
                                        _ptr(out_p)))
        return out_s, out_p

    def rerank_device(self, mode, qvecs_ptr, nq, q_off_ptr, cand_ptr, lex_ptr, alpha, k, max_cand,
                      out_ff_ptr=0, out_int_ptr=0, topk_score_ptr=0, topk_pos_ptr=0, stream=0):
        """ffx_rerank on raw device pointers (ints); asynchronous on `stream`."""
        vp = lambda x: C.c_void_p(x) if x else None  # noqa: E731
        check(lib().ffx_rerank(self.handle, int(mode), vp(qvecs_ptr), int(nq), vp(q_off_ptr),
                               vp(cand_ptr), vp(lex_ptr), float(alpha), int(k), int(max_cand),
                               vp(out_ff_ptr), vp(out_int_ptr), vp(topk_score_ptr), vp(topk_pos_ptr),
                               vp(stream)))


def pq_encode(vecs, codewords, device: int = 0) -> np.ndarray:
    """ffx_pq_encode: nearest codeword per subspace on the GPU; uint8 codes [n, M]."""
    vecs = _arr(vecs, np.float32)
    codewords = _arr(codewords, np.float32)
    M, Ks, Ds = codewords.shape
    if vecs.ndim != 2 or vecs.shape[1] != M * Ds:
        raise ValueError(f"expected vectors of shape [n, {M * Ds}], got {vecs.shape}")
    codes = np.empty((vecs.shape[0], M), np.uint8)
    check(lib().ffx_pq_encode(int(device), _ptr(vecs), vecs.shape[0], M, Ks, Ds, _ptr(codewords), _ptr(codes)))
    return codes


def pq_kmeans(vecs, init_codewords, iters: int, device: int = 0) -> np.ndarray:
    """ffx_pq_kmeans: `iters` Lloyd rounds per subspace on the GPU from `init_codewords` [M, Ks, Ds]."""
    vecs = _arr(vecs, np.float32)
    codewords = np.array(init_codewords, dtype=np.float32, order="C", copy=True)
    M, Ks, Ds = codewords.shape
    if vecs.ndim != 2 or vecs.shape[1] != M * Ds:
        raise ValueError(f"expected vectors of shape [n, {M * Ds}], got {vecs.shape}")
    check(lib().ffx_pq_kmeans(int(device), _ptr(vecs), vecs.shape[0], M, Ks, Ds, _ptr(codewords), int(iters)))
    return codewords


def sgemm(a, b, trans_a: bool = False, device: int = 0) -> np.ndarray:
    """ffx_sgemm: `a @ b` (or `a.T @ b`) in fp32 on the GPU; a is [m, k] (or [k, m]), b is [k, n]."""
    a, b = _arr(a, np.float32), _arr(b, np.float32)
    m, k = (a.shape[1], a.shape[0]) if trans_a else a.shape
    if b.shape[0] != k:
        raise ValueError(f"shapes {a.shape} and {b.shape} do not multiply")
    out = np.empty((m, b.shape[1]), np.float32)
    check(lib().ffx_sgemm(int(device), int(trans_a), m, b.shape[1], k, _ptr(a), _ptr(b), _ptr(out)))
    return out


def merge_topk(device, shard_scores_ptr, shard_pos_ptr, n_shards, nq, k, out_score_ptr, out_pos_ptr,
               stream=0):
    check(lib().ffx_merge_topk(int(device), C.c_void_p(shard_scores_ptr), C.c_void_p(shard_pos_ptr),
                               int(n_shards), int(nq), int(k), C.c_void_p(out_score_ptr),
                               C.c_void_p(out_pos_ptr), C.c_void_p(stream) if stream else None))
