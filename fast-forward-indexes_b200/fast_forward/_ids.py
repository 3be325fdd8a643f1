"""Id dictionaries of an index: string ids -> integers, in C++ (libffx `ffx_dict_*`).

The reference keeps `dict[str, list[int]]` / `dict[str, int]` and hashes ids one Python object
at a time (index/memory.py:84-95, index/util.py:29-41, index/disk.py:408-417).  Here id
columns stay in Arrow buffers (what pandas >= 3 string columns are made of; anything else is
converted once) and are coded by the library on all host cores; only integers go on to the
GPU.  No CUDA device is needed for this module.
"""

from __future__ import annotations

import ctypes as C
from collections.abc import Iterable, Sequence

import numpy as np

from fast_forward import _ffx

try:  # pandas >= 3 depends on it; older environments fall back to a plain-Python encoder
    import pyarrow as pa
except ImportError:  # pragma: no cover - depends on the environment
    pa = None


class _Strings:
    """One Arrow-layout view: n strings, int64 offsets, UTF-8 bytes, optional validity bitmap."""

    __slots__ = ("n", "offsets", "data", "validity", "bit_offset", "keep")

    def __init__(self, n, offsets, data, validity, bit_offset, keep):
        self.n, self.offsets, self.data, self.validity, self.bit_offset, self.keep = n, offsets, data, validity, bit_offset, keep


def _from_arrow(arr) -> _Strings:
    if arr.type != pa.large_string():
        arr = arr.cast(pa.large_string())
    validity, offsets, data = arr.buffers()
    has_nulls = validity is not None and arr.null_count > 0
    return _Strings(len(arr), offsets.address + 8 * arr.offset, data.address if data is not None else 0,
                    validity.address if has_nulls else 0, arr.offset if has_nulls else 0, arr)


def _from_python(values: Sequence) -> _Strings:
    """Fallback without pyarrow: encode one string at a time."""
    enc = [b"" if v is None else str(v).encode("utf-8") for v in values]
    offsets = np.zeros(len(enc) + 1, np.int64)
    np.cumsum(np.fromiter(map(len, enc), np.int64, len(enc)), out=offsets[1:])
    data = np.frombuffer(b"".join(enc), np.uint8) if offsets[-1] else np.zeros(1, np.uint8)
    validity = np.packbits(np.fromiter((v is not None for v in values), bool, len(enc)), bitorder="little") \
        if any(v is None for v in values) else None
    return _Strings(len(enc), offsets.ctypes.data, data.ctypes.data,
                    validity.ctypes.data if validity is not None else 0, 0, (offsets, data, validity))


def string_views(values) -> list[_Strings]:
    """Arrow-layout views (one per chunk) of a pandas Series / Index / array, a pyarrow array, a
    numpy array or a sequence of `str | None` — zero-copy where the data already is Arrow."""
    if pa is None:
        return [_from_python(list(values))]
    if hasattr(values, "array") and not isinstance(values, (pa.Array, pa.ChunkedArray)):
        values = values.array  # pandas Series / Index -> extension array
    if isinstance(values, (pa.Array, pa.ChunkedArray)):
        arr = values
    elif hasattr(values, "__arrow_array__") and getattr(getattr(values, "dtype", None), "storage", "") == "pyarrow":
        arr = pa.array(values)  # pandas ArrowStringArray: zero copy
    elif hasattr(values, "to_numpy"):
        arr = pa.array(values.to_numpy(dtype=object, na_value=None), type=pa.large_string())
    else:
        if isinstance(values, np.ndarray) and values.dtype.kind in "US":
            values = values.astype(object)
        arr = pa.array(values if isinstance(values, (list, np.ndarray)) else list(values), type=pa.large_string())
    chunks = arr.chunks if isinstance(arr, pa.ChunkedArray) else [arr]
    return [_from_arrow(c) for c in chunks if len(c)]


def _vp(address: int):
    return C.c_void_p(address) if address else None


class IdDict:
    """string -> int64 map (one `ffx_dict`)."""

    def __init__(self) -> None:
        self._h = C.c_void_p()
        _ffx.check(_ffx.lib().ffx_dict_create(C.byref(self._h)))

    def close(self) -> None:
        if self._h is not None and self._h.value:
            _ffx.lib().ffx_dict_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self) -> int:
        return int(_ffx.lib().ffx_dict_size(self._h))

    def clone(self) -> "IdDict":
        """A deep copy (ffx_dict_clone)."""
        other = IdDict.__new__(IdDict)
        other._h = C.c_void_p()
        _ffx.check(_ffx.lib().ffx_dict_clone(self._h, C.byref(other._h)))
        return other

    # ---- building ---------------------------------------------------------------------------
    def insert_ordinal(self, values) -> np.ndarray:
        """Document ids: ordinal (order of first appearance) per value, -1 for None."""
        views = string_views(values)
        out = np.empty(sum(v.n for v in views), np.int64)
        at = 0
        for v in views:
            _ffx.check(_ffx.lib().ffx_dict_insert_ordinal(self._h, _vp(v.offsets), _vp(v.data), _vp(v.validity),
                                                          v.bit_offset, v.n, C.c_void_p(out[at:].ctypes.data)))
            at += v.n
        return out

    def insert_unique(self, values, first_value: int, dry_run: bool = False) -> int:
        """Passage ids: value i -> first_value + i.  Returns -1, or the index of the first value
        that already exists (nothing is inserted then)."""
        views = string_views(values)
        if len(views) > 1:  # one batch must be all-or-nothing: make it one chunk
            views = string_views(pa.concat_arrays([v.keep for v in views]))
        for v in views:
            dup = C.c_int64(-1)
            rc = _ffx.lib().ffx_dict_insert_unique(self._h, _vp(v.offsets), _vp(v.data), _vp(v.validity),
                                                   v.bit_offset, v.n, int(first_value), int(dry_run), C.byref(dup))
            if dup.value >= 0:
                return int(dup.value)
            _ffx.check(rc)
        return -1

    # ---- coding ------------------------------------------------------------------------------
    def lookup(self, values, threads: int = 0) -> tuple[np.ndarray, int]:
        """int32 value per string (-1 = absent / None) and the index of the first absent one
        (-1 when all are present)."""
        views = string_views(values)
        out = np.empty(sum(v.n for v in views), np.int32)
        first_missing, at = -1, 0
        for v in views:
            miss = C.c_int64(-1)
            _ffx.check(_ffx.lib().ffx_dict_lookup(self._h, _vp(v.offsets), _vp(v.data), _vp(v.validity), v.bit_offset,
                                                  v.n, C.c_void_p(out[at:].ctypes.data), C.byref(miss), int(threads)))
            if miss.value >= 0 and first_missing < 0:
                first_missing = at + int(miss.value)
            at += v.n
        return out, first_missing

    # ---- reading back ------------------------------------------------------------------------
    def export(self):
        """(keys, values): keys in insertion order (pyarrow large_string array, or a list without
        pyarrow) and their int64 values."""
        n = len(self)
        offsets = np.empty(n + 1, np.int64)
        data = np.empty(max(int(_ffx.lib().ffx_dict_key_bytes(self._h)), 1), np.uint8)
        values = np.empty(n, np.int64)
        _ffx.check(_ffx.lib().ffx_dict_export(self._h, C.c_void_p(offsets.ctypes.data), C.c_void_p(data.ctypes.data),
                                              C.c_void_p(values.ctypes.data)))
        if pa is None:
            raw = data.tobytes()
            return [raw[offsets[i]:offsets[i + 1]].decode("utf-8") for i in range(n)], values
        keys = pa.Array.from_buffers(pa.large_string(), n, [None, pa.py_buffer(offsets), pa.py_buffer(data)])
        return keys, values

    def keys(self) -> list[str]:
        keys, _ = self.export()
        return keys if isinstance(keys, list) else keys.to_pylist()


def factorize(values) -> tuple[np.ndarray, object]:
    """(int32 code per string, the distinct strings in code order) for a null-free string column,
    on all host cores (libffx `ffx_factorize`); the numbering is arbitrary but consistent."""
    views = string_views(values)
    if len(views) != 1:
        merged = pa.concat_arrays([v.keep for v in views]) if views else pa.array([], pa.large_string())
        views = string_views(merged) or [_from_arrow(merged)]
    v = views[0]
    if v.validity:
        raise ValueError("factorize: the column holds nulls")
    codes = np.empty(v.n, np.int32)
    handle, n_keys, key_bytes = C.c_void_p(), C.c_int64(), C.c_int64()
    _ffx.check(_ffx.lib().ffx_factorize(_vp(v.offsets), _vp(v.data), v.n, C.c_void_p(codes.ctypes.data) if v.n else None,
                                        C.byref(handle), C.byref(n_keys), C.byref(key_bytes), 0))
    try:
        offsets = np.empty(n_keys.value + 1, np.int64)
        data = np.empty(max(key_bytes.value, 1), np.uint8)
        _ffx.check(_ffx.lib().ffx_factor_export(handle, C.c_void_p(offsets.ctypes.data), C.c_void_p(data.ctypes.data)))
    finally:
        _ffx.lib().ffx_factor_free(handle)
    keys = pa.Array.from_buffers(pa.large_string(), n_keys.value, [None, pa.py_buffer(offsets), pa.py_buffer(data)])
    return codes, keys


def csr_from_ordinals(row_doc: np.ndarray, n_docs: int) -> tuple[np.ndarray, np.ndarray]:
    """doc -> rows CSR (offsets [n_docs+1], rows) from per-row document ordinals (-1 = none)."""
    row_doc = np.ascontiguousarray(row_doc, np.int64)
    off = np.empty(n_docs + 1, np.int64)
    rows = np.empty(int((row_doc >= 0).sum()), np.int64)
    _ffx.check(_ffx.lib().ffx_csr_build(C.c_void_p(row_doc.ctypes.data), len(row_doc), int(n_docs),
                                        C.c_void_p(off.ctypes.data), C.c_void_p(rows.ctypes.data) if len(rows) else None))
    return off, rows


def first_text(values, index: int) -> str:
    """The `index`-th id of a column, as text (for error messages)."""
    if hasattr(values, "iloc"):
        return str(values.iloc[index])
    return str(values[index])


def as_id_list(values: Iterable) -> list:
    return values if isinstance(values, list) else list(values)
