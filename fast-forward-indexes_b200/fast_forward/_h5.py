"""ctypes view of libffx's HDF5 reader (`ffx_h5_*`, csrc/ffx_h5.cpp, include/ffx.h).

`OnDiskIndex.load` opens index files through this instead of h5py: the file is mapped, its
chunks are handed to the staging buffers as pointers into the mapping, and the attributes /
small datasets of the `quantizer` group come back as numpy values shaped like the ones h5py
returns (reference: src/fast_forward/index/disk.py:355-418).
"""

from __future__ import annotations

import ctypes as C
import os
from collections.abc import Iterator

import numpy as np

from fast_forward._ffx import check, lib


def _numpy_type(cls: int, size: int, signed: int) -> np.dtype:
    try:
        if cls == 0:
            return np.dtype(f"{'i' if signed else 'u'}{size}")
        if cls == 1:
            return np.dtype(f"f{size}")
        if cls == 3:
            return np.dtype(f"S{size}")
        if cls == 8:  # h5py stores numpy bool as a one-byte enum
            return np.dtype(np.bool_) if size == 1 else np.dtype(f"{'i' if signed else 'u'}{size}")
    except TypeError:
        pass
    raise ValueError(f"HDF5 datatype (class {cls}, {size} bytes) has no numpy counterpart")


class H5File:
    """A read-only HDF5 file.  Object names are '/'-separated paths from the root."""

    def __init__(self, path: os.PathLike | str) -> None:
        self._h = C.c_void_p()
        check(lib().ffx_h5_open(os.fsencode(path), C.byref(self._h)))

    def close(self) -> None:
        if self._h:
            lib().ffx_h5_close(self._h)
            self._h = C.c_void_p()

    def __enter__(self) -> "H5File":
        return self

    def __exit__(self, *exc) -> None:
        self.close()

    def __del__(self) -> None:
        try:
            self.close()
        except Exception:  # interpreter shutdown
            pass

    # ---- structure ------------------------------------------------------------------------
    def kind(self, path: str) -> int:
        """0 = absent, 1 = group, 2 = dataset."""
        out = C.c_int()
        check(lib().ffx_h5_kind(self._h, path.encode(), C.byref(out)))
        return out.value

    def __contains__(self, path: str) -> bool:
        return self.kind(path) != 0

    def _text(self, fn, *args) -> bytes:
        need = C.c_int64()
        check(fn(self._h, *args, None, 0, C.byref(need)))
        buf = C.create_string_buffer(max(need.value, 1))
        check(fn(self._h, *args, buf, need.value, C.byref(need)))
        return buf.raw[: need.value]

    def keys(self, path: str = "/") -> list[str]:
        return [n.decode("utf-8", "replace") for n in self._text(lib().ffx_h5_list, path.encode()).split(b"\n") if n]

    # ---- attributes -----------------------------------------------------------------------
    def attr(self, path: str, name: str):
        info = (C.c_int64 * 13)()
        need = C.c_int64()
        check(lib().ffx_h5_attr_read(self._h, path.encode(), name.encode(), info, None, 0, C.byref(need)))
        buf = C.create_string_buffer(max(need.value, 1))
        check(lib().ffx_h5_attr_read(self._h, path.encode(), name.encode(), info, buf, need.value, C.byref(need)))
        raw = buf.raw[: need.value]
        cls, size, signed, rank, count = (int(v) for v in info[:5])
        shape = tuple(int(v) for v in info[5:5 + rank])
        if cls == 3:
            texts = [t.decode("utf-8", "replace") for t in raw.split(b"\0")] if count else []
            texts += [""] * (count - len(texts))
            return texts[0] if rank == 0 else np.array(texts, dtype=object).reshape(shape)
        values = np.frombuffer(raw, _numpy_type(cls, size, signed), count).copy()
        return values[0] if rank == 0 else values.reshape(shape)

    def attrs(self, path: str = "/") -> dict:
        names = [n.decode("utf-8", "replace") for n in self._text(lib().ffx_h5_attr_names, path.encode()).split(b"\n") if n]
        return {n: self.attr(path, n) for n in names}

    # ---- datasets -------------------------------------------------------------------------
    def info(self, path: str) -> dict:
        """dtype, shape, chunk_rows (None unless chunked), row_bytes of a dataset."""
        raw = (C.c_int64 * 16)()
        check(lib().ffx_h5_dataset_info(self._h, path.encode(), raw))
        rank = int(raw[3])
        return {
            "dtype": _numpy_type(int(raw[0]), int(raw[1]), int(raw[2])),
            "shape": tuple(int(v) for v in raw[4:4 + rank]),
            "chunk_rows": int(raw[13]) if raw[12] == 2 else None,
            "row_bytes": int(raw[14]),
        }

    def read(self, path: str, lo: int = 0, hi: int | None = None) -> np.ndarray:
        """rows [lo, hi) of axis 0 as a fresh array (`fp[path][lo:hi]`)."""
        meta = self.info(path)
        shape = meta["shape"]
        if not shape:  # scalar dataset
            out = np.empty((), meta["dtype"])
            check(lib().ffx_h5_read_rows(self._h, path.encode(), 0, 1, out.ctypes.data_as(C.c_void_p)))
            return out
        hi = shape[0] if hi is None else min(hi, shape[0])
        lo = min(lo, hi)
        out = np.empty((hi - lo,) + shape[1:], meta["dtype"])
        if out.size:
            check(lib().ffx_h5_read_rows(self._h, path.encode(), lo, hi - lo, out.ctypes.data_as(C.c_void_p)))
        return out

    def spans(self, path: str, lo: int, hi: int) -> Iterator[tuple[int, np.ndarray]]:
        """Rows [lo, hi) as (first_row, array) runs, one per stretch that is contiguous in the
        file (an HDF5 chunk).  The arrays are read-only VIEWS of the mapped file, valid until
        `close()`; chunks that were never written come back as zero arrays."""
        meta = self.info(path)
        inner, dt = meta["shape"][1:], meta["dtype"]
        per_row = int(np.prod(inner, dtype=np.int64)) if inner else 1
        row = lo
        while row < hi:
            ptr, n = C.c_void_p(), C.c_int64()
            check(lib().ffx_h5_row_span(self._h, path.encode(), row, C.byref(ptr), C.byref(n)))
            take = min(int(n.value), hi - row)
            if ptr.value:
                flat = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), (take * meta["row_bytes"],))
                block = flat.view(dt).reshape((take,) + inner) if per_row else flat.view(dt)
                block.flags.writeable = False
            else:
                block = np.zeros((take,) + inner, dt)
            yield row, block
            row += take
