"""TEST INFRASTRUCTURE — a tiny stand-in for the subset of h5py that `OnDiskIndex` uses.

Neither h5py nor libhdf5 is installed in the build image or on the GPU box, so the write
side of `fast_forward.index.disk` (which needs h5py) cannot run there.  This module mimics the
h5py calls made by disk.py (File as a context manager, attrs, create_dataset with
maxshape/chunks, resize, slice and increasing-list indexing, fixed-width byte strings, groups,
`in`, `del`) on an in-memory tree, and persists that tree as genuine HDF5 bytes: written by
tests/h5_writer.py on close, re-read through the product's own native reader
(`fast_forward._h5.H5File`) on open.  So every OnDiskIndex test crosses the real file format,
and `OnDiskIndex.load` reads these files exactly as it would read h5py's.
tests/test_disk.py installs it as `h5py` only when the real package is missing.
"""

import os

import numpy as np

import h5_writer


class Dataset:
    def __init__(self, data, maxshape=None, chunks=None):
        self._a = data
        self.attrs = {}
        self.maxshape = maxshape
        self.chunks = chunks if chunks is not True else (min(len(data), 1024) or 1,) + data.shape[1:]

    shape = property(lambda self: self._a.shape)
    dtype = property(lambda self: self._a.dtype)

    def resize(self, size, axis=0):
        assert axis == 0 and self.maxshape is not None and self.maxshape[0] is None
        grown = np.zeros((size,) + self._a.shape[1:], self._a.dtype)
        keep = min(size, self._a.shape[0])
        grown[:keep] = self._a[:keep]
        self._a = grown

    @staticmethod
    def _check(key):
        if isinstance(key, list):
            if any(b <= a for a, b in zip(key, key[1:])):
                raise TypeError("Indexing elements must be in increasing order")
        return key

    def __getitem__(self, key):
        return self._a[self._check(key)].copy()

    def __setitem__(self, key, value):
        if self._a.dtype.kind == "S":
            value = np.asarray([v.encode() if isinstance(v, str) else v for v in np.atleast_1d(value)],
                               dtype=self._a.dtype)
        self._a[self._check(key)] = value

    def __len__(self):
        return len(self._a)


class Group:
    def __init__(self):
        self.attrs = {}
        self._items = {}

    def _walk(self, path, create=False):
        node = self
        for part in [p for p in path.split("/") if p]:
            if part not in node._items:
                if not create:
                    raise KeyError(path)
                node._items[part] = Group()
            node = node._items[part]
        return node

    def __contains__(self, path):
        try:
            self._walk(path)
            return True
        except KeyError:
            return False

    def __getitem__(self, path):
        return self._walk(path)

    def __delitem__(self, path):
        parent, _, leaf = path.rpartition("/")
        del self._walk(parent)._items[leaf]

    def items(self):
        return self._items.items()

    def create_group(self, path):
        return self._walk(path, create=True)

    def create_dataset(self, name, shape=None, dtype=None, data=None, maxshape=None, chunks=None):
        parent, _, leaf = name.rpartition("/")
        node = self._walk(parent, create=True)
        arr = np.array(data) if data is not None else np.zeros(shape, dtype)
        node._items[leaf] = Dataset(arr, maxshape, chunks)
        return node._items[leaf]


class File(Group):
    def __init__(self, path, mode="r"):
        super().__init__()
        self._path, self._mode = str(path), mode
        if mode in ("r", "a") and os.path.exists(self._path):
            from fast_forward._h5 import H5File

            with H5File(self._path) as src:
                self._adopt(src, "/", self)
        elif mode == "r":
            raise FileNotFoundError(path)

    @classmethod
    def _adopt(cls, src, path, node):
        node.attrs = src.attrs(path)
        for name in src.keys(path):
            child = path.rstrip("/") + "/" + name
            if src.kind(child) == 2:
                meta = src.info(child)
                chunked = meta["chunk_rows"] is not None
                ds = Dataset(src.read(child), (None,) + meta["shape"][1:] if chunked else None,
                             (meta["chunk_rows"],) + meta["shape"][1:] if chunked else None)
                ds.attrs = src.attrs(child)
                node._items[name] = ds
            else:
                node._items[name] = Group()
                cls._adopt(src, child, node._items[name])

    def _tree(self, node):
        out = h5_writer.Group()
        out.attrs = dict(node.attrs)
        for name, child in node._items.items():
            if isinstance(child, Dataset):
                ds = h5_writer.Dataset(child._a, child.chunks if child.maxshape is not None else None, child.maxshape)
                ds.attrs = dict(getattr(child, "attrs", {}))
                out.children[name] = ds
            else:
                out.children[name] = self._tree(child)
        return out

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        if self._mode != "r" and exc[0] is None:
            h5_writer.write_hdf5(self._tree(self), self._path)
        return False
