"""Dict-backed stand-in for the few h5py calls the self-play writers make (TEST INFRASTRUCTURE ONLY).

h5py / libhdf5 are not part of this image.  The reference's `Self_Play.play` (Self_Play.py:178-208) and this repo's
`ReplayWriter` HDF5 branch only use: `h5py.File(path, mode)` as a context manager, `file.keys()`, `file[name]` (a dataset
that supports `ds[i]`, `ds[i] = x`, `ds[i] += x`, `ds[:]`), `name in file`, and `file.create_dataset(name, data=...,
dtype=..., maxshape=..., chunks=...)`.  Files live in the process-wide `STORE` dict keyed by absolute path, so a test
can read back exactly what a writer produced.  `install()` registers the stub as the `h5py` module.
"""
import os
import sys
import types

import numpy as np

STORE = {}


class File:
    def __init__(self, path, mode="r"):
        self.path = os.path.abspath(path)
        if mode in ("w", "w-", "x"):
            STORE[self.path] = {}
        elif mode in ("a",):
            STORE.setdefault(self.path, {})
        elif self.path not in STORE:        # "r", "r+"
            raise FileNotFoundError(self.path)
        self.mode = mode
        self._d = STORE[self.path]

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def close(self):
        pass

    def keys(self):
        return self._d.keys()

    def __contains__(self, k):
        return k in self._d

    def __getitem__(self, k):
        return self._d[k]

    def create_dataset(self, name, shape=None, dtype=None, data=None, maxshape=None, chunks=None, **kw):
        if self.mode == "r":
            raise OSError("file is read-only")
        if name in self._d:
            raise ValueError("Unable to create dataset (name already exists)")
        if data is None:
            data = np.zeros(shape, dtype=dtype)
        arr = np.array(data, dtype=dtype if dtype is not None else np.asarray(data).dtype)
        if maxshape is not None and len(maxshape) != arr.ndim:
            raise ValueError("maxshape rank %d != data rank %d" % (len(maxshape), arr.ndim))
        self._d[name] = arr
        return arr


def install():
    mod = types.ModuleType("h5py")
    mod.File = File
    mod.__stub__ = True
    sys.modules["h5py"] = mod
    return mod


def exists(path):
    return os.path.abspath(path) in STORE
