"""EmbeddingStore - the `<video>_cls.h5` contract (cbas.py:413-421 writer, cbas.py:485-507 reader).

Layout: dataset "cls", shape (N, D) float16, chunks (8192, D), maxshape (None, D); file attributes
`encoder_model_identifier` and `schema_version = "1.0"` (stamped only when a project is loaded, cbas.py:414).
Files are written as `<out>.tmp` and published with os.replace by the caller.

Backend: h5py when it is importable (the reference's own library), otherwise the native implementation in
`hdf5_min` (same on-disk format).  Both expose the same two small classes.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np

try:  # pragma: no cover - h5py is absent where this package is developed
    import h5py  # type: ignore
    HAVE_H5PY = True
except Exception:  # ImportError, or a broken libhdf5
    h5py = None
    HAVE_H5PY = False

from . import hdf5_min

DATASET = "cls"
CHUNK_ROWS = 8192
SCHEMA_VERSION = "1.0"


def backend_name() -> str:
    """'h5py' when the reference's own HDF5 library is importable, else the native writer/reader."""
    return "h5py" if HAVE_H5PY else "native"


class EmbeddingWriter:
    """Append float16 rows to a new embedding file."""

    def __init__(self, path: str, width: int, attrs: Optional[Dict[str, str]] = None, backend: Optional[str] = None):
        self.path, self.width = path, int(width)
        self.backend = backend or ("h5py" if HAVE_H5PY else "native")
        self.rows = 0
        if self.backend == "h5py":
            self._f = h5py.File(path, "w")
            for k, v in (attrs or {}).items():
                self._f.attrs[k] = v
            self._d = self._f.create_dataset(DATASET, shape=(0, self.width), maxshape=(None, self.width),
                                             dtype="f2", chunks=(CHUNK_ROWS, self.width))
        elif self.backend == "native":
            self._w = hdf5_min.Writer(path, DATASET, self.width, "f2", CHUNK_ROWS, attrs)
        else:
            raise ValueError(f"unknown store backend '{self.backend}'")

    def append(self, emb: np.ndarray) -> None:
        emb = np.asarray(emb)
        if emb.ndim != 2 or emb.shape[1] != self.width:
            raise ValueError(f"expected [n,{self.width}] embeddings, got {emb.shape}")
        block = emb.astype(np.float16, copy=False)  # the reference's dset[...] = float32 array cast (cbas.py:438)
        if self.backend == "h5py":
            n = self._d.shape[0]
            self._d.resize(n + len(block), axis=0)
            self._d[n:] = block
        else:
            self._w.append(block)
        self.rows += len(block)

    def flush(self) -> None:
        (self._f if self.backend == "h5py" else self._w).flush()

    def close(self) -> None:
        (self._f if self.backend == "h5py" else self._w).close()

    def abort(self) -> None:
        try:
            if self.backend == "h5py":
                self._f.close()
            else:
                self._w.abort()
        except Exception:
            pass


class EmbeddingReader:
    """Read an embedding file: `.shape`, `.attrs`, `read(start, stop)` -> float16 rows."""

    def __init__(self, path: str, backend: Optional[str] = None):
        self.backend = backend or ("h5py" if HAVE_H5PY else "native")
        self._f = h5py.File(path, "r") if self.backend == "h5py" else hdf5_min.File(path, "r")
        if DATASET not in self._f:
            self._f.close()
            raise KeyError(f"{path} has no '{DATASET}' dataset")
        self._d = self._f[DATASET]
        self.shape = tuple(self._d.shape)
        self.attrs = {k: (v.decode() if isinstance(v, bytes) else v) for k, v in dict(self._f.attrs).items()}

    def read(self, start: int, stop: int) -> np.ndarray:
        return np.asarray(self._d[start:stop])

    def close(self) -> None:
        self._f.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
        return False
