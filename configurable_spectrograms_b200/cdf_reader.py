"""Native CDF v3 ingest (``csg_cdf_*`` in libcsgpu, ``csrc/cdf.cpp``) behind a small Python class.

The reference reads every FAST file through ``cdflib`` -- four full ``varget`` calls per load, one of
them the data-sized ``pitch_angle`` variable of which 64 numbers are used (``CS/cdf_utils.py:247-253``).
Here only the records that are needed are decoded, straight into memory the caller provides (a pinned
staging slot of the batch driver), so a cube crosses host memory once on its way to HBM.
"""

from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

_NP_TYPES = {1: np.int8, 2: np.int16, 4: np.int32, 8: np.int64, 11: np.uint8, 12: np.uint16, 14: np.uint32,
             21: np.float32, 44: np.float32, 22: np.float64, 45: np.float64, 31: np.float64, 33: np.int64,
             41: np.int8, 51: np.uint8, 52: np.uint8}
_MAGIC = b"\xcd\xf3\x00\x01"


def is_cdf_v3(path: str) -> bool:
    """True when the file starts with the CDF v3 magic number (an empty marker file does not)."""
    try:
        with open(path, "rb") as f:
            return f.read(4) == _MAGIC
    except OSError:
        return False


class CdfFile:
    """One open CDF file.  ``with CdfFile(path) as cdf: cdf.read("data", out=pinned_view)``."""

    def __init__(self, path: str):
        self.lib = _lib.load_library()
        handle = C.c_void_p()
        if self.lib.csg_cdf_open(str(path).encode(), C.byref(handle)) != 0:
            raise _lib.CsgError(f"{path}: {self.lib.csg_cdf_last_error().decode()}")
        self.handle = handle
        self.path = str(path)

    def close(self):
        if self.handle:
            self.lib.csg_cdf_close(self.handle)
            self.handle = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _info(self, name, index=0):
        info = np.zeros(1, dtype=_lib.CDF_VAR)
        if self.lib.csg_cdf_var_info(self.handle, name.encode() if name is not None else None, index, info.ctypes.data) != 0:
            raise KeyError(self.lib.csg_cdf_last_error().decode())
        return info[0]

    def variables(self) -> list[str]:
        return [self._info(None, i)["name"].decode() for i in range(self.lib.csg_cdf_var_count(self.handle))]

    def shape(self, name: str) -> tuple[int, ...]:
        """``(records, *dims)`` -- what ``cdflib``'s ``varget`` would return for a record-varying variable."""
        info = self._info(name)
        return (int(info["n_records"]),) + tuple(int(d) for d in info["dims"][: info["n_dims"]])

    def dtype(self, name: str):
        info = self._info(name)
        dt = _NP_TYPES.get(int(info["data_type"]))
        if dt is None or info["elem_bytes"] == 0:
            raise TypeError(f"{name}: CDF data type {int(info['data_type'])} is not supported")
        return np.dtype(dt)

    def read(self, name: str, rec0: int = 0, n_rec: int | None = None, out: np.ndarray | None = None) -> np.ndarray:
        """Records ``[rec0, rec0 + n_rec)`` as a C-ordered array ``(n_rec, *dims)``; decoded into ``out``
        (same dtype, enough room, C-contiguous) when given."""
        shape = self.shape(name)
        dt = self.dtype(name)
        n_rec = shape[0] - rec0 if n_rec is None else int(n_rec)
        want = (n_rec,) + shape[1:]
        count = int(np.prod(want, dtype=np.int64))
        if out is None:
            out = np.empty(want, dtype=dt)
        else:
            if out.dtype != dt or not out.flags.c_contiguous or out.size < count:
                raise ValueError(f"{name}: `out` must be C-contiguous {dt} with room for {count} values")
            out = out.reshape(-1)[:count].reshape(want)
        if self.lib.csg_cdf_read(self.handle, name.encode(), int(rec0), n_rec, out.ctypes.data, out.nbytes) != 0:
            raise _lib.CsgError(f"{self.path}:{name}: {self.lib.csg_cdf_last_error().decode()}")
        return out


def read_fast_variables(path: str, data_alloc=None) -> dict[str, np.ndarray]:
    """The four variables of ``load_fast_cdf_dataset`` (``CS/cdf_utils.py:247-253``): ``time_unix`` and
    ``data`` in full, record 0 only of ``energy`` and ``pitch_angle`` (the reference takes
    ``energy[0, 0, :]`` and ``pitch_angle[0, :, 0]``).  ``data_alloc(shape, dtype)`` may hand out the
    memory the cube is decoded into (pinned staging)."""
    with CdfFile(path) as cdf:
        times = cdf.read("time_unix")
        shape, dt = cdf.shape("data"), cdf.dtype("data")
        out = data_alloc(shape, dt) if data_alloc is not None else None
        data = cdf.read("data", out=out)
        energy = cdf.read("energy", 0, 1)
        pitch = cdf.read("pitch_angle", 0, 1)
    return {"time_unix": times, "data": data, "energy": energy, "pitch_angle": pitch}
