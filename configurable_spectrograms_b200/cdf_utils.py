"""CDF access and cusp-boundary lookup (reference ``cdf_utils.py``).

CDF *decoding* is out of scope for the GPU path: with
``cdflib`` installed real files are read through it; otherwise a ``<file>.npz`` side-car
holding the four variables is used (synthetic data, ``synth.py``).
"""

from __future__ import annotations

import os
from pathlib import Path

import numpy as np

from .constants import CDF_DATA_DIRECTORY, CDF_VARIABLE_NAMES, FILTERED_ORBITS_CSV_PATH
from .logging_utils import log_error, log_message

filtered_orbits_cache: dict = {}
orbit_column_cache: dict = {}
cdf_type_cache: dict = {}

INSTRUMENT_TAGS = ("ees", "eeb", "ies", "ieb")


def _read_variables(cdf_path: str, names, data_alloc=None) -> list[np.ndarray]:
    """The named variables of a file: a binary CDF v3 goes through the native reader (``cdf_reader``:
    ``energy`` / ``pitch_angle`` are cut to record 0, all the loader uses), an empty marker file with a
    ``<file>.npz`` side-car through numpy (synthetic data, ``synth.py``), anything else through cdflib."""
    from . import cdf_reader

    if cdf_reader.is_cdf_v3(str(cdf_path)):
        if tuple(names) == tuple(CDF_VARIABLE_NAMES):
            got = cdf_reader.read_fast_variables(str(cdf_path), data_alloc)
            return [got[n] for n in names]
        with cdf_reader.CdfFile(str(cdf_path)) as cdf:
            return [cdf.read(n) for n in names]
    side_car = str(cdf_path) + ".npz"
    if os.path.exists(side_car):
        return _read_npz(side_car, names, data_alloc)
    try:
        import cdflib
    except ImportError as exc:
        raise ImportError(
            f"cdflib is not installed and no side-car '{side_car}' exists: cannot read {cdf_path}"
        ) from exc
    with cdflib.CDF(cdf_path) as cdf:
        return [np.asarray(cdf.varget(n)) for n in names]


def _read_npz(side_car: str, names, data_alloc=None) -> list[np.ndarray]:
    """Members of an ``np.savez`` archive.  With ``data_alloc`` the (uncompressed, C-ordered) ``data``
    member is read from the file straight into the memory the allocator hands out -- one pass over the
    bytes, no intermediate array, no CRC pass -- which is what staging into pinned slots wants; every
    other case goes through ``np.load``."""
    if data_alloc is not None and "data" in names:
        import zipfile
        from numpy.lib import format as npy_format

        try:
            with zipfile.ZipFile(side_car) as zf:
                info = zf.getinfo("data.npy")
                direct = info.compress_type == zipfile.ZIP_STORED
            if direct:
                with open(side_car, "rb") as f:
                    f.seek(info.header_offset)
                    local = f.read(30)
                    name_len, extra_len = int.from_bytes(local[26:28], "little"), int.from_bytes(local[28:30], "little")
                    f.seek(info.header_offset + 30 + name_len + extra_len)
                    version = npy_format.read_magic(f)
                    shape, fortran, dtype = (npy_format.read_array_header_1_0(f) if version == (1, 0)
                                             else npy_format.read_array_header_2_0(f))
                    if not fortran and not dtype.hasobject:
                        out = data_alloc(shape, dtype)
                        if out is not None and out.flags.c_contiguous:
                            flat = out.reshape(-1).view(np.uint8)
                            got = f.readinto(memoryview(flat))
                            if got != flat.nbytes:
                                raise OSError(f"{side_car}: data.npy is truncated ({got} of {flat.nbytes} bytes)")
                            with np.load(side_car) as z:
                                return [out if n == "data" else np.asarray(z[n]) for n in names]
        except (KeyError, ValueError, zipfile.BadZipFile):
            pass  # not the layout np.savez writes: the generic path below decides
    with np.load(side_car) as z:
        return [np.asarray(z[n]) for n in names]


def load_filtered_orbits(csv_path: str = FILTERED_ORBITS_CSV_PATH):
    """Tab-separated cusp-index table, cached per path; ``None`` when unreadable (``:26-52``)."""
    if csv_path in filtered_orbits_cache:
        return filtered_orbits_cache[csv_path]
    import pandas as pd

    try:
        frame = pd.read_csv(csv_path, sep="\t")
    except OSError as exc:
        log_error(f"Error loading CSV {csv_path}: {exc}")
        return None
    filtered_orbits_cache[csv_path] = frame
    return frame


def get_timestamps_for_orbit(filtered_orbits_dataframe, orbit_number, instrument_type, time_unix_array) -> list[float]:
    """Cusp boundary timestamps of an orbit: one value for a degenerate index pair, two
    otherwise, ``[]`` when anything is missing (``:55-123``).

    >>> import pandas as pd, numpy as np
    >>> orbits = pd.DataFrame({"orbit": [42], "ees min index": [1], "ees max index": [3]})
    >>> get_timestamps_for_orbit(orbits, 42, "ees", np.array([100.0, 200.0, 300.0, 400.0]))
    [200.0, 400.0]
    >>> get_timestamps_for_orbit(orbits, 99, "ees", np.array([100.0, 200.0, 300.0, 400.0]))
    []
    """
    frame = filtered_orbits_dataframe
    if frame is None or instrument_type is None or time_unix_array is None:
        return []
    key = (id(frame), instrument_type)
    cached = orbit_column_cache.get(key)
    if cached is None or cached[0] is not frame:  # (the entry keeps its frame alive, so an id is never re-used under it)
        lowered = {c: c.lower() for c in frame.columns}
        orbit_col = next(c for c, l in lowered.items() if "orbit" in l)
        lo_col = next(c for c, l in lowered.items() if instrument_type in l and "min index" in l)
        hi_col = next(c for c, l in lowered.items() if instrument_type in l and "max index" in l)
        # first row per orbit number, like ``frame[frame[orbit] == n].iloc[0]`` -- looked up once per file of a
        # directory run, where the boolean-mask filter of the whole table cost milliseconds per call
        first_rows: dict = {}
        for number, lo_value, hi_value in zip(frame[orbit_col].tolist(), frame[lo_col].tolist(), frame[hi_col].tolist()):
            try:
                first_rows.setdefault(number, (lo_value, hi_value))
            except TypeError:  # an unhashable cell can never equal an orbit number
                pass
        if len(orbit_column_cache) > 64:
            orbit_column_cache.clear()
        cached = orbit_column_cache[key] = (frame, first_rows)
    first_rows = cached[1]
    try:
        hit = first_rows.get(orbit_number)
    except TypeError:
        hit = None
    if hit is None:
        return []
    try:
        lo = int(hit[0])
        hi = int(hit[1])
    except (TypeError, ValueError):
        log_message("[WARN] Non-integer indices found in orbit row, using 0.")
        return []
    last = len(time_unix_array) - 1
    lo = max(0, min(lo, last))
    hi = max(0, min(hi, last))
    if lo == hi:
        return [float(time_unix_array[lo])]
    return [float(time_unix_array[lo]), float(time_unix_array[hi])]


def get_cdf_file_type(cdf_file_path: str) -> str | None:
    """``'ees'|'eeb'|'ies'|'ieb'``, ``'orb'`` for ephemeris files, else ``None`` (``:126-154``).

    >>> get_cdf_file_type("fa_esa_l2_eeb_20000101001737_13312_v02.cdf")
    'eeb'
    >>> get_cdf_file_type("fa_k0_orb_13312_v01.cdf")
    'orb'
    """
    lowered = cdf_file_path.lower()
    if "_orb_" in lowered:
        return "orb"
    for tag in INSTRUMENT_TAGS:
        if f"_{tag}_" in lowered:
            return tag
    log_error(f"Unknown CDF file type for path: {cdf_file_path}")
    return None


def get_variable_shape(cdf_path: str, variable_name: str):
    kind = cdf_type_cache.get(cdf_path)
    if kind is None:
        kind = cdf_type_cache[cdf_path] = get_cdf_file_type(cdf_path)
    if kind is None or kind == "orb":
        return None
    try:
        (value,) = _read_variables(cdf_path, [variable_name])
        return value.shape if isinstance(value, np.ndarray) else None
    except Exception as exc:
        log_error(f"Error reading {cdf_path} for variable {variable_name}: {exc}")
        return None


def get_cdf_var_shapes(cdf_folder_path: str = CDF_DATA_DIRECTORY, variable_names=CDF_VARIABLE_NAMES):
    paths = [str(p) for p in Path(cdf_folder_path).rglob("*.[cC][dD][fF]")]
    return {name: [get_variable_shape(p, name) for p in paths] for name in variable_names}


def load_fast_cdf_dataset(cdf_path: str, variable_names=tuple(CDF_VARIABLE_NAMES), data_alloc=None) -> dict[str, np.ndarray]:
    """``{'times','data','energy','pitch_angle'}`` with 1-D bin arrays and ``data`` as a
    (time, pitch, energy) array or transposed *view* (``:222-256``) -- the view is kept
    un-copied because numpy's summation order, reproduced on the GPU, depends on it.

    ``data_alloc(shape, dtype)`` (batch driver): memory the cube is decoded into -- a pinned staging
    slot -- when the file goes through the native reader."""
    times, data, energy_full, pitch_full = _read_variables(cdf_path, variable_names, data_alloc)
    energy = energy_full[0, 0, :] if energy_full.ndim == 3 else energy_full
    pitch = pitch_full[0, :, 0] if pitch_full.ndim == 3 else pitch_full
    if data.shape[1] == len(energy) and data.shape[2] == len(pitch):
        data = np.transpose(data, (0, 2, 1))
    return {"times": times, "data": data, "energy": energy, "pitch_angle": pitch}
