"""Deterministic synthetic FAST ESA orbits (no network, no cdflib on the box).

Shapes follow the sample-orbit listing in the reference's
``FAST CDF variables.txt`` (``data (800,64,96)`` for ees/ies ``:88,:207``,
``(903,64,96)`` for eeb/ieb ``:326,:445``; ``energy (k,P',96)``;
``pitch_angle (T,64,96)``; ``time_unix (T,)``).  Only the slices the
reference actually consumes are kept full size: ``energy[0,0,:]`` and
``pitch_angle[0,:,0]`` (``cdf_utils.py:252-253``), so the side variables are
stored with a singleton time axis.

Files are written as ``<name>.cdf`` (empty marker, discovery globs ``*.cdf``,
``fast/orbit_discovery.py:155``) plus ``<name>.cdf.npz`` holding the four
variables; :func:`configurable_spectrograms_b200.cdf_utils.load_fast_cdf_dataset`
reads the side-car when cdflib is unavailable.
"""

from __future__ import annotations

import os
from datetime import datetime, timedelta, timezone

import numpy as np

N_ENERGY = 96
N_PITCH = 64
INSTRUMENTS = ("ees", "eeb", "ies", "ieb")
#: nominal record counts of the sample orbit (``FAST CDF variables.txt:88,207,326,445``)
NOMINAL_T = {"ees": 800, "ies": 800, "eeb": 903, "ieb": 903}
#: cube bytes of one nominal 4-instrument orbit, float32
ORBIT_CUBE_BYTES_F32 = sum(NOMINAL_T[i] * N_PITCH * N_ENERGY * 4 for i in INSTRUMENTS)


def energy_bins(n_energy: int = N_ENERGY, descending: bool = True) -> np.ndarray:
    """96 log-spaced energies 30 keV -> 4 eV (descending like the ESA sweep)."""
    e = np.geomspace(4.0, 30000.0, n_energy).astype(np.float32)
    return e[::-1].copy() if descending else e


def pitch_angle_bins(n_pitch: int = N_PITCH, quirks: bool = False) -> np.ndarray:
    """Pitch-angle bin centres in degrees.

    With ``quirks`` the grid is stretched so the first/last bins fall outside
    [0, 360] and one bin is NaN: the reference's "all (0, 360)" group then
    differs from the unmasked sum (``fast/plotting.py:124-127`` vs ``:278``).
    """
    if not quirks:
        return ((np.arange(n_pitch) + 0.5) * (360.0 / n_pitch)).astype(np.float32)
    pa = np.linspace(-4.0, 364.0, n_pitch).astype(np.float32)
    pa[n_pitch // 3] = np.nan
    # put bins exactly on closed-interval edges (30, 150, 210, 330)
    pa[5], pa[26], pa[37], pa[57] = 30.0, 150.0, 210.0, 330.0
    return pa


def make_cube(
    rng: np.random.Generator,
    n_time: int,
    n_pitch: int = N_PITCH,
    n_energy: int = N_ENERGY,
    dtype=np.float32,
    intensity: float = 3.0,
    integer_counts: bool = True,
    nan_fraction: float = 0.01,
    quirks: bool = False,
    cusp_window: tuple[int, int] | None = None,
) -> np.ndarray:
    """One (time, pitch-angle, energy) counts cube."""
    t = np.arange(n_time, dtype=np.float64)[:, None, None]
    p = np.arange(n_pitch, dtype=np.float64)[None, :, None]
    e = np.arange(n_energy, dtype=np.float64)[None, None, :]
    # high-energy channels (low index: the sweep is descending) are nearly empty so
    # the 99 %-coverage energy of the extrema pass lands below the 4000 eV cap
    cutoff = 1e-4 + 1.0 / (1.0 + np.exp(-(e - 0.30 * n_energy) / 0.8))
    lam = (
        intensity
        * cutoff
        * (
            0.15
            + np.exp(-(((e - 0.55 * n_energy) / (0.22 * n_energy)) ** 2))
            * (1.0 + 0.6 * np.cos(2 * np.pi * p / n_pitch))
            * (1.0 + 0.5 * np.sin(2 * np.pi * t / max(n_time, 1) * 3.0))
        )
    )
    if cusp_window is not None:
        lo, hi = cusp_window
        bump = np.zeros((n_time, 1, 1))
        bump[max(lo, 0) : max(hi, lo + 1)] = 4.0
        lam = lam * (1.0 + bump * np.exp(-(((e - 0.7 * n_energy) / (0.1 * n_energy)) ** 2)))
    if integer_counts:
        cube = rng.poisson(lam).astype(dtype)
    else:
        cube = rng.gamma(shape=2.0, scale=lam / 2.0 + 1e-3).astype(dtype)
    if nan_fraction > 0:
        cube[rng.random(cube.shape) < nan_fraction] = np.nan
    if quirks and n_time >= 8:
        cube[n_time // 5] = np.nan  # an all-NaN time row
        cube[:, :, n_energy - 3] = np.nan  # an all-NaN energy column
        cube[2, 3, 40] = -1e31  # CDF fill value: summed as-is by the reference
        cube[3, 7, 41] = np.inf
        cube[4, 9, 42] = -np.inf
        cube[5, 1, 43] = np.inf
        cube[5, 2, 43] = -np.inf  # inf + -inf -> NaN in the collapsed matrix
        cube[6, :, 44] = 0.0
        cube[7, 0, 45] = -0.0
    return cube


def make_times(n_time: int, start: float = 946684800.0, cadence: float = 2.5) -> np.ndarray:
    return start + cadence * np.arange(n_time, dtype=np.float64)


def make_file_arrays(
    rng: np.random.Generator,
    instrument: str,
    n_time: int | None = None,
    start: float = 946684800.0,
    dtype=np.float32,
    intensity: float | None = None,
    integer_counts: bool = True,
    nan_fraction: float = 0.01,
    quirks: bool = False,
    cusp_window: tuple[int, int] | None = None,
    stored_layout: str = "tpe",
) -> dict[str, np.ndarray]:
    """The four CDF variables of one instrument file.

    ``stored_layout='tep'`` stores ``data`` as (time, energy, pitch) so the
    loader's conditional transpose (``cdf_utils.py:254-255``) produces the
    non-contiguous (time, pitch, energy) *view* (summation layout B).
    """
    if n_time is None:
        n_time = NOMINAL_T[instrument]
    if intensity is None:
        intensity = 3.0 if instrument.startswith("e") else 0.3
    cadence = 2.5 if instrument.endswith("s") else 0.6
    cube = make_cube(
        rng,
        n_time,
        dtype=dtype,
        intensity=intensity,
        integer_counts=integer_counts,
        nan_fraction=nan_fraction,
        quirks=quirks,
        cusp_window=cusp_window,
    )
    if stored_layout == "tep":
        cube = np.ascontiguousarray(np.transpose(cube, (0, 2, 1)))
    energy = energy_bins()
    pa = pitch_angle_bins(quirks=quirks)
    if stored_layout == "tep":
        energy_full = np.broadcast_to(energy[None, None, :], (1, N_PITCH, N_ENERGY)).copy()
        pa_full = np.broadcast_to(pa[None, :, None], (1, N_PITCH, N_ENERGY)).copy()
    else:
        energy_full = np.broadcast_to(energy[None, None, :], (1, 1, N_ENERGY)).copy()
        pa_full = np.broadcast_to(pa[None, :, None], (1, N_PITCH, 1)).copy()
    return {
        "time_unix": make_times(n_time, start=start, cadence=cadence),
        "data": cube,
        "energy": energy_full,
        "pitch_angle": pa_full,
    }


def fast_filename(instrument: str, start_unix: float, orbit: int) -> str:
    """``fa_esa_l2_{inst}_{YYYYMMDDhhmmss}_{orbit}_v02.cdf`` (``fast/orbit_discovery.py:92-126``)."""
    stamp = datetime.fromtimestamp(start_unix, tz=timezone.utc).strftime("%Y%m%d%H%M%S")
    return f"fa_esa_l2_{instrument}_{stamp}_{orbit}_v02.cdf"


CUSP_CSV_COLUMNS = (
    "Orbit Number\tFolder Path\torb File\torb min Index\torb Max Index\t"
    "eeb\teeb File\teeb min Index\teeb Max Index\tees\tees File\tees min Index\tees Max Index\t"
    "ieb\tieb File\tieb min Index\tieb Max Index\ties\ties File\ties min Index\ties Max Index"
)


def write_fast_directory(
    root: str,
    n_orbits: int,
    seed: int = 3,
    first_orbit: int = 13000,
    n_time: dict[str, int] | None = None,
    jitter_time: bool = True,
    storm_orbits: tuple[int, ...] = (1,),
    missing: dict[int, str] | None = None,
    cusp_every: int = 3,
    dtype=np.float32,
    integer_counts: bool = True,
    quirks_every: int = 0,
    instruments: tuple[str, ...] = INSTRUMENTS,
    csv_path: str | None = None,
) -> dict:
    """Write a synthetic FAST data tree + cusp TSV; return a manifest.

    ``storm_orbits`` (indices into the orbit sequence) get x8 intensity so the
    reference's running-max-over-prefix z extrema differs from the final-pool
    percentile (hard part 1).
    """
    rng = np.random.default_rng(seed)
    base_t = dict(NOMINAL_T if n_time is None else n_time)
    missing = missing or {}
    manifest = {"root": root, "orbits": {}, "csv": None}
    rows = [CUSP_CSV_COLUMNS]
    t0 = datetime(2000, 1, 1, tzinfo=timezone.utc)
    for k in range(n_orbits):
        orbit = first_orbit + k
        start_dt = t0 + timedelta(minutes=133 * k)
        start = start_dt.timestamp()
        folder = os.path.join(root, f"{start_dt.year:04d}", f"{start_dt.month:02d}")
        os.makedirs(folder, exist_ok=True)
        files = {}
        csv_idx = {}
        has_cusp = cusp_every > 0 and (k % cusp_every == 0)
        for inst in instruments:
            if missing.get(k) == inst:
                continue
            T = base_t[inst]
            if jitter_time:
                T = int(T + rng.integers(-T // 8, T // 8 + 1))
            window = None
            if has_cusp:
                lo = int(T * 0.4)
                hi = lo + max(2, int(T * (0.06 if inst.endswith("s") else 0.25)))
                window = (lo, min(hi, T - 1))
                csv_idx[inst] = window
            scale = 8.0 if k in storm_orbits else 1.0
            arrays = make_file_arrays(
                rng,
                inst,
                n_time=T,
                start=start,
                dtype=dtype,
                intensity=scale * (3.0 if inst.startswith("e") else 0.3),
                integer_counts=integer_counts,
                quirks=(quirks_every > 0 and k % quirks_every == 0),
                cusp_window=window,
            )
            name = fast_filename(inst, start, orbit)
            path = os.path.join(folder, name)
            open(path, "wb").close()
            np.savez(path + ".npz", **arrays)
            files[inst] = path
        manifest["orbits"][orbit] = files

        def _cell(inst, j):
            return str(csv_idx[inst][j]) if inst in csv_idx else ""

        rows.append(
            "\t".join(
                [str(orbit), folder, f"fa_k0_orb_{orbit}_v01.cdf", "0", "0"]
                + [
                    x
                    for inst in ("eeb", "ees", "ieb", "ies")
                    for x in ("True", os.path.basename(files.get(inst, "")), _cell(inst, 0), _cell(inst, 1))
                ]
            )
        )
    if csv_path is None:
        csv_path = os.path.join(root, "FAST_Cusp_Indices.csv")
    with open(csv_path, "w") as f:
        f.write("\n".join(rows) + "\n")
    manifest["csv"] = csv_path
    return manifest
