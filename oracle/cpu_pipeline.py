"""TEST INFRASTRUCTURE / CPU BASELINE ONLY -- never imported by the product path.

The reference's numeric work for one batch step, restated on numpy exactly as the
reference structures it (``CS/`` = ``src/configurable_spectrograms/``):

* ``CS/fast/extrema.py:245-300``  serial pre-pass: per (orbit, instrument) nansum, positive
  pool, re-concatenate + nanpercentile after every step (quadratic);
* ``CS/fast/batch_directory.py:237-243`` every orbit submitted twice when extrema are on;
* ``CS/fast/process_orbit.py:148-253`` per submission: 4 instruments x {given, raw}
  pitch-angle grids + {given, raw} instrument grids;
* ``CS/fast/plotting.py:121-150`` 4 gathers + nansums + percentiles per pitch-angle grid,
  ``CS/plotting.py:612-656`` two ``make_spectrogram`` per row (each re-collapsing),
* the norm + LUT of ``imshow`` at cell resolution (``oracle/restate.py`` R9).

Used by ``bench.py`` for the reported ``cpu_baseline`` and the ``--impl reference`` arm
(the real reference cannot travel to the GPU box: it needs cdflib + matplotlib), fanned
out over a fork ``ProcessPoolExecutor`` like the reference does.  matplotlib's Agg
rasterisation and PNG encoding are NOT included, so this baseline is optimistic for
the reference.
"""

from __future__ import annotations

import os
import time
from concurrent.futures import ProcessPoolExecutor

import numpy as np

from oracle import restate as R

PA_GROUPS = (
    [(0.0, 360.0)],
    [(0.0, 30.0), (330.0, 360.0)],
    [(150.0, 210.0)],
    [(40.0, 140.0), (210.0, 330.0)],
)
ORDER = ("ees", "eeb", "ies", "ieb")
# float32 log10: numpy's native routine (what the reference executes; the timing baseline) or the
# correctly rounded definition the parity tests grade against (oracle/restate.py log10_like)
NATIVE_LOG = True


def _group_mask(pa, ranges):
    m = np.zeros_like(pa, dtype=bool)
    with np.errstate(invalid="ignore"):
        for lo, hi in ranges:
            m |= (pa >= lo) & (pa <= hi)
    return m


def _zoom(lines, minutes):
    if not lines:
        return None
    if len(lines) == 1:
        return lines[0], minutes * 60
    return 0.5 * (lines[0] + lines[1]), max(minutes * 60, abs(lines[1] - lines[0]) * 1.5)


def pitch_angle_grid(ds, lines, z_scale, lut, y_min=None, y_max=None, z_min=None, z_max=None, minutes=6.25):
    """Numeric work of ``FAST_plot_pitch_angle_grid``: returns the panels' index planes."""
    times, data, energy, pa = ds["times"], ds["data"], ds["energy"], ds["pitch_angle"]
    y_lo = 0 if y_min is None else y_min
    y_hi = 4000 if y_max is None else y_max
    valid = (energy >= y_lo) & (energy <= y_hi)
    rows = []
    for ranges in PA_GROUPS:
        pa_data = data[:, _group_mask(pa, ranges), :]
        with np.errstate(invalid="ignore", over="ignore"):
            full = np.nansum(pa_data, axis=1)
        full = full[:, ~np.all(np.isnan(full), axis=0) & valid].T
        if full.size == 0:
            continue
        with np.errstate(invalid="ignore"):
            vmin, vmax = R.compute_percentile_bounds(full, 1, 99, z_min, z_max)
        rows.append((pa_data, vmin, vmax))
    return _multirow(times, energy, rows, lines, z_scale, lut, z_min, z_max, minutes)


def instrument_grid(datasets, lines, z_scale, lut, extrema=None, y_scale="linear", minutes=6.25):
    rows = []
    times0 = energy0 = None
    per_row_axes = []
    for inst in ORDER:
        ds = datasets.get(inst)
        if ds is None:
            continue
        times, data, energy = ds["times"], ds["data"], ds["energy"]
        if isinstance(extrema, dict):
            kp = f"{inst}_{y_scale}_{z_scale}"
            y_lo, y_hi = extrema.get(f"{kp}_y_min", 0), extrema.get(f"{kp}_y_max", 4000)
            rz = (extrema.get(f"{kp}_z_min"), extrema.get(f"{kp}_z_max"))
        else:
            y_lo, y_hi, rz = 0, 4000, (None, None)
        with np.errstate(invalid="ignore", over="ignore"):
            full = np.nansum(data, axis=1)
        full = full[:, ~np.all(np.isnan(full), axis=0) & ((energy >= y_lo) & (energy <= y_hi))].T
        if full.size == 0:
            continue
        with np.errstate(invalid="ignore"):
            vmin, vmax = R.compute_percentile_bounds(full, 1, 99, *rz)
        rows.append((data, vmin, vmax))
        per_row_axes.append((times, energy))
    out = []
    zoom = _zoom(lines, minutes)
    need = False
    if zoom is not None:
        for (data, _, _), (times, _) in zip(rows, per_row_axes):
            m = (times >= zoom[0] - zoom[1] / 2) & (times <= zoom[0] + zoom[1] / 2)
            if np.any(~np.isnan(data[m])):
                need = True
                break
    for (data, vmin, vmax), (times, energy) in zip(rows, per_row_axes):
        out.append(_one_panel(times, energy, data, z_scale, vmin, vmax, lut, x_min=times[0], x_max=times[-1]))
        if need:
            out.append(_one_panel(times, energy, data, z_scale, vmin, vmax, lut, center=zoom[0], window=zoom[1]))
    return out


def _one_panel(times, energy, cube, z_scale, vmin, vmax, lut, **kw):
    with np.errstate(all="ignore"):
        p = R.panel(times, energy, cube, z_scale=z_scale, z_min=vmin, z_max=vmax, **kw)
        if p is None:
            return None
        try:
            idx, rgba = R.rasterise(p, lut, native_log=NATIVE_LOG)
        except ValueError:
            return None
    return idx, rgba


def _multirow(times, energy, rows, lines, z_scale, lut, z_min, z_max, minutes):
    zoom = _zoom(lines, minutes)
    need = False
    if zoom is not None:
        m = (times >= zoom[0] - zoom[1] / 2) & (times <= zoom[0] + zoom[1] / 2)
        for pa_data, _, _ in rows:
            if np.any(~np.isnan(pa_data[m])):
                need = True
                break
    out = []
    for pa_data, vmin, vmax in rows:
        lo = vmin if z_min is None else z_min
        hi = vmax if z_max is None else z_max
        out.append(_one_panel(times, energy, pa_data, z_scale, lo, hi, lut, x_min=times[0], x_max=times[-1]))
        if need:
            out.append(_one_panel(times, energy, pa_data, z_scale, lo, hi, lut, center=zoom[0], window=zoom[1]))
    return out


_SHARED: dict = {}  # inherited by fork workers: the reference's workers load files themselves,
#                     so the cubes must not be pickled through the pool either


def _process_indexed(job):
    index, with_extrema = job
    _o, datasets, lines = _SHARED["orbits"][index]
    extrema = _SHARED["state"] if with_extrema else None
    return process_orbit((datasets, lines, _SHARED["y"], _SHARED["z"], extrema, _SHARED["lut"]))


def process_orbit(args):
    """Numeric work of one ``FAST_process_single_orbit`` submission."""
    datasets, lines, y_scale, z_scale, extrema, lut = args
    n_panels = 0
    for inst in ORDER:
        ds = datasets.get(inst)
        if ds is None:
            continue
        ov = R.extrema_overrides(extrema, inst, y_scale, z_scale)
        for kw in (dict(y_min=ov[0], y_max=ov[1], z_min=ov[2], z_max=ov[3]), {}):
            n_panels += sum(p is not None for p in pitch_angle_grid(ds, lines.get(inst), z_scale, lut, **kw))
    first_lines = next((lines.get(i) for i in ORDER if i in datasets), None)
    for ge in (extrema, None):
        n_panels += sum(p is not None for p in instrument_grid(datasets, first_lines, z_scale, lut, ge, y_scale))
    return n_panels


def run_step(orbits, y_scale="linear", z_scale="log", max_percentile=99.0, lut=None, workers=None):
    """One batch step over ``orbits`` = [(orbit, datasets, lines)]: extrema pre-pass (serial, as in
    the reference) then both submissions of every orbit over a fork process pool.

    Returns ``(seconds, extrema_state, n_panels)``.
    """
    if lut is None:
        lut = R.lut_with_extremes(np.random.default_rng(0).integers(0, 256, (256, 4), dtype=np.uint8))
    workers = workers or os.cpu_count() or 1
    t0 = time.perf_counter()
    files = [(o, {i: (ds["energy"], ds["data"]) for i, ds in dsets.items()}) for o, dsets, _ in orbits]
    state = R.global_extrema(files, ORDER, y_scale, z_scale, state={}, max_percentile=max_percentile)
    _SHARED.update(orbits=orbits, state=state, y=y_scale, z=z_scale, lut=lut)
    jobs = [(k, flag) for k in range(len(orbits)) for flag in (False, True)]
    if workers > 1:
        import multiprocessing as mp

        with ProcessPoolExecutor(max_workers=workers, mp_context=mp.get_context("fork")) as pool:
            n = sum(pool.map(_process_indexed, jobs))
    else:
        n = sum(_process_indexed(j) for j in jobs)
    _SHARED.clear()
    return time.perf_counter() - t0, state, n
