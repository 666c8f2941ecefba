"""Batch driver: every orbit of a directory through the GPU path (reference ``fast/batch_directory.py``).

Same signature, progress-JSON schema, output tree and return value as the reference's
``FAST_plot_spectrograms_directory``.  What changes is the execution model: the reference fans
orbits out to a ``ProcessPoolExecutor`` (one orbit per task, ten figures each re-loading and
re-collapsing the files); here the orbits of this rank are ONE shard resident in HBM:

  K1  every cube collapsed once (all pitch-angle groups + total + zoom flags)
  K2b the global-extrema pre-pass from the same collapsed matrices (``compute_global_extrema``
      with its resumable JSON cache; NCCL exchange when several ranks share the directory)
  K2a 1st / 99th percentiles and safe_vmin of every panel of every figure, one launch
  K3  every panel rasterised, one launch

after which host threads compose the figures (labels, ticks, cusp markers) and DEFLATE the
PNGs.  ``max_workers`` sizes that thread pool.  Multi-GPU: launch one process per GPU with
``torchrun`` -- ranks take contiguous blocks of the ascending orbit sequence.
"""

from __future__ import annotations

import json
import os
import signal
from concurrent.futures import ThreadPoolExecutor
from typing import Any

import numpy as np

from .. import _lib
from ..cdf_utils import get_cdf_file_type, get_timestamps_for_orbit, load_fast_cdf_dataset, load_filtered_orbits
from ..colormaps import get_lut
from ..constants import DEFAULT_ZOOM_WINDOW_MINUTES
from ..figure import close_all_axes_and_clear
from ..logging_utils import configure_log_batch, flush_log_buffer, log_exception
from .constants import DEFAULT_INSTRUMENT_ORDER, FAST_CDF_DATA_FOLDER_PATH, FAST_OUTPUT_BASE, FAST_PLOTTING_PROGRESS_JSON
from .extrema import compute_global_extrema
from .orbit_discovery import _add_to_orbit_list, _classify_error_reason, _parse_year_month, discover_orbit_files
from .pipeline import BatchStep, ShardPlan
from .plotting import figure_from_spec
from .process_orbit import figure_filename

_INSTRUMENT_KEYS = DEFAULT_INSTRUMENT_ORDER

__all__ = ["FAST_plot_spectrograms_directory"]


def _rank_world():
    try:
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size(), dist
    except ImportError:
        pass
    return 0, 1, None


def FAST_plot_spectrograms_directory(
    directory_path: str = FAST_CDF_DATA_FOLDER_PATH,
    output_base: str = FAST_OUTPUT_BASE,
    y_scale: str = "linear",
    z_scale: str = "log",
    zoom_duration_minutes: float = DEFAULT_ZOOM_WINDOW_MINUTES,
    instrument_order: tuple[str, ...] = _INSTRUMENT_KEYS,
    verbose: bool = True,
    progress_json_path: str | None = FAST_PLOTTING_PROGRESS_JSON,
    ignore_progress_json: bool = False,
    use_tqdm: bool | None = None,
    colormap: str = "viridis",
    cusp_marker_style: str = "both",
    cusp_marker_kwargs: dict | None = None,
    max_workers: int = 4,
    orbit_timeout_seconds: int | float = 60,
    instrument_timeout_seconds: int | float = 30,
    retry_timeouts: bool = True,
    flush_batch_size: int = 10,
    log_flush_batch_size: int | None = None,
    max_processing_percentile: float | None = None,
    override_plots: bool = True,
) -> list[dict[str, Any]]:
    """Plot every orbit under ``directory_path`` (reference ``:32-433``).

    Returns one result dict per submission (``{"orbit", "status", "errors"}``); like the
    reference every orbit is submitted twice when ``max_processing_percentile`` is given (once
    without, once with the global extrema, ``:237-243``).  Progress keys:
    ``{y}_{z}_last_orbit``, ``{y}_{z}_error_plotting``, ``orbit_{y}_{z}_timed_out`` and the
    per-reason error lists (``:277-321``).  Raises ``KeyboardInterrupt`` on SIGINT / SIGTERM.

    ``orbit_timeout_seconds``, ``instrument_timeout_seconds`` and ``retry_timeouts`` are accepted for
    call compatibility and have nothing to act on: the reference times out and retries worker
    *processes* (``:344-420,455-492``); here an orbit is a slice of a few kernel launches, so the
    ``orbit_{y}_{z}_timed_out`` list is written and stays empty.
    """
    interrupted = {"flag": False}

    def _signal_handler(signum, frame):
        interrupted["flag"] = True
        log_exception(f"[INTERRUPT] Signal {signum} received. Requesting shutdown...", level="message")
        raise KeyboardInterrupt

    previous = {}
    try:
        for sig in (signal.SIGINT, signal.SIGTERM):
            previous[sig] = signal.signal(sig, _signal_handler)
    except (ValueError, OSError) as exc:
        log_exception("[WARN] Could not register signal handlers", exc, level="message")

    try:
        return _run_directory(
            directory_path, output_base, y_scale, z_scale, zoom_duration_minutes, tuple(instrument_order), verbose,
            progress_json_path, ignore_progress_json, colormap, cusp_marker_style, cusp_marker_kwargs, max_workers,
            flush_batch_size, log_flush_batch_size, max_processing_percentile, override_plots,
        )
    finally:
        for sig, handler in previous.items():
            try:
                signal.signal(sig, handler)
            except (ValueError, OSError):
                pass


def _run_directory(directory_path, output_base, y_scale, z_scale, zoom_duration_minutes, instrument_order, verbose,
                   progress_json_path, ignore_progress_json, colormap, cusp_marker_style, cusp_marker_kwargs, max_workers,
                   flush_batch_size, log_flush_batch_size, max_processing_percentile, override_plots):
    rank, world, dist = _rank_world()
    frame = load_filtered_orbits()
    configure_log_batch(log_flush_batch_size or flush_batch_size)

    orbit_to_instruments = discover_orbit_files(directory_path, instrument_order)
    sorted_orbits = sorted(orbit_to_instruments.items(), key=lambda kv: kv[0])
    total_orbits = len(sorted_orbits)

    # ---- resume state (reference :173-213)
    progress_key = f"{y_scale}_{z_scale}_last_orbit"
    error_key = f"{y_scale}_{z_scale}_error_plotting"
    timeout_key = f"orbit_{y_scale}_{z_scale}_timed_out"
    progress: dict[str, Any] = {}
    last_completed, error_orbits = None, set()
    if progress_json_path is not None and not ignore_progress_json:
        try:
            with open(progress_json_path) as f:
                progress = json.load(f)
            last_completed = progress.get(progress_key)
            error_orbits = set(progress.get(error_key, []))
        except (OSError, json.JSONDecodeError) as exc:
            log_exception(f"[ERROR] Failed to load progress JSON from {progress_json_path}. Starting fresh.", exc, level="error")
            progress = {}
    start_idx = 0
    if last_completed is not None:
        start_idx = next((i for i, (o, _) in enumerate(sorted_orbits) if o > last_completed), total_orbits)
        log_exception(f"[RESUME] Skipping {start_idx} orbits (up to orbit {last_completed}). "
                      f"{len(error_orbits)} error orbits will also be skipped.", level="message")
    else:
        log_exception("[RESUME] No previous progress found. Starting from the first orbit. "
                      f"{len(error_orbits)} error orbits will be skipped if present.", level="message")
    pending = [o for o, _ in sorted_orbits[start_idx:] if o not in error_orbits]
    flush_batch_size = max(1, flush_batch_size)

    # ---- this rank's contiguous block of the ascending orbit sequence, resident in one shard
    per = (total_orbits + world - 1) // world if total_orbits else 0
    lo, hi = rank * per, min(total_orbits, (rank + 1) * per)
    comm = None
    if world > 1:
        import torch

        from ..comm import TorchComm

        comm = TorchComm(dist, torch.device("cuda", torch.cuda.current_device()))
    need_extrema = max_processing_percentile is not None
    mine = sorted_orbits[lo:hi]
    load = [(o, files) for o, files in mine if need_extrema or o in set(pending)]
    results: list[dict[str, Any]] = []
    ctx = _lib.default_context(_current_device())
    shard = ShardPlan(ctx, y_scale, z_scale, zoom_duration_minutes, instrument_order=instrument_order)
    shard.first_orbit_index = lo
    load_errors: dict[int, list[str]] = {}
    loaded_orbits = []
    for orbit, files in load:
        datasets, lines = {}, {}
        for inst in DEFAULT_INSTRUMENT_ORDER:
            path = files.get(inst)
            if not path or inst not in instrument_order:
                continue
            try:
                detected = get_cdf_file_type(path)
                if detected is None or detected == "orb":
                    continue
                ds = load_fast_cdf_dataset(path)
                datasets[inst] = ds
                lines[inst] = get_timestamps_for_orbit(frame, orbit, detected, ds["times"])
            except Exception as exc:
                err = f"[FAIL] Plotting Orbit {orbit} pitch angle grid for {inst}"
                log_exception(err, exc, level="error")
                load_errors.setdefault(orbit, []).append(err)
        shard.add_orbit(orbit, datasets, lines)
        loaded_orbits.append(orbit)
    # ranks that loaded only part of the sequence still index it globally
    if not need_extrema:
        shard.first_orbit_index = 0
    shard.upload()
    shard.collapse()

    # ---- global extrema pre-pass (reference :159-171), from the collapsed matrices already in HBM
    global_extrema = None
    if need_extrema:
        global_extrema = compute_global_extrema(
            directory_path, y_scale, z_scale, instrument_order, compute_mins=False,
            max_percentile=float(max_processing_percentile), log_floor_cutoff=0.1, log_floor_value=-1.0,
            flush_batch_size=flush_batch_size, _shard=shard, _comm=comm,
        )

    # ---- every figure of this rank's pending orbits: K2a + K3; the figures are planned on host
    # threads from panel references, then composed and PNG-encoded on the device (K4): no raster
    # ever crosses PCIe uncompressed
    my_pending = [o for o in pending if o in set(loaded_orbits)]
    submissions = (False, True) if need_extrema else (False,)
    sequence = [(o, {i: True for i in files}) for o, files in sorted_orbits]
    step = BatchStep(shard, sequence, comm=comm, lut259=get_lut(colormap), plot_orbits=my_pending, submissions=submissions)
    if my_pending:
        step.run(state=global_extrema if global_extrema is not None else {}, collapse=False)
        step.finish()
    b = shard.batch
    norms = b.norms() if b.n_panels else None

    def render(orbit, with_extrema):
        """One submission of one orbit (= one FAST_process_single_orbit call of the reference)."""
        result: dict[str, Any] = {"orbit": orbit, "status": "ok", "errors": []}
        saves: list[tuple[str, Any]] = []  # (path, figure): encoded and written after the planning pass
        for err in load_errors.get(orbit, []):
            result["status"] = "error"
            result["errors"].append(err)
        files = orbit_to_instruments[orbit]
        first_path = next((files[k] for k in DEFAULT_INSTRUMENT_ORDER if k in files), None)
        year, month = _parse_year_month(first_path) if first_path else ("unknown", "unknown")
        out_dir = os.path.join(output_base, str(year), str(month), str(orbit))
        os.makedirs(out_dir, exist_ok=True)
        first, last = step.figure_ranges.get((orbit, with_extrema), (0, 0))
        for spec in shard.figures[first:last]:
            what = f"pitch angle grid for {spec.instrument}" if spec.kind == "pitch-angle" else "instrument grid"
            try:
                fig, _canvas = figure_from_spec(shard, spec, colormap, cusp_marker_style, cusp_marker_kwargs, norms=norms,
                                                device_rasters=True)
                if fig is None:
                    continue
                path = os.path.join(out_dir, figure_filename(spec, y_scale, z_scale, colormap))
                if not override_plots and os.path.exists(path):
                    log_exception(f"[SKIP] Plot already exists, skipping: {path}", level="message")
                    close_all_axes_and_clear(fig)
                else:
                    saves.append((path, fig))
            except Exception as exc:
                err = f"[FAIL] Plotting Orbit {orbit} {what}"
                log_exception(err, exc, level="error")
                result["status"] = "error"
                if err not in result["errors"]:
                    result["errors"].append(err)
        return result, saves

    def record(result, pdisk):
        orbit = result["orbit"]
        pdisk[progress_key] = orbit
        pdisk.setdefault(error_key, [])
        pdisk.setdefault(timeout_key, [])
        if result.get("status") == "error":
            _add_to_orbit_list(pdisk, error_key, orbit)
            for msg in result.get("errors") or []:
                reason = _classify_error_reason(msg)
                inst = next((c for c in _INSTRUMENT_KEYS if c in msg.lower()), "unknown")
                _add_to_orbit_list(pdisk, f"{inst}_{y_scale}_{z_scale}_error-{reason}", orbit)
                _add_to_orbit_list(pdisk, f"{y_scale}_{z_scale}_error-{reason}", orbit)

    jobs = [(o, flag) for o in my_pending for flag in submissions]
    pdisk = dict(progress)
    since_flush = 0
    with ThreadPoolExecutor(max_workers=max(1, int(max_workers))) as pool:
        planned = list(pool.map(lambda j: render(*j), jobs))
    rendered = [r for r, _s in planned]
    # the reference runs the submissions one after the other: a later one finds the earlier one's file
    # and skips it unless override_plots is set, in which case the later one wins
    by_path: dict[str, Any] = {}
    for _r, job_saves in planned:
        for path, fig in job_saves:
            if path in by_path and not override_plots:
                log_exception(f"[SKIP] Plot already exists, skipping: {path}", level="message")
                close_all_axes_and_clear(fig)
                continue
            if path in by_path:
                close_all_axes_and_clear(by_path[path])
            by_path[path] = fig
    saves = list(by_path.items())
    # ---- K4: compose + DEFLATE on the device, files written by a thread pool; progress is recorded
    # only once the orbit's PNGs are on disk
    if saves:
        from ..png import write_figures_device

        write_figures_device(ctx, b.d_rgba.ptr, saves, max_workers=max(1, int(max_workers)))
        for path, fig in saves:
            log_exception(f"[SAVED] {path}", level="message")
            close_all_axes_and_clear(fig)
    for result in rendered:
        results.append(result)
        if verbose:
            log_exception(f"[BATCH] Completed orbit {result['orbit']}: {result['status']}", level="message")
        if progress_json_path is not None and rank == 0:
            record(result, pdisk)
            since_flush += 1
            if since_flush >= flush_batch_size:
                _write_json(progress_json_path, pdisk)
                since_flush = 0

    if world > 1:  # every rank returns every result; rank 0 owns the progress file
        gathered: list = [None] * world
        dist.all_gather_object(gathered, results)
        results = [r for part in gathered for r in part]
        if rank == 0 and progress_json_path is not None:
            pdisk = dict(progress)
            for r in sorted(results, key=lambda r: r["orbit"]):
                record(r, pdisk)
    if progress_json_path is not None and rank == 0 and (results or os.path.exists(progress_json_path)):
        _write_json(progress_json_path, pdisk)
    flush_log_buffer(force=True)
    return results


def _write_json(path, data):
    try:
        with open(path, "w") as f:
            json.dump(data, f, indent=2)
    except OSError as exc:
        log_exception("[ERROR] Failed to write progress JSON", exc, level="error")


def _current_device() -> int:
    try:
        import torch

        if torch.cuda.is_available():
            return int(torch.cuda.current_device())
    except ImportError:
        pass
    return 0
