"""Generic batch plotting (reference ``generic_batch.py``): one ``generic_plot_spectrogram_set``
per item, saved as ``output_dir/<item>/generic.png``, resumable through ``run_batch``.

The GPU context is one per process and not thread-safe, so the workers are threads: the GPU
section of an item (collapse, bounds, raster) is serialised, the PNG DEFLATE of finished
figures runs concurrently on the other workers.
"""

from __future__ import annotations

import functools
import os
import threading
from collections.abc import Callable
from concurrent.futures import ThreadPoolExecutor
from typing import Any

from .batch_runner import run_batch
from .constants import PLOTTING_PROGRESS_JSON_PATH
from .figure import close_all_axes_and_clear
from .logging_utils import log_error
from .plotting import generic_plot_spectrogram_set

__all__ = ["generic_batch_plot"]

_gpu_lock = threading.Lock()


def generic_batch_plot(
    items,
    output_dir: str,
    build_datasets_fn: Callable[[Any], list[dict]],
    zoom_center_fn: Callable[[Any], float | None] | None = None,
    zoom_window_seconds: float | None = None,
    vertical_lines_fn: Callable[[Any], list[float] | None] | None = None,
    y_scale: str = "linear",
    z_scale: str = "linear",
    colormap: str = "viridis",
    cusp_marker_style: str = "both",
    cusp_marker_kwargs: dict | None = None,
    max_workers: int = 2,
    progress_json_path: str = PLOTTING_PROGRESS_JSON_PATH,
    ignore_progress_json: bool = False,
    flush_batch_size: int = 10,
    log_flush_batch_size: int | None = None,
    install_signal_handlers: bool = True,
) -> list[tuple[Any, str]]:
    """Plot every item (reference ``:15-129``).  Returns ``[(item, status)]`` with status in
    ``{'ok', 'no_data', 'error'}``."""
    os.makedirs(output_dir, exist_ok=True)

    def worker(item):
        try:
            datasets = build_datasets_fn(item)
            if not datasets:
                return (item, "no_data")
            center = zoom_center_fn(item) if zoom_center_fn else None
            vertical_lines = vertical_lines_fn(item) if vertical_lines_fn else None
            with _gpu_lock:
                fig, _canvas = generic_plot_spectrogram_set(
                    datasets, zoom_center=center, zoom_window_seconds=zoom_window_seconds, vertical_lines=vertical_lines,
                    y_scale=y_scale, z_scale=z_scale, colormap=colormap, cusp_marker_style=cusp_marker_style,
                    cusp_marker_kwargs=cusp_marker_kwargs, show=False,
                )
            if fig is not None:
                item_dir = os.path.join(output_dir, str(item))
                os.makedirs(item_dir, exist_ok=True)
                fig.savefig(os.path.join(item_dir, "generic.png"), dpi=150)
                close_all_axes_and_clear(fig)
            return (item, "ok")
        except Exception as exc:
            log_error(f"[GENERIC-FAIL] Item {item}: {exc}")
            return (item, "error")

    return run_batch(
        items,
        worker,
        functools.partial(ThreadPoolExecutor, max_workers=max_workers),
        progress_json_path=progress_json_path,
        ignore_progress_json=ignore_progress_json,
        flush_batch_size=flush_batch_size,
        log_flush_batch_size=log_flush_batch_size,
        install_signal_handlers=install_signal_handlers,
    )
