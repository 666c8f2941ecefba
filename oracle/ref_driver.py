"""TEST INFRASTRUCTURE / CPU BASELINE ONLY -- never imported by the product path.

Drives the UNMODIFIED reference (``oracle/_ref/configurable_spectrograms``, copied from the
reference checkout by ``oracle/make_ref.sh``) through its own public batch entry point,
``FAST_plot_spectrograms_directory`` (``CS/fast/batch_directory.py:32``): serial global-extrema
pre-pass, then a fork ``ProcessPoolExecutor`` over ``FAST_process_single_orbit`` (``:337-343``),
every orbit submitted twice (``:237-243``), progress / extrema JSON and PNG tree included.

The two third-party packages the reference imports and this image lacks are served by
``oracle/stubs.py``: ``cdflib`` reads ``<file>.cdf.npz`` side-cars, ``matplotlib`` is the
rendering stand-in described there (norm + LUT restated in numpy, PNG through Pillow).  Real
Agg rendering (resampling to 4800 x 2400, text, colourbars) is NOT included, so the timing is
optimistic for the reference.

Used by ``bench.py --impl reference`` / ``cpu_baseline`` / ``api_e2e`` and by
``tests/test_oracle.py`` (when ``oracle/_ref`` exists).
"""

from __future__ import annotations

import os
import shutil
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "configurable_spectrograms", "fast", "batch_directory.py"))


def install(render: str | None = "cell"):
    """Stubs in ``sys.modules`` + ``oracle/_ref`` first on ``sys.path``; returns the reference package."""
    if not available():
        raise RuntimeError("oracle/_ref is missing: run `sh oracle/make_ref.sh` where /root/reference exists")
    from oracle import stubs

    stubs.install(reference_src=REF_DIR)
    stubs.set_render(render)
    import configurable_spectrograms  # noqa: F401  (the reference, unmodified)

    src = os.path.dirname(os.path.abspath(configurable_spectrograms.__file__))
    if os.path.dirname(src) != REF_DIR:
        raise RuntimeError(f"configurable_spectrograms resolved to {src}, not to oracle/_ref")
    return configurable_spectrograms


def prepare_directory(work: str, n_orbits: int, seed: int = 4) -> dict:
    """A synthetic FAST tree of ``n_orbits`` nominal-shape orbits under ``work/FAST_data`` (+ the cusp
    TSV the reference reads from the CWD): the workload of ``bench.py`` -- 4 instruments, a cusp
    window on every third orbit, storm orbits early in the sequence."""
    root = os.path.dirname(HERE)
    if root not in sys.path:
        sys.path.insert(0, root)
    from configurable_spectrograms_b200 import synth

    os.makedirs(work, exist_ok=True)
    man = synth.write_fast_directory(os.path.join(work, "FAST_data"), n_orbits, seed=seed, jitter_time=False,
                                     storm_orbits=(1, 2, 5), cusp_every=3)
    shutil.copy(man["csv"], os.path.join(work, "FAST_Cusp_Indices.csv"))
    return man


def run_directory(work: str, workers: int | None = None, render: str | None = "cell", colormap: str = "turbo",
                  y_scale: str = "linear", z_scale: str = "log", max_percentile: float = 99.0) -> dict:
    """One whole batch step of the reference over ``work/FAST_data`` (fresh progress / extrema JSON and
    output tree every call).  Returns ``{"seconds", "results", "pngs", "png_bytes", "workers"}``."""
    install(render)
    import configurable_spectrograms.cdf_utils as cu
    from configurable_spectrograms.fast.batch_directory import FAST_plot_spectrograms_directory

    workers = workers or os.cpu_count() or 1
    cwd = os.getcwd()
    os.chdir(work)
    try:
        for name in ("progress.json", "FAST_calculated_extrema.json"):
            if os.path.exists(name):
                os.remove(name)
        shutil.rmtree("FAST_plots_ref", ignore_errors=True)
        cu.filtered_orbits_cache.clear()
        t0 = time.perf_counter()
        results = FAST_plot_spectrograms_directory(
            "./FAST_data", output_base="./FAST_plots_ref/", y_scale=y_scale, z_scale=z_scale, colormap=colormap,
            max_processing_percentile=max_percentile, max_workers=workers, progress_json_path="./progress.json",
            verbose=False, use_tqdm=False,
        )
        seconds = time.perf_counter() - t0
        pngs, nbytes = 0, 0
        for dirpath, _dirs, files in os.walk("./FAST_plots_ref"):
            for fn in files:
                pngs += 1
                nbytes += os.path.getsize(os.path.join(dirpath, fn))
    finally:
        os.chdir(cwd)
    return {"seconds": seconds, "results": results, "pngs": pngs, "png_bytes": nbytes, "workers": workers}
