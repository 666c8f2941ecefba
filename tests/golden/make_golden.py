#!/usr/bin/env python
"""Generate golden fixtures by running the UNMODIFIED reference here.

Run in the build container only (``/root/reference`` does not exist on the GPU
box): ``python tests/golden/make_golden.py``.  The reference is imported from
``/root/reference/src`` under the two stubs in ``oracle/stubs.py`` (cdflib ->
``.npz`` side-cars, recording matplotlib) and driven through its own public
functions; every ``imshow`` input it produces is captured and written next to
the inputs that produced it, so the tests need neither the reference nor the
generator at run time.

Outputs (all under ``tests/golden/``):
  pa_grid.npz        FAST_plot_pitch_angle_grid, quirky ees file, cusp zoom, linear+log z
  inst_grid.npz      FAST_plot_instrument_grid, 4 instruments, raw + given extrema
  generic_set.npz    generic_plot_spectrogram_set (config 1), linear+log, f32+f64, layout-B view
  extrema_tree.npz   inputs of a 6-orbit tree (storm orbit early, one missing file)
  extrema_tree.json  compute_global_extrema results for the four (y,z) combos in CLI order
                     + the list of PNG names FAST_plot_spectrograms_directory saved
  doctests.json      the reference's own doctest pins on this path, evaluated here
"""

from __future__ import annotations

import json
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import stubs  # noqa: E402

stubs.install()

from configurable_spectrograms_b200 import synth  # noqa: E402


def _panels():
    out = []
    for r in stubs.RECORDED:
        out.append(r)
    return out


def _pack_panels(prefix, panels, store):
    store[f"{prefix}_n"] = np.array(len(panels))
    for i, r in enumerate(panels):
        store[f"{prefix}_{i}_matrix"] = r["matrix"]
        store[f"{prefix}_{i}_meta"] = np.array(
            [1.0 if r["mode"] == "log" else 0.0, float(r["vmin"]), float(r["vmax"])], dtype=np.float64
        )
        store[f"{prefix}_{i}_extent"] = np.array(r["extent"], dtype=np.float64)


def main():
    work = tempfile.mkdtemp(prefix="golden_")
    os.chdir(work)
    rng = np.random.default_rng(20)

    # ------------------------------------------------------------------ pa grid
    from configurable_spectrograms.cdf_utils import load_filtered_orbits
    from configurable_spectrograms.fast.plotting import FAST_plot_instrument_grid, FAST_plot_pitch_angle_grid

    os.makedirs("one/2000/01", exist_ok=True)
    T = 200
    arrays = synth.make_file_arrays(rng, "ees", n_time=T, quirks=True, cusp_window=(80, 92))
    path = os.path.join("one/2000/01", synth.fast_filename("ees", arrays["time_unix"][0], 777))
    open(path, "wb").close()
    np.savez(path + ".npz", **arrays)
    with open("FAST_Cusp_Indices.csv", "w") as f:
        f.write(synth.CUSP_CSV_COLUMNS + "\n")
        f.write("\t".join(["777", "x", "orb", "0", "0", "True", "f", "", "", "True", "f", "80", "92"] + [""] * 8) + "\n")
    df = load_filtered_orbits()
    store = {f"in_{k}": v for k, v in arrays.items()}
    store["cusp_idx"] = np.array([80, 92])
    for zs in ("linear", "log"):
        stubs.reset_recording()
        fig, _ = FAST_plot_pitch_angle_grid(
            path, filtered_orbits_df=df, orbit_number=777, scale_function_z=zs, show=False, colormap="turbo"
        )
        assert fig is not None
        _pack_panels(f"raw_{zs}", _panels(), store)
        stubs.reset_recording()
        fig, _ = FAST_plot_pitch_angle_grid(
            path,
            filtered_orbits_df=df,
            orbit_number=777,
            scale_function_z=zs,
            show=False,
            colormap="turbo",
            y_min=0.0,
            y_max=2900.0,
            z_min=0.0,
            z_max=460.0,
        )
        _pack_panels(f"given_{zs}", _panels(), store)
    np.savez_compressed(os.path.join(HERE, "pa_grid.npz"), **store)

    # ---------------------------------------------------------------- inst grid
    store = {}
    files = {}
    for inst, T in (("ees", 120), ("eeb", 150), ("ies", 120), ("ieb", 150)):
        a = synth.make_file_arrays(rng, inst, n_time=T, quirks=(inst == "ies"), cusp_window=(40, 60))
        p = os.path.join("one/2000/01", synth.fast_filename(inst, a["time_unix"][0], 778))
        open(p, "wb").close()
        np.savez(p + ".npz", **a)
        files[inst] = p
        for k, v in a.items():
            store[f"in_{inst}_{k}"] = v
    with open("FAST_Cusp_Indices.csv", "a") as f:
        f.write(
            "\t".join(
                ["778", "x", "orb", "0", "0"]
                + ["True", "f", "40", "60"] * 4
            )
            + "\n"
        )
    import configurable_spectrograms.cdf_utils as cu

    cu.filtered_orbits_cache.clear()
    df = load_filtered_orbits()
    ext = {
        f"{i}_linear_log_{k}": v
        for i, (ym, zm) in {"ees": (2900, 456.0), "eeb": (3100, 855.0), "ies": (2500, 56.0), "ieb": (2700, 85.0)}.items()
        for k, v in (("y_min", 0), ("y_max", ym), ("z_min", 0), ("z_max", zm))
    }
    store["extrema_json"] = np.array(json.dumps(ext))
    for tag, ge in (("raw", None), ("given", ext)):
        stubs.reset_recording()
        fig, _ = FAST_plot_instrument_grid(
            files,
            filtered_orbits_df=df,
            orbit_number=778,
            scale_function_y="linear",
            scale_function_z="log",
            show=False,
            colormap="viridis",
            global_extrema=ge,
        )
        assert fig is not None
        _pack_panels(tag, _panels(), store)
    np.savez_compressed(os.path.join(HERE, "inst_grid.npz"), **store)

    # -------------------------------------------------------------- generic set
    from configurable_spectrograms.plotting import generic_plot_spectrogram_set

    store = {}
    cases = {
        "f32": synth.make_file_arrays(rng, "ees", n_time=40, integer_counts=False),
        "f64": synth.make_file_arrays(rng, "ees", n_time=24, integer_counts=False, dtype=np.float64, quirks=True),
        "tep": synth.make_file_arrays(rng, "ees", n_time=24, integer_counts=False, stored_layout="tep"),
    }
    for name, a in cases.items():
        p = os.path.join("one", f"generic_{name}.cdf")
        open(p, "wb").close()
        np.savez(p + ".npz", **a)
        for k, v in a.items():
            store[f"in_{name}_{k}"] = v
        from configurable_spectrograms.cdf_utils import load_fast_cdf_dataset

        ds = load_fast_cdf_dataset(p)
        for zs in ("linear", "log"):
            stubs.reset_recording()
            fig, _ = generic_plot_spectrogram_set(
                [{"x": ds["times"], "y": ds["energy"], "data": ds["data"], "label": name}],
                z_scale=zs,
                colormap="viridis",
                show=False,
            )
            assert fig is not None
            _pack_panels(f"{name}_{zs}", _panels(), store)
        # zoomed variant (center/window) on the f32 case
        if name == "f32":
            stubs.reset_recording()
            generic_plot_spectrogram_set(
                [{"x": ds["times"], "y": ds["energy"], "data": ds["data"]}],
                zoom_center=float(ds["times"][20]),
                zoom_window_seconds=50.0,
                z_scale="log",
                show=False,
            )
            _pack_panels("f32_zoom_log", _panels(), store)
    np.savez_compressed(os.path.join(HERE, "generic_set.npz"), **store)

    # ------------------------------------------------------------- extrema tree
    from configurable_spectrograms.fast.batch_directory import FAST_plot_spectrograms_directory
    from configurable_spectrograms.fast.extrema import compute_global_extrema

    tree = os.path.join(work, "tree")
    os.makedirs(tree)
    os.chdir(tree)
    man = synth.write_fast_directory(
        "./FAST_data",
        6,
        seed=31,
        n_time={"ees": 28, "ies": 28, "eeb": 34, "ieb": 34},
        storm_orbits=(1,),
        missing={3: "ieb"},
        cusp_every=2,
        quirks_every=4,
    )
    shutil.copy(man["csv"], "./FAST_Cusp_Indices.csv")
    cu.filtered_orbits_cache.clear()
    store = {}
    for orbit, fl in man["orbits"].items():
        for inst, p in fl.items():
            with np.load(p + ".npz") as z:
                for k in z.files:
                    store[f"{orbit}_{inst}_{k}"] = z[k]
            store[f"{orbit}_{inst}_relpath"] = np.array(os.path.relpath(p, tree))
    store["csv"] = np.array(open("./FAST_Cusp_Indices.csv").read())
    np.savez_compressed(os.path.join(HERE, "extrema_tree.npz"), **store)

    results = {"combos": [], "pool95": None}
    order = ("ees", "eeb", "ies", "ieb")
    for ys, zs in (("linear", "linear"), ("linear", "log"), ("log", "linear"), ("log", "log")):
        ex = compute_global_extrema("./FAST_data", ys, zs, order, max_percentile=99.0)
        results["combos"].append({"y": ys, "z": zs, "extrema": ex})
    # an independent fresh run at the default percentile, with mins
    ex95 = compute_global_extrema(
        "./FAST_data", "linear", "linear", order, extrema_json_path="./other.json", compute_mins=True
    )
    results["pool95_mins"] = ex95
    # the whole batch driver (fork pool inherits the stubs): which PNGs get written
    stubs.reset_recording()
    os.remove("./FAST_calculated_extrema.json")
    res = FAST_plot_spectrograms_directory(
        "./FAST_data",
        output_base="./FAST_plots/",
        y_scale="linear",
        z_scale="log",
        colormap="cividis",
        max_processing_percentile=99,
        max_workers=2,
        progress_json_path="./progress.json",
    )
    pngs = []
    for d, _s, fs in os.walk("./FAST_plots"):
        for fn in fs:
            pngs.append(os.path.relpath(os.path.join(d, fn), "./FAST_plots"))
    results["batch_status"] = sorted((r["orbit"], r["status"]) for r in res)
    results["batch_pngs"] = sorted(pngs)
    results["batch_extrema"] = json.load(open("./FAST_calculated_extrema.json"))
    results["batch_progress"] = json.load(open("./progress.json"))
    with open(os.path.join(HERE, "extrema_tree.json"), "w") as f:
        json.dump(results, f, indent=1, sort_keys=True)

    # ----------------------------------------------------------------- doctests
    from configurable_spectrograms.cdf_utils import get_timestamps_for_orbit
    from configurable_spectrograms.fast.extrema import _extrema_overrides
    from configurable_spectrograms.percentile_utils import compute_percentile_bounds, round_extrema
    import pandas as pd

    orbits = pd.DataFrame({"orbit": [42], "ees min index": [1], "ees max index": [3]})
    times = np.array([100.0, 200.0, 300.0, 400.0])
    pins = {
        "round_extrema(1234,'up')": round_extrema(1234, "up"),
        "round_extrema(0.0123,'down')": round_extrema(0.0123, "down"),
        "compute_percentile_bounds([[1,2,3,100]],0,100)": list(
            compute_percentile_bounds(np.array([[1.0, 2.0, 3.0, 100.0]]), 0, 100)
        ),
        "compute_percentile_bounds([1,2,3],z_min=-5,z_max=5)": list(
            compute_percentile_bounds(np.array([1.0, 2.0, 3.0]), z_min=-5.0, z_max=5.0)
        ),
        "_extrema_overrides": list(
            _extrema_overrides(
                {"ees_linear_linear_y_max": 1234, "ees_linear_linear_z_min": 0.0123}, "ees", "linear", "linear"
            )
        ),
        "get_timestamps_for_orbit(42)": get_timestamps_for_orbit(orbits, 42, "ees", times),
        "get_timestamps_for_orbit(99)": get_timestamps_for_orbit(orbits, 99, "ees", times),
    }
    with open(os.path.join(HERE, "doctests.json"), "w") as f:
        json.dump(pins, f, indent=1)
    shutil.rmtree(work, ignore_errors=True)
    for fn in sorted(os.listdir(HERE)):
        print(fn, os.path.getsize(os.path.join(HERE, fn)))


if __name__ == "__main__":
    main()
