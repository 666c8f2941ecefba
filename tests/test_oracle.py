"""CPU: pin the oracle (oracle/restate.py) against numpy, the reference's own
doctest pins, and the golden vectors captured from the unmodified reference."""


import numpy as np
import pytest

from oracle import restate as R
from tests.helpers import bits, dataset_from_arrays, load_json, load_npz, panels, same_bits, same_float

PA_GROUPS = {  # CS/fast/constants.py:36-41, row order CS/fast/plotting.py:26-31
    "all": [(0.0, 360.0)],
    "down": [(0.0, 30.0), (330.0, 360.0)],
    "up": [(150.0, 210.0)],
    "perp": [(40.0, 140.0), (210.0, 330.0)],
}


def group_mask(pa, ranges):
    m = np.zeros_like(pa, dtype=bool)
    with np.errstate(invalid="ignore"):
        for lo, hi in ranges:
            m |= (pa >= lo) & (pa <= hi)
    return m


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("shape", [(7, 64, 96), (5, 10, 12), (3, 1, 4), (4, 129, 6), (2, 200, 8), (3, 96, 64), (2, 5, 7)])
def test_nansum_orders_match_numpy(dtype, shape):
    rng = np.random.default_rng(1)
    c = rng.gamma(2.0, 3.0, shape).astype(dtype)
    c[rng.random(shape) < 0.05] = np.nan
    c[0, 0, 0] = -0.0
    assert np.array_equal(bits(R.nansum_layout_a(c)), bits(np.nansum(c, axis=1)))
    sel = rng.random(shape[1]) < 0.5
    assert np.array_equal(bits(R.nansum_layout_a(c, sel)), bits(np.nansum(c[:, sel, :], axis=1)))
    stored = np.ascontiguousarray(np.transpose(c, (0, 2, 1)))
    view = np.transpose(stored, (0, 2, 1))
    assert np.array_equal(bits(R.nansum_layout_b(stored)), bits(np.nansum(view, axis=1)))
    if 8 <= shape[1] <= 128:
        assert np.array_equal(bits(R.nansum_layout_b_vec(stored)), bits(np.nansum(view, axis=1)))


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_nanpercentile_matches_numpy(dtype):
    rng = np.random.default_rng(2)
    for n in [1, 2, 3, 5, 10, 100, 1001, 59200, 300000]:
        for p in [0, 1, 5, 50, 95, 99, 99.5, 100, 33.3]:
            v = rng.gamma(2.0, 30.0, n).astype(dtype)
            if n > 3:
                v[rng.random(n) < 0.1] = np.nan
            with np.errstate(invalid="ignore"):
                assert same_float(np.nanpercentile(v, p), R.nanpercentile(v, p)), (n, p)
    v = rng.poisson(3.0, 59200).astype(dtype)
    v[5], v[6] = np.inf, -np.inf
    for p in [0, 1, 99, 100]:
        with np.errstate(invalid="ignore"):
            assert same_float(np.nanpercentile(v, p), R.nanpercentile(v, p))
    assert np.isnan(R.nanpercentile(np.array([np.nan], dtype=dtype), 50))


def test_nanpercentile_float32_index_at_large_n():
    """At n ~ 3e7 the float32 virtual index is several positions off the float64 one."""
    rng = np.random.default_rng(3)
    v = rng.gamma(2.0, 30.0, 20_000_000).astype(np.float32)
    for p in (95, 99):
        assert float(np.nanpercentile(v, p)) == R.nanpercentile(v, p)


def test_reference_doctest_pins():
    pins = load_json("doctests.json")
    assert R.round_extrema(1234, "up") == pins["round_extrema(1234,'up')"] == 1300.0
    assert R.round_extrema(0.0123, "down") == pins["round_extrema(0.0123,'down')"] == 0.012
    assert list(R.compute_percentile_bounds(np.array([[1.0, 2.0, 3.0, 100.0]]), 0, 100)) == [1.0, 100.0]
    assert pins["compute_percentile_bounds([[1,2,3,100]],0,100)"] == [1.0, 100.0]
    assert list(R.compute_percentile_bounds(np.array([1.0, 2.0, 3.0]), z_min=-5.0, z_max=5.0)) == [-5.0, 5.0]
    got = R.extrema_overrides({"ees_linear_linear_y_max": 1234, "ees_linear_linear_z_min": 0.0123}, "ees", "linear", "linear")
    assert list(got) == pins["_extrema_overrides"] == [None, 1300.0, 0.012, None]
    assert R.extrema_overrides(None, "ees", "linear", "linear") == (None, None, None, None)


def _zoom_from_lines(lines, zoom_minutes=6.25):
    """CS/plotting.py:586-596."""
    if len(lines) == 1:
        return lines[0], zoom_minutes * 60
    c = 0.5 * (lines[0] + lines[1])
    return c, max(zoom_minutes * 60, abs(lines[1] - lines[0]) * 1.5)


def test_panel_restatement_matches_reference_pa_grid():
    g = load_npz("pa_grid.npz")
    ds = dataset_from_arrays({k[3:]: v for k, v in g.items() if k.startswith("in_")})
    lo, hi = g["cusp_idx"]
    lines = [float(ds["times"][lo]), float(ds["times"][hi])]
    center, dur = _zoom_from_lines(lines)
    for zs in ("linear", "log"):
        for variant, zb in (("raw", (None, None)), ("given", (0.0, 460.0))):
            ref = panels(g, f"{variant}_{zs}")
            assert len(ref) == 8
            k = 0
            for name in ("all", "down", "up", "perp"):
                sel = group_mask(ds["pitch_angle"], PA_GROUPS[name])
                pa_data = ds["data"][:, sel, :]
                y_hi = 4000 if variant == "raw" else 2900.0
                emask = (ds["energy"] >= 0) & (ds["energy"] <= y_hi)
                with np.errstate(invalid="ignore", over="ignore"):
                    full = np.nansum(pa_data, axis=1)[:, emask].T
                    vmin, vmax = R.compute_percentile_bounds(full, 1, 99, *zb)
                for zoom in (False, True):
                    kw = dict(center=center, window=dur) if zoom else dict(x_min=ds["times"][0], x_max=ds["times"][-1])
                    got = R.panel(ds["times"], ds["energy"], pa_data, z_scale=zs, z_min=vmin, z_max=vmax, **kw)
                    assert got["mode"] == ref[k]["mode"]
                    assert same_float(got["vmin"], ref[k]["vmin"]) and same_float(got["vmax"], ref[k]["vmax"])
                    assert same_bits(got["matrix"], ref[k]["matrix"])
                    k += 1


def test_panel_restatement_matches_reference_generic():
    g = load_npz("generic_set.npz")
    for name in ("f32", "f64", "tep"):
        ds = dataset_from_arrays({k[len(name) + 4 :]: v for k, v in g.items() if k.startswith(f"in_{name}_")})
        for zs in ("linear", "log"):
            ref = panels(g, f"{name}_{zs}")
            assert len(ref) == 1
            got = R.panel(ds["times"], ds["energy"], ds["data"], y_max=ds["energy"].max(), z_scale=zs)
            assert same_bits(got["matrix"], ref[0]["matrix"])
            assert same_float(got["vmin"], ref[0]["vmin"]) and same_float(got["vmax"], ref[0]["vmax"])


def _tree_files():
    tree = load_npz("extrema_tree.npz")
    orbits = sorted({int(k.split("_")[0]) for k in tree if k[0].isdigit()})
    order = ("ees", "eeb", "ies", "ieb")
    files = []
    for o in orbits:
        per = {}
        for inst in order:
            if f"{o}_{inst}_data" in tree:
                ds = dataset_from_arrays({v: tree[f"{o}_{inst}_{v}"] for v in ("time_unix", "data", "energy", "pitch_angle")})
                per[inst] = (ds["energy"], ds["data"])
        files.append((o, per))
    return files, order


def test_global_extrema_restatement_matches_reference():
    files, order = _tree_files()
    gold = load_json("extrema_tree.json")
    state = {}
    for combo in gold["combos"]:  # CLI order: one shared cache (batch_multi_plot_FAST_spectrograms.py:88-93)
        state = R.global_extrema(files, order, combo["y"], combo["z"], state=state, max_percentile=99.0)
        assert state == combo["extrema"], (combo["y"], combo["z"])
    st = R.global_extrema(files, order, "linear", "linear", max_percentile=95.0, compute_mins=True)
    assert st == gold["pool95_mins"]
    # fresh cache, linear/log: every orbit is scanned with the running max
    st = R.global_extrema(files, order, "linear", "log", max_percentile=99.0)
    assert st == gold["batch_extrema"]


def test_running_max_differs_from_final_pool():
    """The storm orbit makes the prefix running max exceed the final-pool percentile."""
    files, _ = _tree_files()
    gold = load_json("extrema_tree.json")["batch_extrema"]
    pool = []
    for _o, per in files:
        if "ees" in per:
            with np.errstate(invalid="ignore", over="ignore"):
                c = np.nansum(per["ees"][1], axis=1)
            pool.append(c[np.isfinite(c) & (c > 0)])
    final = np.ceil(np.nanpercentile(np.concatenate(pool), 99.0))
    assert gold["ees_linear_log_z_max"] > final


def test_norm_and_colormap_index_semantics():
    m = np.array([[0.0, 0.5, 1.0, 2.0, -1.0, np.nan]], dtype=np.float32)
    idx = R.colormap_index(R.normalize(m, 0.0, 1.0))
    assert idx.tolist() == [[0, 128, 255, R.I_OVER, R.I_UNDER, R.I_BAD]]
    lg = np.array([[1.0, 10.0, 100.0, 1000.0, 0.5]], dtype=np.float32)
    idx = R.colormap_index(R.lognorm(lg, 1.0, 100.0))
    assert idx.tolist() == [[0, 128, 255, R.I_OVER, R.I_UNDER]]
    with pytest.raises(ValueError):
        R.lognorm(lg, 10.0, 1.0)
    with pytest.raises(ValueError):
        R.lognorm(lg, float("nan"), 1.0)


def test_native_float32_log10_flip_rate_is_tiny():
    """Informational bound: the correctly rounded log10 the product defines vs this machine's np.log10."""
    rng = np.random.default_rng(5)
    m = rng.gamma(2.0, 80.0, (96, 4000)).astype(np.float32) + np.float32(0.5)
    a = R.colormap_index(R.lognorm(m, 1.0, 2000.0))
    b = R.colormap_index(R.lognorm(m, 1.0, 2000.0, native_log=True))
    assert np.mean(a != b) < 1e-3


def test_oracle_ref_reproduces_the_golden_batch_run(tmp_path):
    """``oracle/_ref`` (the unmodified reference, copied by ``oracle/make_ref.sh``) driven by
    ``oracle/ref_driver.py`` -- the bench's reference arm -- over the golden 6-orbit tree: statuses, PNG
    tree and extrema JSON equal the fixtures captured from ``/root/reference`` itself, with the rendering
    stand-in on (every PNG is a real image).  A subprocess keeps the cdflib / matplotlib stubs out of
    this interpreter.  Skipped where ``oracle/_ref`` has not been built."""
    import json
    import os
    import subprocess
    import sys

    from oracle import ref_driver as RD
    from tests.test_gpu_api import _write_tree

    if not RD.available():
        pytest.skip("oracle/_ref not built (sh oracle/make_ref.sh needs the reference checkout)")
    _write_tree(tmp_path)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import json, sys, os\n"
        f"sys.path.insert(0, {root!r})\n"
        "from oracle import ref_driver as RD\n"
        f"r = RD.run_directory({str(tmp_path)!r}, workers=2, render='cell', colormap='cividis')\n"
        f"os.chdir({str(tmp_path)!r})\n"
        "pngs = sorted(os.path.relpath(os.path.join(d, f), './FAST_plots_ref') for d, _s, fs in os.walk('./FAST_plots_ref') for f in fs)\n"
        "from PIL import Image\n"
        "sizes = [Image.open(os.path.join('./FAST_plots_ref', p)).size for p in pngs[:3]]\n"
        "print('RESULT' + json.dumps({'status': sorted((x['orbit'], x['status']) for x in r['results']), 'pngs': pngs,\n"
        "      'extrema': json.load(open('FAST_calculated_extrema.json')), 'sizes': sizes}))\n"
    )
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-3000:]
    got = json.loads(next(line for line in out.stdout.splitlines() if line.startswith("RESULT"))[6:])
    gold = load_json("extrema_tree.json")
    assert [tuple(x) for x in got["status"]] == [tuple(x) for x in gold["batch_status"]]
    assert got["pngs"] == gold["batch_pngs"]
    assert got["extrema"] == gold["batch_extrema"]
    assert all(w > 8 and h > 8 for w, h in got["sizes"])


def test_norm_and_colormap_index_against_real_matplotlib():
    """R9 is restated, not pinned, because matplotlib is not installable offline.  This test pins it by
    itself the day ``import matplotlib`` succeeds (skipped until then): ``Normalize`` / ``LogNorm`` followed by
    ``Colormap.__call__(bytes=True)`` must colour every cell like ``restate.normalize | lognorm`` ->
    ``colormap_index`` -> LUT row (reference call sites ``CS/plotting.py:276-287,316-324``)."""
    mpl = pytest.importorskip("matplotlib")
    if not hasattr(mpl, "__version__") or not hasattr(mpl, "colormaps"):
        pytest.skip("a matplotlib stand-in is on the path, not matplotlib")
    from matplotlib import colors

    from oracle import restate as R

    rng = np.random.default_rng(9)
    cmap = mpl.colormaps["viridis"].with_extremes(bad=(0, 0, 0, 0))
    lut = R.lut_with_extremes((np.asarray(cmap(np.arange(256), bytes=True))).astype(np.uint8))
    for dtype in (np.float32, np.float64):
        m = (rng.gamma(0.7, 300.0, (64, 500)) + 1e-3).astype(dtype)
        m[rng.random(m.shape) < 0.02] = np.nan
        for vmin, vmax in ((0.5, 2500.0), (1.0, 1.0e4), (3.0, 900.0)):
            got = np.asarray(cmap(colors.LogNorm(vmin=vmin, vmax=vmax)(m), bytes=True))
            want = lut[R.colormap_index(R.lognorm(m, vmin, vmax, native_log=True))]
            assert np.array_equal(got, want), ("log", dtype, vmin, vmax)
            got = np.asarray(cmap(colors.Normalize(vmin=vmin, vmax=vmax)(np.ma.masked_invalid(m)), bytes=True))
            want = lut[R.colormap_index(R.normalize(m, vmin, vmax))]
            assert np.array_equal(got, want), ("linear", dtype, vmin, vmax)
