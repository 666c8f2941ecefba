"""CPU tests of the host-side API pieces that need no GPU: PNG round trip, raster figure
composition, cusp markers, batch_runner's resumable progress."""

import functools
import json
from concurrent.futures import ThreadPoolExecutor

import numpy as np


def test_png_round_trip(tmp_path):
    from configurable_spectrograms_b200 import png

    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (37, 53, 4), dtype=np.uint8)
    data = png.encode_rgba(img)
    assert data[:8] == b"\x89PNG\r\n\x1a\n"
    assert np.array_equal(png.decode_rgba(data), img)
    png.write_many([(tmp_path / f"{i}.png", img[i:]) for i in range(3)], max_workers=2)
    assert np.array_equal(png.decode_rgba((tmp_path / "2.png").read_bytes()), img[2:])
    from PIL import Image  # an independent decoder agrees

    assert np.array_equal(np.asarray(Image.open(tmp_path / "0.png")), img)


def test_figure_compose_and_markers():
    from configurable_spectrograms_b200.cusp_marking import draw_cusp_both_markers, draw_cusp_bracket_marker
    from configurable_spectrograms_b200.figure import FigureCanvas, SpectrogramFigure, close_all_axes_and_clear

    fig = SpectrogramFigure(figsize=(24, 6))
    canvas = FigureCanvas(fig)
    assert canvas.figure is fig and fig.canvas is canvas
    axes = [fig.add_subplot(2, 2, k + 1) for k in range(4)]
    for k, ax in enumerate(axes[:3]):
        rgba = np.full((5, 20 + 10 * k, 4), 40 * (k + 1), dtype=np.uint8)
        rgba[0] = 255  # lowest energy row: must end up at the bottom of the image
        ax.imshow(rgba, extent=(0.0, 1.0, 4.0, 4000.0), vmin=1.0, vmax=10.0)
        ax.set_xlim(0.0, 1.0)
    artists = draw_cusp_both_markers(axes[0], [0.25, 0.75], line_color="white")
    assert len(artists) == 5  # two lines per position + the bracket
    assert draw_cusp_bracket_marker(axes[0], []) == []
    one = draw_cusp_bracket_marker(axes[1], [0.5], caption="cusp")
    assert len(one) == 2 and axes[1].texts[-1]["text"] == "cusp"
    img = fig.compose(row_height=20, gap=2)
    assert img.shape[2] == 4 and img.dtype == np.uint8
    panel = axes[0].render()
    assert (panel[-1, 0] == 255).all() and panel.shape == (5, 20, 4)
    assert (panel[:, round(0.25 * 19)] == (255, 255, 255, 255)).all()  # the white cusp line on top of the black one
    close_all_axes_and_clear(fig)
    assert fig.axes == []


def test_run_batch_progress_and_resume(tmp_path):
    from configurable_spectrograms_b200.batch_runner import run_batch

    path = tmp_path / "progress.json"

    def worker(item):
        if item == "boom":
            raise RuntimeError("x")
        return item, ("no_data" if item == "empty" else "ok")

    factory = functools.partial(ThreadPoolExecutor, max_workers=2)
    res = run_batch(["a", "empty", "boom", "b"], worker, factory, progress_json_path=str(path), flush_batch_size=2,
                    install_signal_handlers=False)
    assert sorted(res) == [("a", "ok"), ("b", "ok"), ("boom", "error"), ("empty", "no_data")]
    state = json.load(open(path))
    assert sorted(state["completed_items"]) == ["'a'", "'b'"] and state["errors"] == ["'boom'"]
    assert state["no_data"] == ["'empty'"] and state["last_index"] == 3 and state["schema_version"] == 1
    res = run_batch(["a", "b", "c"], worker, factory, progress_json_path=str(path), install_signal_handlers=False)
    assert res == [("c", "ok")]
    res = run_batch(["a"], worker, factory, progress_json_path=str(path), ignore_progress_json=True, install_signal_handlers=False)
    assert res == [("a", "ok")]


def test_date2num_matches_datetime_path():
    from datetime import datetime, timezone

    from configurable_spectrograms_b200.plotting import date2num
    from oracle import stubs

    rng = np.random.default_rng(1)
    t = 946684800.0 + rng.random(2000) * 3e7
    t[:4] = [946684800.0, 946684800.5, 946684800.0000005, 946684800.9999995]
    ref = np.array([stubs.date2num(datetime.fromtimestamp(float(x), tz=timezone.utc)) for x in t])
    assert np.array_equal(date2num(t).view(np.uint64), ref.view(np.uint64))
    assert date2num(946684800.0) == ref[0]
