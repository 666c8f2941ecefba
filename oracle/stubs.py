"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Two tiny ``sys.modules`` stubs that let the UNMODIFIED reference
(``/root/reference/src/configurable_spectrograms``) run in a container that
has neither ``cdflib`` nor ``matplotlib``:

* ``cdflib.CDF(path)`` context manager whose ``varget(name)`` serves arrays
  from ``path + ".npz"`` (the only cdflib API the reference touches:
  ``cdf_utils.py:180-181,247-251``).
* a *recording* ``matplotlib``: ``Axes.imshow`` appends the matrix and its
  normalisation arguments to ``RECORDED`` -- that tuple is the reference's
  numeric result for one panel (``plotting.py:280-287,316-324``).

Used by ``tests/golden/make_golden.py`` (fixture generation, this container
only), by the optional live-reference tests and by ``oracle/ref_driver.py`` (the
bench's reference arm).

Rendering stand-in (``set_render("cell" | "display")``):
matplotlib is not installable, so for TIMING the reference the stub's ``imshow``
applies the numpy restatement of ``Normalize`` / ``LogNorm`` + the 256-entry LUT
(``oracle/restate.py`` R9) and ``savefig`` composes the panels and writes the PNG
with Pillow (zlib level 6 -- the encoder matplotlib itself uses).  ``"cell"``: panels at
cell resolution, energy rows repeated to ~148 px (the geometry the product's own PNGs
have); ``"display"``: nearest-neighbour resample into the subplot boxes of a
``figsize x dpi`` canvas (4800 x 2400 for the FAST grids, ``CS/plotting.py:606`` +
``CS/fast/process_orbit.py:110``).  Text, ticks, colourbars and Agg anti-aliasing are NOT
drawn in either mode, so the timing is optimistic for the reference.  With rendering off
(the default) the stub only records -- the golden fixtures depend on that.
"""

from __future__ import annotations

import sys
import types
from datetime import datetime, timezone

import numpy as np

REFERENCE_SRC = "/root/reference/src"

#: panels captured since the last ``reset_recording()``
RECORDED: list[dict] = []
#: figures saved since the last reset: (path, dpi, n_axes)
SAVED: list[tuple] = []


#: None = record only; "cell" / "display" = colour-map in imshow and write real PNGs in savefig
RENDER: dict = {"mode": None, "native_log": True}


def set_render(mode: str | None, native_log: bool = True) -> None:
    if mode not in (None, "cell", "display"):
        raise ValueError(mode)
    RENDER["mode"], RENDER["native_log"] = mode, native_log


_LUTS: dict = {}


def _lut_for(name) -> np.ndarray:
    """A deterministic smooth (259, 4) uint8 table per colormap name (the real tables ship with
    matplotlib, which is absent; the cost of the lookup does not depend on the colours)."""
    key = str(name)
    lut = _LUTS.get(key)
    if lut is None:
        from oracle import restate as R

        seed = sum(ord(c) for c in key) % 7
        x = np.linspace(0.0, 1.0, 256)
        rgb = np.stack([0.5 + 0.5 * np.cos(2 * np.pi * (x * (0.7 + 0.1 * seed) + ph)) for ph in (0.0, 0.33, 0.67)], axis=1)
        table = np.concatenate([(rgb * 255).astype(np.uint8), np.full((256, 1), 255, np.uint8)], axis=1)
        lut = _LUTS[key] = R.lut_with_extremes(table)
    return lut


def reset_recording() -> None:
    RECORDED.clear()
    SAVED.clear()


# ----------------------------------------------------------------------------
# cdflib
# ----------------------------------------------------------------------------
class _CDF:
    def __init__(self, path):
        self._path = str(path)
        self._npz = None

    def __enter__(self):
        self._npz = np.load(self._path + ".npz")
        return self

    def __exit__(self, *exc):
        if self._npz is not None:
            self._npz.close()
        return False

    def varget(self, name):
        if self._npz is None:
            self.__enter__()
        return self._npz[name]


# ----------------------------------------------------------------------------
# matplotlib (recording)
# ----------------------------------------------------------------------------
_EPOCH = datetime(1970, 1, 1, tzinfo=timezone.utc)


def date2num(d):
    """Days since 1970-01-01 UTC, following matplotlib's ``_dt64_to_ordinalf``."""

    def one(x):
        if x.tzinfo is not None:
            x = x.astimezone(timezone.utc).replace(tzinfo=None)
        d64 = np.datetime64(x, "us")
        dsec = d64.astype("datetime64[s]")
        extra = (d64 - dsec).astype("timedelta64[ns]")
        dt = (dsec - np.datetime64("1970-01-01T00:00:00", "s")).astype(np.float64)
        dt += extra.astype(np.float64) / 1.0e9
        return dt / 86400.0

    if isinstance(d, datetime):
        return one(d)
    arr = np.asarray(d, dtype=object)
    if arr.size == 0:
        return np.array([], dtype=np.float64)
    return np.array([one(x) for x in arr.ravel()], dtype=np.float64).reshape(arr.shape)


def num2date(x, tz=None):
    from datetime import timedelta

    return _EPOCH + timedelta(days=float(x))


class DateFormatter:
    def __init__(self, fmt, tz=None):
        self.fmt = fmt


class LogNorm:
    def __init__(self, vmin=None, vmax=None, clip=False):
        self.vmin, self.vmax = vmin, vmax
        # matplotlib validates lazily (at draw); mirror the two raises that matter
        # so reference control flow that would fail under real matplotlib is visible.


class _Label:
    def set_fontsize(self, *_a, **_k):
        pass


class _AxisObj:
    def __init__(self):
        self.label = _Label()

    def set_major_formatter(self, *_a, **_k):
        pass


class _Colorbar:
    def __init__(self):
        self.ax = _Axes(None)


class _Axes:
    def __init__(self, figure):
        self.figure = figure
        self._xlim = (0.0, 1.0)
        self.xaxis = _AxisObj()
        self.yaxis = _AxisObj()
        self.calls: list[tuple] = []

    def set_xlim(self, a, b=None):
        self._xlim = (a, b)

    def get_xlim(self):
        return self._xlim

    def imshow(self, matrix, aspect=None, origin=None, extent=None, cmap=None, norm=None, vmin=None, vmax=None):
        if RENDER["mode"] is not None:
            from oracle import restate as R

            m = np.asarray(matrix)
            with np.errstate(all="ignore"):
                if norm is not None:
                    x = R.lognorm(m, norm.vmin, norm.vmax, native_log=RENDER["native_log"])
                else:
                    x = R.normalize(m, vmin, vmax)
                self.image = _lut_for(cmap)[R.colormap_index(x)]
            self.extent = tuple(float(v) for v in extent) if extent is not None else None
            return {"rendered": True}
        rec = {
            "matrix": np.array(matrix, copy=True),
            "cmap": cmap,
            "extent": tuple(float(v) for v in extent) if extent is not None else None,
            "origin": origin,
            "aspect": aspect,
        }
        if norm is not None:
            rec.update(mode="log", vmin=norm.vmin, vmax=norm.vmax)
        else:
            rec.update(mode="linear", vmin=vmin, vmax=vmax)
        RECORDED.append(rec)
        return rec

    def axvline(self, *a, **k):
        self.calls.append(("axvline", a, k))
        return object()

    def plot(self, *a, **k):
        self.calls.append(("plot", a, k))
        return (object(),)

    def text(self, *a, **k):
        self.calls.append(("text", a, k))
        return object()

    def get_xaxis_transform(self):
        return None

    def __getattr__(self, name):
        # set_xlabel / set_ylabel / set_title / set_yticks / set_yticklabels /
        # set_yscale / tick_params ... : accept and ignore.
        if name.startswith(("set_", "tick_params")):
            return lambda *a, **k: None
        raise AttributeError(name)


class Figure:
    _count = 0

    def __init__(self, figsize=None, **_k):
        self.figsize = figsize
        self.axes: list[_Axes] = []
        self.canvas = None
        Figure._count += 1
        self.number = None

    def add_subplot(self, *a, **k):
        ax = _Axes(self)
        ax.grid = tuple(int(v) for v in a[:3]) if len(a) >= 3 else (1, 1, 1)
        self.axes.append(ax)
        return ax

    def _compose(self, dpi):
        """The figure as one RGBA image: panels (row 0 = lowest energy -> flipped) with their vertical
        cusp lines, on the subplot grid."""
        cells, n_rows, n_cols = {}, 1, 1
        for ax in self.axes:
            img = getattr(ax, "image", None)
            r, c, idx = getattr(ax, "grid", (1, 1, 1))
            n_rows, n_cols = max(n_rows, r), max(n_cols, c)
            if img is None or img.size == 0:
                continue
            panel = np.ascontiguousarray(img[::-1])
            ext = getattr(ax, "extent", None)
            if ext is not None and panel.shape[1] > 0:
                span = (ext[1] - ext[0]) or 1.0
                for kind, a, k in ax.calls:
                    if kind != "axvline" or not a:
                        continue
                    col = int(round((float(a[0]) - ext[0]) / span * (panel.shape[1] - 1)))
                    if 0 <= col < panel.shape[1]:
                        half = 1 if float(k.get("linewidth", 1)) >= 4 else 0
                        panel[:, max(0, col - half) : col + half + 1] = (255, 0, 0, 255) if k.get("color") == "red" else (0, 0, 0, 255)
            cells[((idx - 1) // c, (idx - 1) % c)] = panel
        if not cells:
            return np.full((1, 1, 4), 255, np.uint8)
        if RENDER["mode"] == "display":
            W = int(round((self.figsize or (10, 3))[0] * (dpi or 100)))
            H = int(round((self.figsize or (10, 3))[1] * (dpi or 100)))
            canvas = np.full((H, W, 4), 255, np.uint8)
            # subplot boxes of a plain grid: 8 % margins, 6 % gaps (tight_layout differs by a few pixels)
            bw, bh = 0.84 * W / n_cols, 0.84 * H / n_rows
            for (i, j), panel in cells.items():
                x0, y0 = int(0.08 * W + j * bw + 0.03 * bw), int(0.08 * H + i * bh + 0.03 * bh)
                w, h = int(0.94 * bw), int(0.94 * bh)
                rows = (np.arange(h) * panel.shape[0]) // h
                cols = (np.arange(w) * panel.shape[1]) // w
                canvas[y0 : y0 + h, x0 : x0 + w] = panel[rows][:, cols]
            return canvas
        gap, row_height = 8, 148
        reps = {k: max(1, row_height // p.shape[0]) for k, p in cells.items()}
        heights = [max([cells[k].shape[0] * reps[k] for k in cells if k[0] == i] or [0]) for i in range(n_rows)]
        widths = [max([cells[k].shape[1] for k in cells if k[1] == j] or [0]) for j in range(n_cols)]
        canvas = np.full((sum(heights) + gap * (n_rows + 1), sum(widths) + gap * (n_cols + 1), 4), 255, np.uint8)
        y = gap
        for i in range(n_rows):
            x = gap
            for j in range(n_cols):
                p = cells.get((i, j))
                if p is not None:
                    p = np.repeat(p, reps[(i, j)], axis=0)
                    canvas[y : y + p.shape[0], x : x + p.shape[1]] = p
                x += widths[j] + gap
            y += heights[i] + gap
        return canvas

    def colorbar(self, im, ax=None, label=None, ticks=None, format=None):
        if isinstance(im, dict):
            im["colorbar_ticks"] = list(ticks) if ticks is not None else None
        return _Colorbar()

    def delaxes(self, ax):
        if ax in self.axes:
            self.axes.remove(ax)

    def clf(self):
        self.axes.clear()

    def savefig(self, path, dpi=None, **_k):
        if RENDER["mode"] is not None:
            from PIL import Image

            Image.fromarray(self._compose(dpi), "RGBA").save(path, format="PNG", compress_level=6)
            return
        SAVED.append((str(path), dpi, len(self.axes)))
        with open(path, "wb") as f:
            f.write(b"stub-figure")

    def suptitle(self, *a, **k):
        pass

    def tight_layout(self, *a, **k):
        pass

    def subplots_adjust(self, *a, **k):
        pass

    def text(self, *a, **k):
        pass


class FigureCanvasAgg:
    def __init__(self, figure=None):
        self.figure = figure
        if figure is not None:
            figure.canvas = self

    def close(self):
        pass


class _Gcf:
    @staticmethod
    def destroy(num):
        pass


def install(reference_src: str = REFERENCE_SRC) -> None:
    """Register the stubs and put the reference on ``sys.path`` (idempotent)."""
    if "cdflib" not in sys.modules:
        m = types.ModuleType("cdflib")
        m.CDF = _CDF
        sys.modules["cdflib"] = m
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        mpl.use = lambda *a, **k: None
        mpl.__stub__ = True
        colors = types.ModuleType("matplotlib.colors")
        colors.LogNorm = LogNorm
        dates = types.ModuleType("matplotlib.dates")
        dates.date2num = date2num
        dates.num2date = num2date
        dates.DateFormatter = DateFormatter
        helpers = types.ModuleType("matplotlib._pylab_helpers")
        helpers.Gcf = _Gcf
        backends = types.ModuleType("matplotlib.backends")
        agg = types.ModuleType("matplotlib.backends.backend_agg")
        agg.FigureCanvasAgg = FigureCanvasAgg
        figure = types.ModuleType("matplotlib.figure")
        figure.Figure = Figure
        mpl.colors, mpl.dates, mpl._pylab_helpers = colors, dates, helpers
        mpl.backends, mpl.figure = backends, figure
        backends.backend_agg = agg
        sys.modules.update(
            {
                "matplotlib": mpl,
                "matplotlib.colors": colors,
                "matplotlib.dates": dates,
                "matplotlib._pylab_helpers": helpers,
                "matplotlib.backends": backends,
                "matplotlib.backends.backend_agg": agg,
                "matplotlib.figure": figure,
            }
        )
    if reference_src not in sys.path:
        sys.path.insert(0, reference_src)
