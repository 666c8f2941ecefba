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
only) and by the optional live-reference tests.
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
        self.axes.append(ax)
        return ax

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
