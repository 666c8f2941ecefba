"""Raster figures: the host-side stand-in for the matplotlib objects the reference draws into.

The reference hands every panel to ``Axes.imshow`` and lets Agg resample, colour-map and
composite it (``plotting.py:280-287,316-324``).  On this path the GPU has already produced
the colour-mapped cells (K3), so a figure here is a grid of finished RGBA rasters plus the
annotations the host still owns (titles, labels, ticks, cusp markers), recorded through the
same small slice of the Axes / Figure API the reference touches.  ``savefig`` composes the
panels at cell resolution into one PNG (``png.py``); resampling to the reference's display
resolution is the next item on the scope list.

The classes are duck-typed like their matplotlib namesakes so the mirrored functions
(``plotting.make_spectrogram`` ...) and ``cusp_marking`` read like the reference's.
"""

from __future__ import annotations

import numpy as np

from . import png

_COLORS = {
    "black": (0, 0, 0, 255),
    "red": (255, 0, 0, 255),
    "white": (255, 255, 255, 255),
    "blue": (0, 0, 255, 255),
    "green": (0, 128, 0, 255),
}


def _rgba(color) -> tuple[int, int, int, int]:
    if isinstance(color, str):
        return _COLORS.get(color, (0, 0, 0, 255))
    c = tuple(color)
    if all(0.0 <= float(v) <= 1.0 for v in c):
        c = tuple(int(round(float(v) * 255)) for v in c)
    return (int(c[0]), int(c[1]), int(c[2]), int(c[3]) if len(c) > 3 else 255)


class _Label:
    def __init__(self):
        self.text, self.fontsize = "", None

    def set_fontsize(self, size):
        self.fontsize = size


class _AxisSide:
    def __init__(self):
        self.label = _Label()
        self.major_formatter = None

    def set_major_formatter(self, formatter):
        self.major_formatter = formatter


class DeviceRaster:
    """A panel that stays in HBM: where its RGBA raster lives in a batch's ``d_rgba`` buffer
    (pixel offset, energies, time steps).  Stands in for the host array in ``imshow``; figures made
    of such panels are composed and PNG-encoded on the device (``png.encode_figures_device``)."""

    __slots__ = ("offset", "ne", "nt")

    def __init__(self, offset: int, ne: int, nt: int):
        self.offset, self.ne, self.nt = int(offset), int(ne), int(nt)

    @property
    def shape(self):
        return (self.ne, self.nt, 4)

    @property
    def size(self):
        return self.ne * self.nt * 4


class RasterImage:
    """What ``imshow`` returned: one colour-mapped panel."""

    def __init__(self, rgba, index, extent, cmap, vmin, vmax, norm):
        self.rgba, self.index, self.extent = rgba, index, extent
        self.cmap, self.vmin, self.vmax, self.norm = cmap, vmin, vmax, norm


class Colorbar:
    def __init__(self, image, label, ticks=None, fmt=None):
        self.image, self.label, self.ticks, self.format = image, label, ticks, fmt
        self.ax = PanelAxes(None)


class PanelAxes:
    """One subplot: a raster, its axes metadata and the marker primitives drawn on top."""

    def __init__(self, figure):
        self.figure = figure
        self.images: list[RasterImage] = []
        self.lines: list[dict] = []
        self.texts: list[dict] = []
        self.xaxis, self.yaxis = _AxisSide(), _AxisSide()
        self.title = ""
        self._xlim = (0.0, 1.0)
        self.yticks = None
        self.yticklabels = None
        self.yscale = "linear"
        self.tick_params_calls: list[dict] = []

    # -- the Axes calls the path makes (plotting.py:234-387, cusp_marking.py)
    def imshow(self, rgba, aspect="auto", origin="lower", extent=None, cmap=None, norm=None, vmin=None, vmax=None,
               index=None):
        """Store an already colour-mapped (E', T', 4) uint8 raster (row 0 = lowest energy)."""
        img = RasterImage(rgba if isinstance(rgba, DeviceRaster) else np.asarray(rgba), index, extent, cmap, vmin, vmax, norm)
        self.images.append(img)
        return img

    def set_xlim(self, left, right):
        self._xlim = (float(left), float(right))

    def get_xlim(self):
        return self._xlim

    def set_xlabel(self, text, **kw):
        self.xaxis.label.text = text

    def set_ylabel(self, text, **kw):
        self.yaxis.label.text = text
        if "fontsize" in kw:
            self.yaxis.label.fontsize = kw["fontsize"]

    def set_title(self, text, **kw):
        self.title = text

    def set_yticks(self, ticks):
        self.yticks = list(ticks)

    def set_yticklabels(self, labels):
        self.yticklabels = list(labels)

    def set_yscale(self, scale):
        self.yscale = scale

    def tick_params(self, **kw):
        self.tick_params_calls.append(kw)

    def axvline(self, x, **kw):
        line = {"kind": "vline", "x": float(x), **kw}
        self.lines.append(line)
        return line

    def plot(self, xs, ys, **kw):
        line = {"kind": "polyline", "x": [float(v) for v in xs], "y": [float(v) for v in ys], **kw}
        self.lines.append(line)
        return (line,)

    def text(self, x, y, s, **kw):
        entry = {"x": float(x), "y": float(y), "text": s, **kw}
        self.texts.append(entry)
        return entry

    def get_xaxis_transform(self):
        return "xaxis"  # x in data units, y in axes fraction

    # -- composition
    def marker_columns(self) -> list[tuple[int, int, tuple[int, int, int, int]]]:
        """(column, half width, colour) of every vertical marker burnt into the panel, in drawing order."""
        if not self.images:
            return []
        img = self.images[-1]
        n_cols = img.rgba.shape[1]
        out = []
        if img.extent is not None and n_cols > 0:
            x0, x1 = float(img.extent[0]), float(img.extent[1])
            span = (x1 - x0) or 1.0
            for ln in self.lines:
                if ln["kind"] != "vline":
                    continue
                col = int(round((ln["x"] - x0) / span * (n_cols - 1)))
                if 0 <= col < n_cols:
                    half = 1 if float(ln.get("linewidth", 1)) >= 4 else 0
                    out.append((col, half, _rgba(ln.get("color", "black"))))
        return out

    def render(self) -> np.ndarray | None:
        """(rows, cols, 4) uint8, image row 0 at the TOP, vertical markers burnt in."""
        if not self.images:
            return None
        img = self.images[-1]
        if isinstance(img.rgba, DeviceRaster):
            raise TypeError("this panel lives on the device: encode the figure with png.encode_figures_device")
        out = np.ascontiguousarray(img.rgba[::-1])  # origin="lower": flip for top-down image rows
        for col, half, colour in self.marker_columns():
            out[:, max(0, col - half) : col + half + 1] = colour
        return out


class FigureCanvas:
    def __init__(self, figure):
        self.figure = figure
        figure.canvas = self


class SpectrogramFigure:
    """A grid of panels with the Figure calls the path makes (plotting.py:69-87,461-497,606-693)."""

    _next_number = 1

    def __init__(self, figsize=(10, 3)):
        self.figsize = tuple(figsize)
        self.axes: list[PanelAxes] = []
        self._grid: dict[int, tuple[int, int, int]] = {}
        self.colorbars: list[Colorbar] = []
        self.suptitle_text = None
        self.texts: list[dict] = []
        self.canvas = None
        self.number = SpectrogramFigure._next_number
        SpectrogramFigure._next_number += 1

    def add_subplot(self, n_rows, n_cols, index):
        ax = PanelAxes(self)
        self.axes.append(ax)
        self._grid[id(ax)] = (int(n_rows), int(n_cols), int(index))
        return ax

    def colorbar(self, image, ax=None, label=None, ticks=None, format=None):
        cb = Colorbar(image, label, ticks, format)
        self.colorbars.append(cb)
        return cb

    def suptitle(self, text, **kw):
        self.suptitle_text = text

    def text(self, x, y, s, **kw):
        self.texts.append({"x": x, "y": y, "text": s, **kw})

    def tight_layout(self, **kw):
        pass

    def subplots_adjust(self, **kw):
        pass

    def delaxes(self, ax):
        if ax in self.axes:
            self.axes.remove(ax)
            self._grid.pop(id(ax), None)

    def clf(self):
        self.axes, self._grid, self.colorbars, self.texts = [], {}, [], []

    def layout(self, row_height: int = 148, gap: int = 8):
        """Geometry of :meth:`compose`: ``(H, W, [(axes, y, x, rep)])`` -- every panel at cell
        resolution (time steps are columns), energy rows repeated to about ``row_height`` pixels,
        laid out on the subplot grid with ``gap`` pixels in between.  Shapes only, so it serves
        host rasters and device-resident ones alike."""
        cells = {}
        n_rows = n_cols = 1
        for ax in self.axes:
            r, c, idx = self._grid[id(ax)]
            n_rows, n_cols = max(n_rows, r), max(n_cols, c)
            if not ax.images:
                continue
            ne, nt = ax.images[-1].rgba.shape[:2]
            if ne * nt == 0:
                continue
            rep = max(1, row_height // ne)
            cells[((idx - 1) // c, (idx - 1) % c)] = (ax, ne * rep, nt, rep)
        if not cells:
            return 1, 1, []
        heights = [max([v[1] for (r, _c), v in cells.items() if r == i] or [0]) for i in range(n_rows)]
        widths = [max([v[2] for (_r, c), v in cells.items() if c == j] or [0]) for j in range(n_cols)]
        H = sum(heights) + gap * (n_rows + 1)
        W = sum(widths) + gap * (n_cols + 1)
        placed = []
        y = gap
        for i in range(n_rows):
            x = gap
            for j in range(n_cols):
                v = cells.get((i, j))
                if v is not None:
                    placed.append((v[0], y, x, v[3]))
                x += widths[j] + gap
            y += heights[i] + gap
        return H, W, placed

    def compose(self, row_height: int = 148, gap: int = 8, background=(255, 255, 255, 255)) -> np.ndarray:
        """The figure as one (H, W, 4) uint8 image (see :meth:`layout`)."""
        H, W, placed = self.layout(row_height, gap)
        canvas = np.empty((H, W, 4), dtype=np.uint8)
        canvas[:] = background
        for ax, y, x, rep in placed:
            p = np.repeat(ax.render(), rep, axis=0)
            canvas[y : y + p.shape[0], x : x + p.shape[1]] = p
        return canvas

    def savefig(self, path, dpi=None, compress_level: int = 6, **kw):
        png.write_rgba(path, self.compose(), compress_level=compress_level)


def close_all_axes_and_clear(fig) -> None:
    """Release a figure's panels (reference ``plotting.py:69-89``)."""
    if fig is None:
        return
    try:
        for ax in list(getattr(fig, "axes", [])):
            fig.delaxes(ax)
        fig.clf()
    except Exception:
        pass
