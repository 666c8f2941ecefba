"""256-entry colour tables as kernel input.

The LUT is an *input* of the rasteriser (``csg_rasterise``): ``(259, 4)`` uint8 = 256
colours + under + over + bad, i.e. ``cmap(np.arange(256), bytes=True)`` plus matplotlib's
default extremes (under = first, over = last colour, bad = transparent).  With matplotlib
installed the tables are taken from it; otherwise OpenCV's tables of the same name are
used (rounded, not truncated, bytes -- colour parity is not claimed, index parity is);
without either a smooth procedural ramp keeps the pipeline running.
"""

from __future__ import annotations

import numpy as np

_cache: dict[str, np.ndarray] = {}


def with_extremes(lut256: np.ndarray, under=None, over=None, bad=(0, 0, 0, 0)) -> np.ndarray:
    lut = np.zeros((259, 4), dtype=np.uint8)
    lut[:256] = lut256
    lut[256] = lut256[0] if under is None else under
    lut[257] = lut256[255] if over is None else over
    lut[258] = bad
    return lut


def _from_matplotlib(name: str):
    try:
        import matplotlib

        if getattr(matplotlib, "__stub__", False):
            return None
        cmap = matplotlib.colormaps[name]
        cmap._init() if hasattr(cmap, "_init") and not getattr(cmap, "_isinit", True) else None
        table = cmap(np.arange(256), bytes=True)
        under = (np.asarray(cmap.get_under()) * 255).astype(np.uint8)
        over = (np.asarray(cmap.get_over()) * 255).astype(np.uint8)
        bad = (np.asarray(cmap.get_bad()) * 255).astype(np.uint8)
        return with_extremes(table, under, over, bad)
    except Exception:
        return None


def _from_opencv(name: str):
    try:
        import cv2

        code = getattr(cv2, f"COLORMAP_{name.upper()}", None)
        if code is None:
            return None
        bgr = cv2.applyColorMap(np.arange(256, dtype=np.uint8).reshape(256, 1), code).reshape(256, 3)
        rgba = np.concatenate([bgr[:, ::-1], np.full((256, 1), 255, np.uint8)], axis=1)
        return with_extremes(rgba)
    except Exception:
        return None


def _procedural(name: str):
    x = np.linspace(0.0, 1.0, 256)
    phase = (sum(map(ord, name)) % 7) / 7.0
    r = np.clip(1.5 - np.abs(4 * x - 3 + phase * 0.2), 0, 1)
    g = np.clip(1.5 - np.abs(4 * x - 2), 0, 1)
    b = np.clip(1.5 - np.abs(4 * x - 1 - phase * 0.2), 0, 1)
    rgba = np.stack([r, g, b, np.ones_like(x)], axis=1)
    return with_extremes((rgba * 255).astype(np.uint8))


def get_lut(colormap) -> np.ndarray:
    """``(259, 4)`` uint8 table for a colormap name (or pass a ready table through)."""
    if isinstance(colormap, np.ndarray):
        if colormap.shape == (259, 4):
            return colormap.astype(np.uint8)
        if colormap.shape == (256, 4):
            return with_extremes(colormap.astype(np.uint8))
        raise ValueError("colormap table must be (256, 4) or (259, 4) uint8")
    name = str(colormap)
    hit = _cache.get(name)
    if hit is None:
        hit = _from_matplotlib(name)
        if hit is None:
            hit = _from_opencv(name)
        if hit is None:
            hit = _procedural(name)
        _cache[name] = hit
    return hit
