"""Extrema rounding and percentile colour bounds (reference ``percentile_utils.py``)."""

from __future__ import annotations

import math

import numpy as np


def round_extrema(value: float | int, direction: str) -> float:
    """Round to two significant digits, away from the data (``:8-44``).

    >>> round_extrema(1234, 'up')
    1300.0
    >>> round_extrema(0.0123, 'down')
    0.012
    """
    if value == 0:
        return 0.0
    step = 10 ** (math.floor(math.log10(abs(value))) - 1)
    scaled = value / step
    if direction == "up":
        return float(math.ceil(scaled) * step)
    if direction == "down":
        return float(math.floor(scaled) * step)
    raise ValueError(f"Invalid direction: {direction}")


def compute_percentile_bounds(matrix, low_percentile=1, high_percentile=99, z_min=None, z_max=None):
    """``(z_min, z_max)``: explicit bounds win, missing ones are ``np.nanpercentile`` of the
    flattened matrix computed on the GPU with numpy's exact arithmetic (``:47-89``).

    >>> import numpy as np
    >>> compute_percentile_bounds(np.array([1.0, 2.0, 3.0]), z_min=-5.0, z_max=5.0)
    (-5.0, 5.0)
    """
    if z_min is not None and z_max is not None:
        return float(z_min), float(z_max)
    from .engine import matrix_percentiles

    lo, hi = matrix_percentiles(np.asarray(matrix), low_percentile, high_percentile)
    return (float(z_min) if z_min is not None else lo), (float(z_max) if z_max is not None else hi)
