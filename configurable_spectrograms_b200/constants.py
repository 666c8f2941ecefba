"""Configuration surface shared by the generic and FAST paths.

Same names and values as the reference's ``constants.py`` (``:5-34``) so user code
that imports them keeps working; ``COLLAPSE_FUNCTION`` is the one pluggable
operator on the hot path and here resolves to the GPU ``nansum``.
"""

from .engine import nansum as _gpu_nansum

CDF_DATA_DIRECTORY = "./FAST_data/"
CDF_VARIABLE_NAMES = ["time_unix", "data", "energy", "pitch_angle"]

#: collapses a 3-D cube to 2-D; reference: ``np.nansum`` (``constants.py:12``)
COLLAPSE_FUNCTION = _gpu_nansum

COLORMAP_LINEAR_Y_LINEAR_Z = "viridis"
COLORMAP_LINEAR_Y_LOG_Z = "cividis"
COLORMAP_LOG_Y_LINEAR_Z = "plasma"
COLORMAP_LOG_Y_LOG_Z = "inferno"

PLOT_FIGURE_WIDTH_INCHES = 6.25
PLOT_FIGURE_HEIGHT_INCHES = 2.0
TICK_LABEL_FONT_SIZE = 15
AXIS_LABEL_FONT_SIZE = 18
DEFAULT_ZOOM_WINDOW_MINUTES = 6

FILTERED_ORBITS_CSV_PATH = "./FAST_Cusp_Indices.csv"
PLOTTING_PROGRESS_JSON_PATH = "./batch_multi_plot_progress.json"
OUTPUT_BASE_DIRECTORY = "./plots/"
