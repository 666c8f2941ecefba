"""FAST-specific paths, instrument order, colormaps and pitch-angle groups
(mirror of the reference's ``fast/constants.py:11-41``)."""

from ..constants import (
    COLLAPSE_FUNCTION,
    COLORMAP_LINEAR_Y_LINEAR_Z,
    COLORMAP_LINEAR_Y_LOG_Z,
    COLORMAP_LOG_Y_LINEAR_Z,
    COLORMAP_LOG_Y_LOG_Z,
)

FAST_CDF_DATA_FOLDER_PATH = "./FAST_data/"
FAST_FILTERED_ORBITS_CSV_PATH = "./FAST_Cusp_Indices.csv"
FAST_PLOTTING_PROGRESS_JSON = "./batch_multi_plot_FAST_progress.json"
FAST_OUTPUT_BASE = "./FAST_plots/"
FAST_LOGFILE_PREFIX = "./batch_multi_plot_FAST_log"
FAST_LOGFILE_DATETIME_MARKER_PATH = "./batch_multi_plot_FAST_logfile_datetime.txt"
FAST_EXTREMA_JSON_PATH = "./FAST_calculated_extrema.json"

FAST_COLLAPSE_FUNCTION = COLLAPSE_FUNCTION
CDF_VARIABLES = ("time_unix", "data", "energy", "pitch_angle")
DEFAULT_INSTRUMENT_ORDER = ("ees", "eeb", "ies", "ieb")

DEFAULT_COLORMAP_LINEAR_Y_LINEAR_Z = COLORMAP_LINEAR_Y_LINEAR_Z
DEFAULT_COLORMAP_LINEAR_Y_LOG_Z = COLORMAP_LINEAR_Y_LOG_Z
DEFAULT_COLORMAP_LOG_Y_LINEAR_Z = COLORMAP_LOG_Y_LINEAR_Z
DEFAULT_COLORMAP_LOG_Y_LOG_Z = COLORMAP_LOG_Y_LOG_Z

#: closed degree intervals per category; 210 sits in two groups, 30-40 / 140-150 in none
DEFAULT_PITCH_ANGLE_CATEGORIES: dict[str, list[tuple[float, float]]] = {
    "downgoing\n(0, 30), (330, 360)": [(0.0, 30.0), (330.0, 360.0)],
    "upgoing\n(150, 210)": [(150.0, 210.0)],
    "perpendicular\n(40, 140), (210, 330)": [(40.0, 140.0), (210.0, 330.0)],
    "all\n(0, 360)": [(0.0, 360.0)],
}

#: row order of the pitch-angle grid (reference ``fast/plotting.py:26-31``)
PITCH_ANGLE_ROW_KEYS = (
    "all\n(0, 360)",
    "downgoing\n(0, 30), (330, 360)",
    "upgoing\n(150, 210)",
    "perpendicular\n(40, 140), (210, 330)",
)
