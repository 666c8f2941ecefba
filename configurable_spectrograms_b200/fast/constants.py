"""FAST-specific configuration surface.

The names and values are the reference's (``fast/constants.py:11-41``) because user code and the
mirrored entry points import them.  The pitch-angle grouping is kept as one table of (name, closed
degree intervals) from which both the reference's dict and the row order are derived, so it can also be
handed to the GPU collapse as membership bits (``pipeline.pitch_angle_bits``).
"""

from .. import constants as _generic

# ---- where the FAST batch reads and writes (all relative to the working directory)
FAST_CDF_DATA_FOLDER_PATH = "./FAST_data/"
FAST_OUTPUT_BASE = "./FAST_plots/"
FAST_FILTERED_ORBITS_CSV_PATH = "./FAST_Cusp_Indices.csv"
FAST_EXTREMA_JSON_PATH = "./FAST_calculated_extrema.json"
FAST_PLOTTING_PROGRESS_JSON = "./batch_multi_plot_FAST_progress.json"
FAST_LOGFILE_PREFIX = "./batch_multi_plot_FAST_log"
FAST_LOGFILE_DATETIME_MARKER_PATH = "./batch_multi_plot_FAST_logfile_datetime.txt"

# ---- one colormap per (y scale, z scale): the generic table under both of the reference's names
COLORMAP_LINEAR_Y_LINEAR_Z = DEFAULT_COLORMAP_LINEAR_Y_LINEAR_Z = _generic.COLORMAP_LINEAR_Y_LINEAR_Z
COLORMAP_LINEAR_Y_LOG_Z = DEFAULT_COLORMAP_LINEAR_Y_LOG_Z = _generic.COLORMAP_LINEAR_Y_LOG_Z
COLORMAP_LOG_Y_LINEAR_Z = DEFAULT_COLORMAP_LOG_Y_LINEAR_Z = _generic.COLORMAP_LOG_Y_LINEAR_Z
COLORMAP_LOG_Y_LOG_Z = DEFAULT_COLORMAP_LOG_Y_LOG_Z = _generic.COLORMAP_LOG_Y_LOG_Z

FAST_COLLAPSE_FUNCTION = _generic.COLLAPSE_FUNCTION  # the GPU nansum (engine.nansum)
COLLAPSE_FUNCTION = _generic.COLLAPSE_FUNCTION
CDF_VARIABLES = tuple(_generic.CDF_VARIABLE_NAMES)
DEFAULT_INSTRUMENT_ORDER = ("ees", "eeb", "ies", "ieb")

# ---- pitch-angle groups: (name, closed degree intervals), in the row order of the pitch-angle grid
# (reference ``fast/plotting.py:26-31``).  210 degrees sits in two groups, 30-40 / 140-150 in none.
_GROUPS = (
    ("all", ((0, 360),)),
    ("downgoing", ((0, 30), (330, 360))),
    ("upgoing", ((150, 210),)),
    ("perpendicular", ((40, 140), (210, 330))),
)


def _label(name, spans):
    return name + "\n" + ", ".join(f"({lo}, {hi})" for lo, hi in spans)


PITCH_ANGLE_ROW_KEYS = tuple(_label(name, spans) for name, spans in _GROUPS)
# the reference's dict lists the three partial groups first and "all" last
DEFAULT_PITCH_ANGLE_CATEGORIES: dict[str, list[tuple[float, float]]] = {
    _label(name, spans): [(float(lo), float(hi)) for lo, hi in spans] for name, spans in (_GROUPS[1:] + _GROUPS[:1])
}
