"""Configuration surface shared by the generic and FAST paths.

Every name and value of the reference's ``constants.py`` (``:5-34``) is available here, so user
code that imports them keeps working.  The one entry that is not a plain value is
``COLLAPSE_FUNCTION`` -- the pluggable operator of the hot path (reference: ``np.nansum``,
``constants.py:12``) -- which resolves to the GPU collapse with numpy's summation order.
"""

from .engine import nansum as COLLAPSE_FUNCTION  # (cube, axis=1) -> 2-D sums, bit-exact vs np.nansum

# input: where the cubes live and which CDF variables a cube is made of
CDF_DATA_DIRECTORY, CDF_VARIABLE_NAMES = "./FAST_data/", ["time_unix", "data", "energy", "pitch_angle"]

# colormap per (y scale, z scale) combination
COLORMAP_LINEAR_Y_LINEAR_Z = "viridis"
COLORMAP_LINEAR_Y_LOG_Z = "cividis"
COLORMAP_LOG_Y_LINEAR_Z = "plasma"
COLORMAP_LOG_Y_LOG_Z = "inferno"

# figure geometry and fonts of the single-panel plot (inches / points), default zoom window (minutes)
PLOT_FIGURE_WIDTH_INCHES, PLOT_FIGURE_HEIGHT_INCHES = 6.25, 2.0
TICK_LABEL_FONT_SIZE, AXIS_LABEL_FONT_SIZE = 15, 18
DEFAULT_ZOOM_WINDOW_MINUTES = 6

# generic batch bookkeeping: cusp table, progress file, output tree
FILTERED_ORBITS_CSV_PATH = "./FAST_Cusp_Indices.csv"
PLOTTING_PROGRESS_JSON_PATH = "./batch_multi_plot_progress.json"
OUTPUT_BASE_DIRECTORY = "./plots/"
