"""B200-native batch path behind the Configurable-Spectrograms API.

The reference's module layout and public names are kept (``plotting``, ``cusp_marking``,
``percentile_utils``, ``cdf_utils``, ``batch_runner``, ``generic_batch``, ``constants`` and
``fast.{plotting, process_orbit, batch_directory, extrema, orbit_discovery, constants}``); the
arithmetic of the hot path -- collapse, percentiles / extrema, normalisation + colormap lookup --
runs in ``libcsgpu.so`` (hand-written sm_100a CUDA, C ABI in ``include/csgpu.h``).  There is no
CPU fallback.
"""

__version__ = "0.1.0"
