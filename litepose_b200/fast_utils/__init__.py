"""GPU drop-in for the reference's nano-demo native plugin package (nano_demo/fast_utils):
``fast_utils.plugins`` (find_peaks / find_peaks_out / assign / assign_out, plugins.cpp:111-116)
and ``fast_utils.group`` (the "fast inference" HeatmapParser, group.py:10-47)."""
from . import plugins  # noqa: F401
