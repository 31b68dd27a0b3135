"""Drop-in for the reference's ``utils`` package: ``utils.transforms`` is this repo's module (coordinate helpers
without OpenCV, device warp / final predictions); ``utils.utils``, ``utils.vis``, ``utils.zipreader`` resolve from the
reference tree found further down ``sys.path`` and run unchanged."""
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
try:
    from .._dynpath import DynPath
except ImportError:
    sys.path.insert(0, os.path.dirname(_here))
    try:
        from _dynpath import DynPath
    finally:
        sys.path.pop(0)
__path__ = DynPath(_here, "utils", marker="transforms.py")
