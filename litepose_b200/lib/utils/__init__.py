"""Drop-in for the reference's ``utils`` package: ``utils.transforms`` is this repo's module (coordinate helpers
without OpenCV, device warp / final predictions); ``utils.utils``, ``utils.vis``, ``utils.zipreader`` resolve from the
reference tree found further down ``sys.path`` and run unchanged."""
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
for _p in list(sys.path):
    _cand = os.path.join(_p, "utils")
    if _p and os.path.isdir(_cand) and os.path.abspath(_cand) != _here and _cand not in __path__ \
            and os.path.exists(os.path.join(_cand, "transforms.py")):
        __path__.append(_cand)
