"""Drop-in for the reference's ``core`` package: ``core.group`` is this repo's device parser;
``core.inference`` / ``core.loss`` / ``core.trainer`` resolve from the reference tree found
further down ``sys.path`` (they are reference-owned and run unchanged on our tensors)."""
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
for _p in list(sys.path):
    _cand = os.path.join(_p, "core")
    if _p and os.path.isdir(_cand) and os.path.abspath(_cand) != _here and _cand not in __path__:
        __path__.append(_cand)
