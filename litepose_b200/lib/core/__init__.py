"""Drop-in for the reference's ``core`` package: ``core.group`` is this repo's device parser;
``core.inference`` / ``core.loss`` / ``core.trainer`` resolve from the reference tree found
further down ``sys.path`` (they are reference-owned and run unchanged on our tensors)."""
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
__path__ = DynPath(_here, "core")
