"""Drop-in for the reference's ``models`` package (reference lib/models/__init__.py).

Put ``<repo>/litepose_b200/lib`` in front of ``<reference>/lib`` on ``sys.path``:
``models.pose_mobilenet`` (what valid.py:130 resolves for MODEL.NAME pose_mobilenet) is
the sm_100a-backed module of this repo; every other ``models.<name>`` is resolved lazily
from the reference tree found further down ``sys.path`` (out-of-scope model zoo)."""
import importlib
import os
import sys

from . import pose_mobilenet  # noqa: F401

_here = os.path.dirname(os.path.abspath(__file__))
for _p in list(sys.path):
    _cand = os.path.join(_p, "models")
    if _p and os.path.isdir(_cand) and os.path.abspath(_cand) != _here and _cand not in __path__:
        __path__.append(_cand)


def __getattr__(name):
    try:
        return importlib.import_module(__name__ + "." + name)
    except ImportError as e:
        raise AttributeError(name) from e
