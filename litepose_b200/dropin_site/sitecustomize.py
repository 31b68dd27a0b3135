"""Put this directory (and the repo root) on PYTHONPATH to run the reference's scripts unchanged on the sm_100a path:
the interpreter imports ``sitecustomize`` at start-up, which binds ``models`` / ``core`` / ``utils`` to the mirrors of
litepose_b200/lib before the reference manipulates ``sys.path`` (see litepose_b200/dropin.py)."""
import os
import sys

_root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _root not in sys.path:
    sys.path.insert(0, _root)
from litepose_b200 import dropin as _dropin  # noqa: E402

_dropin.install()
