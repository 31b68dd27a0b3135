"""Drop-in for the reference's ``models`` package (reference lib/models/__init__.py):
``models.pose_mobilenet.get_pose_net`` is what valid.py:130 resolves."""
from . import pose_mobilenet  # noqa: F401
