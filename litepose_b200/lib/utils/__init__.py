"""Coordinate helpers of the evaluation loop (reference lib/utils/transforms.py), see transforms.py."""
