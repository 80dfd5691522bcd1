"""Stub of `matplotlib` (absent here); the reference only draws PNGs with it (fast2q.py:1416-1511)."""
